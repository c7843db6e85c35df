"""The MEX gateway (hybrid_gmres_b200/mex/hgmres_mex.cpp) EXECUTED against a mock of the MATLAB MEX API
(tests/mex_mock/mex_mock.cpp: mxArray structs, mwIndex = 64-bit unsigned CSC, mexErrMsgIdAndTxt that
longjmps like MATLAB's).  MATLAB / Octave do not exist in the image, so this is as close to the drop-in
boundary of SURVEY §8b as a test can get: the same binary entry point `mexFunction`, dispatched on
mexFunctionName(), fed the arrays MATLAB would pass, compared with the executed-reference fixtures."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests.golden_util import load_ref

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def mex(tmp_path_factory):
    out = tmp_path_factory.mktemp("mex") / "libhgmex_mock.so"
    libdir = os.path.join(ROOT, "hybrid_gmres_b200")
    subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"),
                    "-I", os.path.join(libdir, "mex", "stub"), os.path.join(ROOT, "tests", "mex_mock", "mex_mock.cpp"),
                    os.path.join(libdir, "mex", "hgmres_mex.cpp"), "-L", libdir, "-lhgmres", f"-Wl,-rpath,{libdir}",
                    "-o", str(out)], check=True, capture_output=True)
    lib = C.CDLL(str(out))
    vp = C.c_void_p
    lib.mock_new_full.restype = vp
    lib.mock_new_full.argtypes = [C.c_size_t, C.c_size_t, vp]
    lib.mock_new_sparse.restype = vp
    lib.mock_new_sparse.argtypes = [C.c_size_t, C.c_size_t, vp, vp, vp]
    lib.mock_new_string.restype = vp
    lib.mock_new_string.argtypes = [C.c_char_p]
    lib.mock_free.argtypes = [vp]
    lib.mock_shape.argtypes = [vp, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    lib.mock_data.restype = C.POINTER(C.c_double)
    lib.mock_data.argtypes = [vp]
    lib.mock_last_error.restype = C.c_char_p
    lib.mock_call.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp), C.c_int, C.POINTER(vp)]
    yield lib
    lib.mock_unload()


def _mx(lib, v):
    if isinstance(v, str):
        return lib.mock_new_string(v.encode())
    if hasattr(v, "tocsc"):
        c = v.tocsc()
        c.sort_indices()
        jc, ir = c.indptr.astype(np.int64), c.indices.astype(np.int64)
        pr = np.ascontiguousarray(c.data, dtype=np.float64)
        return lib.mock_new_sparse(c.shape[0], c.shape[1], jc.ctypes.data, ir.ctypes.data, pr.ctypes.data)
    a = np.asfortranarray(np.atleast_2d(np.asarray(v, dtype=np.float64)))
    if a.shape[0] == 1 and np.ndim(v) == 1:
        a = np.asfortranarray(a.T)  # vectors are columns
    return lib.mock_new_full(a.shape[0], a.shape[1], a.ctypes.data)


def call(lib, name, nlhs, *args, keep=None):
    """-> list of numpy outputs, or raises RuntimeError with the MATLAB error text"""
    prhs = [(keep[i] if keep and i in keep else _mx(lib, a)) for i, a in enumerate(args)]
    arr = (C.c_void_p * len(prhs))(*prhs)
    plhs = (C.c_void_p * max(nlhs, 1))()
    rc = lib.mock_call(name.encode(), nlhs, plhs, len(prhs), arr)
    try:
        if rc:
            raise RuntimeError(lib.mock_last_error().decode())
        outs = []
        for i in range(max(nlhs, 1)):
            if not plhs[i]:
                outs.append(None)
                continue
            m, n = C.c_size_t(), C.c_size_t()
            lib.mock_shape(plhs[i], C.byref(m), C.byref(n))
            data = np.ctypeslib.as_array(lib.mock_data(plhs[i]), shape=(m.value * n.value,)).copy() if m.value * n.value else np.zeros(0)
            outs.append(data.reshape((m.value, n.value), order="F"))
            lib.mock_free(plhs[i])
        return outs
    finally:
        for i, p in enumerate(prhs):
            if not (keep and i in keep):
                lib.mock_free(p)


def _stats(lib):
    s = (C.c_int * 4)()
    lib.hgmres_mex_cache_stats(s)
    return list(s)


def test_mex_gateway_runs_the_reference_signatures(mex):
    A, B, b, x_true, tol, maxit, lam, k_gcv, r = load_ref("ref_ct16_perturbed")
    m, n = A.shape
    for fn in ("hybrid_ab_gmres_rtp", "hybrid_ba_gmres_rtp"):
        x, err, res, it = call(mex, fn, 4, A, B, b, x_true, tol, float(maxit), lam)
        assert x.shape == (n, 1) and err.shape == res.shape == (int(r[fn + "_it"]), 1) and it.shape == (1, 1)
        assert int(it[0, 0]) == int(r[fn + "_it"])
        assert np.max(np.abs(res.ravel() - r[fn + "_res"]) / r[fn + "_res"]) < 1e-8
        assert np.linalg.norm(x.ravel() - r[fn + "_x"]) < 1e-8 * np.linalg.norm(r[fn + "_x"])
        (x1,) = call(mex, fn, 1, A, B, b, x_true, tol, float(maxit), lam)  # callers use `~` for the rest
        assert np.array_equal(x1, x)
    for fn in ("hybrid_lsqr_solver", "hybrid_lsmr_solver"):
        x, err, res, it = call(mex, fn, 4, A, b, x_true, tol, float(maxit), lam)
        assert int(it[0, 0]) == int(r[fn + "_it"])
        assert np.max(np.abs(res.ravel()[:8] - r[fn + "_res"][:8]) / r[fn + "_res"][:8]) < 1e-8
    x, err, res, it = call(mex, "lsqr_solver", 4, A, b, x_true, tol, float(maxit))
    assert int(it[0, 0]) == int(r["lsqr_solver_it"])
    x, err, res, ar, it = call(mex, "lsmr_solver", 5, A, b, x_true, tol, float(maxit))
    assert int(it[0, 0]) == int(r["lsmr_solver_it"]) and ar.shape == (int(it[0, 0]), 1)
    x, err, res, ar, it = call(mex, "lsmr_solver", 5, A, b)  # lsmr_solver.m:3-5 defaults, NaN history :28
    assert int(it[0, 0]) == int(r["lsmr_defaults_it"]) and np.all(np.isnan(err))
    # full (dense) inputs, as every n = 32 script of the reference passes them
    Ad, Bd, bd, xd, told, maxd, lamd, _, rd = load_ref("ref_deriv2_n32")
    x, err, res, it = call(mex, "hybrid_ba_gmres_rtp", 4, Ad, Bd, bd, xd, told, float(maxd), lamd)
    assert int(it[0, 0]) == int(rd["hybrid_ba_gmres_rtp_it"])
    assert np.max(np.abs(res.ravel()[:4] - rd["hybrid_ba_gmres_rtp_res"][:4]) / rd["hybrid_ba_gmres_rtp_res"][:4]) < 1e-8


def test_mex_gcv_function_memo_is_keyed_on_content(mex):
    A, B, b, x_true, tol, maxit, lam, k_gcv, r = load_ref("ref_ct20_fan_pixel")
    m = A.shape[0]
    # MATLAB passes the SAME arrays to every call of one fminbnd: keep the mxArrays alive across calls
    keepA, keepB, keepb = _mx(mex, A), _mx(mex, B), _mx(mex, b)
    keep = {1: keepA, 2: keepB, 3: keepb}
    s0 = _stats(mex)
    for t in ("ab", "ba"):
        vals = [call(mex, "gcv_function", 1, float(l), None, None, None, float(m), float(k_gcv), t, keep=keep)[0][0, 0]
                for l in r["gcv_lams"]]
        assert np.max(np.abs(np.array(vals) - r[f"gcv_{t}_vals"]) / r[f"gcv_{t}_vals"]) < 1e-7
    s1 = _stats(mex)
    n_calls = 2 * len(r["gcv_lams"])
    assert s1[3] - s0[3] == 2 and s1[2] - s0[2] == n_calls - 2  # one Arnoldi per type, the rest memo hits
    assert s1[1] - s0[1] == 2                                   # A and B uploaded once
    # a perturbed b IN THE SAME BUFFER (same address, same size): must be a miss, not a stale value
    before = call(mex, "gcv_function", 1, 1e-3, None, None, None, float(m), float(k_gcv), "ba", keep=keep)[0][0, 0]
    np.ctypeslib.as_array(mex.mock_data(keepb), shape=(m,))[:] *= 1.5
    after = call(mex, "gcv_function", 1, 1e-3, None, None, None, float(m), float(k_gcv), "ba", keep=keep)[0][0, 0]
    assert _stats(mex)[3] - s1[3] == 1 and abs(after - before) > 1e-6 * abs(before)
    for p in (keepA, keepB, keepb):
        mex.mock_free(p)


def test_mex_errors_release_state_and_keep_working(mex):
    A, B, b, x_true, tol, maxit, lam, k_gcv, r = load_ref("ref_ct16_perturbed")
    with pytest.raises(RuntimeError, match="hgmres:size"):
        call(mex, "hybrid_ba_gmres_rtp", 4, A, B, b[:-1], x_true, tol, float(maxit), lam)
    with pytest.raises(RuntimeError, match="hgmres:nargin"):
        call(mex, "hybrid_ba_gmres_rtp", 4, A, B, b)
    with pytest.raises(RuntimeError, match="hgmres:error"):  # shape mismatch detected inside the C ABI
        call(mex, "hybrid_ba_gmres_rtp", 4, A, A, b, x_true, tol, float(maxit), lam)
    with pytest.raises(RuntimeError, match="hgmres:name"):
        call(mex, "not_a_reference_function", 1, A)
    # `== 0` breakdown at k = 1: MATLAB raises "Output argument x not assigned" for the .m file too
    I, e1 = np.eye(6), np.eye(6)[:, 0].copy()
    with pytest.raises(RuntimeError, match="not assigned"):
        call(mex, "hybrid_ab_gmres_rtp", 4, I, I, e1, np.ones(6), 1e-6, 4.0, 1e-2)
    x, err, res, it = call(mex, "hybrid_ba_gmres_rtp", 4, I, I, e1, np.ones(6), 1e-6, 4.0, 1e-2)
    assert int(it[0, 0]) == 1 and res[0, 0] == 0.0 and np.all(x == 0)
    # after all those longjmps the gateway still works and its cache is consistent
    x, err, res, it = call(mex, "hybrid_ba_gmres_rtp", 4, A, B, b, x_true, tol, float(maxit), lam)
    assert np.max(np.abs(res.ravel() - r["hybrid_ba_gmres_rtp_res"]) / r["hybrid_ba_gmres_rtp_res"]) < 1e-8
    assert mex.mock_locked() == 1
