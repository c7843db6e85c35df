"""CPU tests of the host-side logic and the C-ABI surface (no GPU compute)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle
from oracle.solvers import gcv_from_H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from hybrid_gmres_b200 import _lib
    return _lib.load()


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "hgmres.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hg_[a-z0-9_A-Z]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    from hybrid_gmres_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/hgmres.h but not exported"
        assert n in _lib.PROTOTYPES, f"{n} has no ctypes prototype"
    assert set(_lib.PROTOTYPES) <= set(names)


def test_no_gpu_fails_loudly(lib):
    import hybrid_gmres_b200 as hg
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(hg._lib.HgError) as e:
        hg.Context(0)
    assert e.value.status == 2 and "no CPU fallback" in str(e.value)
    with pytest.raises(hg._lib.HgError):
        hg.hybrid_ba_gmres_rtp(np.eye(3), np.eye(3), np.ones(3), np.ones(3), 1e-6, 2, 1e-2)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "hybrid_gmres_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f


def test_hessenberg_ls_matches_lstsq(lib):
    rng = np.random.default_rng(0)
    for k in (1, 2, 7, 40):
        H = np.triu(rng.standard_normal((k + 1, k)), -1)
        H[np.arange(1, k + 1), np.arange(k)] = np.abs(H[np.arange(1, k + 1), np.arange(k)]) + 0.5
        Hf = np.asfortranarray(H)
        y = np.zeros(k)
        assert lib.hg_host_hessenberg_ls(Hf.ctypes.data, k + 1, k, 2.5, y.ctypes.data) == 0
        rhs = np.zeros(k + 1)
        rhs[0] = 2.5
        ref = oracle.solvers._mldivide_rect(H, rhs)
        assert np.linalg.norm(y - ref) / np.linalg.norm(ref) < 1e-11 * np.linalg.cond(H)


def test_solve_square_spd_and_general(lib):
    rng = np.random.default_rng(1)
    for n in (1, 3, 25):
        G = rng.standard_normal((n + 5, n))
        M = np.asfortranarray(G.T @ G + 1e-3 * np.eye(n))
        M = 0.5 * (M + M.T)
        rhs = rng.standard_normal(n)
        y = np.zeros(n)
        assert lib.hg_host_solve_square(n, M.ctypes.data, n, rhs.ctypes.data, y.ctypes.data) == 0
        assert np.linalg.norm(M @ y - rhs) / np.linalg.norm(rhs) < 1e-10
        N = np.asfortranarray(rng.standard_normal((n, n)) + 3 * np.eye(n))  # non-symmetric -> LU
        assert lib.hg_host_solve_square(n, N.ctypes.data, n, rhs.ctypes.data, y.ctypes.data) == 0
        assert np.linalg.norm(N @ y - rhs) / np.linalg.norm(rhs) < 1e-10
    # symmetric indefinite: Cholesky fails, LU path is taken
    S = np.asfortranarray(np.array([[1.0, 2.0], [2.0, 1.0]]))
    rhs = np.array([1.0, 0.0])
    y = np.zeros(2)
    lib.hg_host_solve_square(2, S.ctypes.data, 2, rhs.ctypes.data, y.ctypes.data)
    assert np.allclose(S @ y, rhs)


def test_singular_values(lib):
    rng = np.random.default_rng(2)
    for n in (1, 4, 20):
        M = np.asfortranarray(np.triu(rng.standard_normal((n, n)), -1))
        if n == 20:
            M[:, -3:] = 0.0  # trailing zero columns as after a gcv breakdown
        s = np.zeros(n)
        assert lib.hg_host_singular_values(n, M.ctypes.data, n, s.ctypes.data) == 0
        ref = np.linalg.svd(M, compute_uv=False)
        assert np.allclose(s, ref, rtol=1e-12, atol=1e-14)


def _gcv_inputs(name="shaw", gcv_type="ab"):
    from oracle.generators import add_noise
    from oracle.solvers import gcv_arnoldi
    A, b_exact, x_true = oracle.generate_test_problem(name, 32)
    B = A.T.copy()
    b = add_noise(b_exact, 1e-2, 0)
    H, beta = gcv_arnoldi(A, B, b, 32, 20, gcv_type)
    return H, beta


@pytest.mark.parametrize("name", ["shaw", "deriv2", "heat"])
@pytest.mark.parametrize("gcv_type", ["ab", "ba"])
def test_host_gcv_matches_oracle(name, gcv_type):
    import hybrid_gmres_b200 as hg
    H, beta = _gcv_inputs(name, gcv_type)
    g = hg.GcvProblem.from_H(H, beta, 32)
    for lam in np.logspace(-9, -1, 9):
        v, ref = g.eval(lam), gcv_from_H(lam, H, beta, 32)
        assert abs(v - ref) <= 1e-7 * ref
    # 1e20 sentinel (gcv_function.m:56-58)
    H2 = np.zeros((3, 2))
    H2[0, 0] = H2[1, 1] = 1.0
    assert hg.GcvProblem.from_H(H2, 1.0, 2.0).eval(0.0) == 1e20


def test_host_fminbnd_is_matlab_fminbnd():
    """C++ fminbnd == the Python restatement bit for bit when driven by the same
    objective; against the oracle's own GCV objective the evaluation count is
    identical and lambda agrees far inside TolX (analyze_regularization.m:37-40)."""
    import hybrid_gmres_b200 as hg
    H, beta = _gcv_inputs("shaw", "ab")
    g = hg.GcvProblem.from_H(H, beta, 32)
    lam, fval, cnt, trace = g.fminbnd(1e-9, 1e-1, 1e-8)
    tr = []
    lam_p, f_p, flag, cnt_p = oracle.fminbnd(g.eval, 1e-9, 1e-1, 1e-8, trace=tr)
    assert lam == lam_p and fval == f_p and cnt == cnt_p and np.array_equal(trace, np.array(tr))
    lam_o, f_o, flag, cnt_o = oracle.fminbnd(lambda l: gcv_from_H(l, H, beta, 32), 1e-9, 1e-1, 1e-8)
    assert cnt == cnt_o
    assert abs(lam - lam_o) < 1e-8 * 1e-3


def test_argument_validation(lib):
    import hybrid_gmres_b200 as hg
    out = C.c_void_p()
    assert lib.hg_gcv_from_H(None, 3, 2, 1.0, 2.0, C.byref(out)) == 1
    assert b"NULL" in lib.hg_last_error()
    H = np.zeros((3, 2), order="F")
    assert lib.hg_gcv_from_H(H.ctypes.data, 2, 2, 1.0, 2.0, C.byref(out)) == 1  # ldh too small
    with pytest.raises(ValueError):
        hg.GcvProblem.from_H(np.zeros((2, 2)), 1.0, 2.0)


def test_mex_gateway_parses_and_binds_the_abi():
    """The MEX gateway cannot be built without MATLAB; it must at least
    parse against the real C ABI header and the stub mex.h."""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    src = os.path.join(ROOT, "hybrid_gmres_b200", "mex", "hgmres_mex.cpp")
    r = subprocess.run([gxx, "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "include"),
                        "-I", os.path.join(ROOT, "hybrid_gmres_b200", "mex", "stub"), src],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    txt = open(src).read()
    for fn in ("hg_hybrid_ab_gmres_rtp", "hg_hybrid_ba_gmres_rtp", "hg_gcv_prepare", "hg_gcv_eval",
               "hg_hybrid_lsqr_solver", "hg_hybrid_lsmr_solver", "hg_lsqr_solver", "hg_lsmr_solver",
               "hg_matrix_from_csc", "hg_matrix_from_dense"):
        assert fn in txt


def test_gcv_surface_host_matches_oracle():
    """hg_gcv_surface (host-side, from a given H) against plot_gcv_surface.m's compute_gcv_surface."""
    import hybrid_gmres_b200 as hg
    from oracle.gcv_surface import compute_gcv_surface
    from oracle.generators import add_noise
    from oracle.solvers import gcv_arnoldi
    A, b_exact, x_true = oracle.generate_test_problem("shaw", 32)
    rng = np.random.default_rng(0)
    B = A.T + 1e-4 * rng.standard_normal(A.shape)  # plot_gcv_surface.m:14-15
    b = add_noise(b_exact, 1e-2, 0)
    lams = np.logspace(-8, -1, 25)
    for t in ("ab", "ba"):
        surf_o, path_o = compute_gcv_surface(t, A, B, b, 12, range(1, 13), lams)
        H, beta = gcv_arnoldi(A, B, b, 32, 12, t)
        surf, path = hg.GcvProblem.from_H(H, beta, 32).surface(lams, 12)
        assert surf.shape == surf_o.shape
        assert np.allclose(surf, surf_o, rtol=1e-6, atol=0)
        assert np.array_equal(path, path_o)


def test_tile_permutation_keeps_tiles_contiguous():
    """ct.tile_permutation: perm[new] = old column-major pixel index; every run of tile*tile new
    indices is one tile x tile block of the image (one 128-byte line for tile = 4)."""
    from hybrid_gmres_b200.ct import tile_permutation
    for N, tile in ((8, 4), (16, 4), (16, 8), (12, 4)):
        q = tile_permutation(N, tile)
        assert q.dtype == np.int32 and sorted(q.tolist()) == list(range(N * N))
        for t in range(N * N // (tile * tile)):
            old = q[t * tile * tile:(t + 1) * tile * tile]
            rows, cols = old % N, old // N  # column-major: index = row + N*col
            assert rows.max() - rows.min() == tile - 1 and cols.max() - cols.min() == tile - 1
            assert rows.min() % tile == 0 and cols.min() % tile == 0
    with pytest.raises(ValueError):
        tile_permutation(10, 4)


def test_bench_view_interleave_covers_every_view_once():
    """bench.build_workload gives rank r the views r, r+P, ...: a partition of the views whose parts
    differ by at most one view (nnz balance) and share the same mix of directions."""
    for nv, P in ((180, 1), (180, 2), (180, 8), (3600, 8), (181, 4)):
        parts = [np.arange(r, nv, P) for r in range(P)]
        allv = np.sort(np.concatenate(parts))
        assert np.array_equal(allv, np.arange(nv))
        sizes = [len(p) for p in parts]
        assert max(sizes) - min(sizes) <= 1
