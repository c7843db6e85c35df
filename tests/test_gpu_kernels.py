"""GPU parity of the individual kernels and data-format conversions, through the
C ABI, against NumPy/SciPy on the same inputs."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu

RTOL = 1e-13  # FP64 mat-vec in a different summation order


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def test_extension_loaded(hg):
    import ctypes
    assert hg._lib.load().hg_version() >= 100
    assert isinstance(hg._lib.load(), ctypes.CDLL)


@pytest.mark.parametrize("which", ["A", "B"])
def test_spmv_ct(hg, ctx, ct64, which):
    A, B, b, x_true = ct64
    M = A if which == "A" else B
    d = hg.DeviceMatrix.from_any(M, ctx)
    x = np.random.default_rng(1).standard_normal(M.shape[1])
    assert _rel(d.matvec(x), M @ x) < RTOL


@pytest.mark.parametrize("shape,density", [((1, 1), 1.0), ((7, 3), 0.5), ((300, 200), 0.02),
                                           ((64, 5000), 0.3), ((5000, 64), 0.001), ((33, 33), 0.0)])
def test_spmv_ragged_and_empty(hg, ctx, shape, density):
    rng = np.random.default_rng(2)
    M = sp.random(shape[0], shape[1], density=density, format="csr", random_state=rng, dtype=np.float64)
    d = hg.DeviceMatrix.from_any(M, ctx)
    assert d.shape == shape and d.nnz == M.nnz
    x = rng.standard_normal(shape[1])
    y = d.matvec(x)
    ref = M @ x
    assert np.allclose(y, ref, rtol=1e-12, atol=1e-14)


def test_spmv_dense_input(hg, ctx):
    from oracle import generate_test_problem
    A, b, x = generate_test_problem("deriv2", 32)
    d = hg.DeviceMatrix.from_any(A, ctx)
    assert d.nnz == 32 * 32
    assert _rel(d.matvec(x), A @ x) < RTOL
    At = hg.DeviceMatrix.from_any(np.asfortranarray(A.T), ctx)
    assert _rel(At.matvec(b), A.T @ b) < RTOL


def test_transpose_bit_exact(hg, ctx, ct64):
    A = ct64[0]
    d = hg.DeviceMatrix.from_any(A, ctx)
    t = d.transpose()
    indptr, indices, data = t.download()
    ref = A.T.tocsr()
    ref.sort_indices()
    assert np.array_equal(indptr, ref.indptr)
    assert np.array_equal(indices, ref.indices)
    assert np.array_equal(data, ref.data)  # values only move: bit exact
    tt = t.transpose()
    i2, j2, v2 = tt.download()
    A2 = A.copy()
    A2.sort_indices()
    assert np.array_equal(i2, A2.indptr) and np.array_equal(j2, A2.indices) and np.array_equal(v2, A2.data)


def test_transpose_long_rows_and_duplicates(hg, ctx):
    rng = np.random.default_rng(3)
    # one column hit by 20000 rows -> an output row longer than the shared-memory sort
    rows = np.concatenate([np.arange(20000), rng.integers(0, 20000, 5000)])
    cols = np.concatenate([np.zeros(20000, dtype=np.int64), rng.integers(1, 50, 5000)])
    vals = rng.standard_normal(rows.shape[0])
    M = sp.csr_matrix((vals, (rows, cols)), shape=(20000, 50))
    d = hg.DeviceMatrix.from_any(M, ctx)
    indptr, indices, data = d.transpose().download()
    ref = M.T.tocsr()
    ref.sort_indices()
    assert np.array_equal(indptr, ref.indptr) and np.array_equal(indices, ref.indices)
    assert np.array_equal(data, ref.data)


@pytest.mark.parametrize("idx_dtype", [np.int32, np.uint64])
def test_csc_upload_matches_csr(hg, ctx, ct48_unmatched, idx_dtype):
    """MATLAB hands sparse matrices as CSC with 64-bit mwIndex."""
    A = ct48_unmatched[0]
    Ac = A.tocsc()
    Ac.sort_indices()
    d = hg.DeviceMatrix.from_csc(Ac.indptr.astype(idx_dtype), Ac.indices.astype(idx_dtype), Ac.data, A.shape, ctx)
    indptr, indices, data = d.download()
    ref = A.copy()
    ref.sort_indices()
    assert np.array_equal(indptr, ref.indptr) and np.array_equal(indices, ref.indices)
    assert np.array_equal(data, ref.data)


@pytest.mark.parametrize("n,k", [(1, 1), (37, 3), (4096, 8), (100003, 17), (70000, 60)])
def test_multidot_and_lincomb(hg, ctx, n, k):
    import ctypes as C
    rng = np.random.default_rng(4)
    V = np.asfortranarray(rng.standard_normal((n, k)))
    w = rng.standard_normal(n)
    c = rng.standard_normal(k)
    lib = ctx._lib
    h = np.zeros(k)
    hg._lib.check(lib.hg_multidot(ctx._h, n, k, V.ctypes.data, n, w.ctypes.data, h.ctypes.data))
    ref = V.T @ w
    assert np.allclose(h, ref, rtol=1e-12, atol=1e-12 * np.linalg.norm(w))
    out = np.zeros(n)
    nrm2 = C.c_double()
    hg._lib.check(lib.hg_lincomb(ctx._h, n, k, V.ctypes.data, n, c.ctypes.data, -1.0, w.ctypes.data,
                                 out.ctypes.data, C.byref(nrm2)))
    ref = w - V @ c
    assert _rel(out, ref) < 1e-14
    assert abs(nrm2.value - ref @ ref) <= 1e-13 * (ref @ ref)


def test_reductions_are_deterministic(hg, ctx, ct64):
    A = ct64[0]
    d = hg.DeviceMatrix.from_any(A, ctx)
    x = np.random.default_rng(5).standard_normal(A.shape[1])
    y1 = d.matvec(x)
    y2 = d.matvec(x)
    assert np.array_equal(y1, y2)


@pytest.fixture
def spmv_mode(hg):
    def _set(v):
        hg.set_option("spmv_mode", v)
    yield _set
    hg.set_option("spmv_mode", 0)


@pytest.mark.parametrize("case", ["ctA", "ctB", "short_rows", "empty_rows", "one_long_row", "tiny"])
def test_streaming_spmv_matches_rowwarp(hg, ctx, ct64, spmv_mode, case):
    """The TMA-staged streaming kernel (forced with spmv_mode=2) on matrices it is and is not
    tuned for: ragged, empty and very long rows, fewer rows than warps."""
    rng = np.random.default_rng(7)
    if case == "ctA":
        M = ct64[0]
    elif case == "ctB":
        M = ct64[1]
    elif case == "short_rows":
        M = sp.random(20000, 3000, density=0.002, format="csr", random_state=rng)
    elif case == "empty_rows":
        M = sp.random(5000, 4000, density=0.05, format="csr", random_state=rng).tolil()
        M[100:900] = 0
        M[-300:] = 0
        M = M.tocsr()
        M.eliminate_zeros()
    elif case == "one_long_row":
        M = sp.vstack([sp.random(1, 300000, density=0.9, format="csr", random_state=rng),
                       sp.random(50, 300000, density=0.001, format="csr", random_state=rng)]).tocsr()
    else:
        M = sp.csr_matrix(np.array([[1.0, 2.0, 0.0], [0.0, 0.0, 0.0], [0.0, 3.0, 4.0]]))
    d = hg.DeviceMatrix.from_any(M, ctx)
    x = rng.standard_normal(M.shape[1])
    ref = M @ x
    spmv_mode(1)
    y1 = d.matvec(x)
    spmv_mode(2)
    y2 = d.matvec(x)
    y2b = d.matvec(x)
    scale = np.abs(M) @ np.abs(x) + 1e-300
    assert np.max(np.abs(y1 - ref) / scale) < 1e-14
    assert np.max(np.abs(y2 - ref) / scale) < 1e-14
    assert np.array_equal(y2, y2b)  # bit-reproducible


def test_streaming_spmv_in_solvers(hg, ctx, ct64, spmv_mode):
    """Whole BA-RTP / LSQR solves with every SpMV (incl. fused epilogues and square-sums)
    forced through the streaming kernel agree with the row-per-warp path."""
    A, B, b, x_true = ct64
    out = {}
    for mode in (1, 2):
        spmv_mode(mode)
        out[mode] = (hg.hybrid_ba_gmres_rtp(A, B, b, x_true, 1e-6, 25, 1e-2, ctx=ctx, residual_mode=1),
                     hg.hybrid_lsqr_solver(A, b, x_true, 1e-6, 8, 1e-2, ctx=ctx))
    for a, c in zip(out[1], out[2]):
        assert a[3] == c[3]
        assert np.linalg.norm(a[0] - c[0]) / np.linalg.norm(a[0]) < 1e-10
        assert np.max(np.abs(a[2] - c[2]) / a[2]) < 1e-10


# ---------------------------------------------------------------------------
# row-per-lane SpMV over 32-row slices, and n-space permutations
# ---------------------------------------------------------------------------
def _uniform_rows(rows, cols, per_row, rng, ragged=0):
    """CSR matrix with `per_row` entries in every row (minus up to `ragged` in random rows)."""
    lens = np.full(rows, per_row) - (rng.integers(0, ragged + 1, rows) if ragged else 0)
    indptr = np.concatenate([[0], np.cumsum(lens)])
    indices = np.concatenate([np.sort(rng.choice(cols, l, replace=False)) for l in lens]).astype(np.int32)
    return sp.csr_matrix((rng.standard_normal(indices.shape[0]), indices, indptr), shape=(rows, cols))


@pytest.mark.parametrize("rows,per_row,ragged", [(64, 8, 0), (1000, 37, 0), (4099, 64, 1), (257, 130, 0), (333, 50, 7)])
def test_spmv_sliced_form_matches_csr(hg, ctx, rows, per_row, ragged):
    """The row-per-lane kernel (forced: spmv_mode 3; random columns would not select it) against
    SciPy and the row-per-warp kernel, 16- and 32-bit indices, rows not a multiple of 32, ragged tails."""
    rng = np.random.default_rng(10)
    M = _uniform_rows(rows, 700, per_row, rng, ragged)
    x = rng.standard_normal(M.shape[1])
    ys = {}
    try:
        for bits in (8, 16, 32):  # byte offsets are taken when every slice column spans < 256, else 16-bit ones
            hg.set_option("spmv_idx16", 0 if bits == 32 else 1)
            hg.set_option("spmv_idx8", 1 if bits == 8 else 0)
            hg.set_option("spmv_mode", 3)
            d = hg.DeviceMatrix.from_any(M, ctx)
            assert d.spmv_form == "sell32" and d.spmv_index_bits in ((8, 16) if bits == 8 else (bits,))
            y = d.matvec(x)
            assert _rel(y, M @ x) < RTOL
            assert np.array_equal(d.matvec(x), y)  # deterministic
            ys[bits] = y
        assert np.array_equal(ys[8], ys[16])  # same kernel shape, same entry order
        ys = {1: ys[16], 0: ys[32]}
        hg.set_option("spmv_mode", 1)
        d1 = hg.DeviceMatrix.from_any(M, ctx)
        assert d1.spmv_form == "csr"
        y1 = d1.matvec(x)
    finally:
        hg.set_option("spmv_mode", 0)
        hg.set_option("spmv_idx16", 1)
        hg.set_option("spmv_idx8", -1)
    # few rows: the 16-bit kernel lets several warps share a slice (another summation order); at full size
    # the 16- and 32-bit kernels are bit-identical (tests/test_gpu_fullsize.py)
    assert _rel(ys[0], ys[1]) < RTOL
    assert _rel(ys[1], y1) < RTOL


def test_spmv_random_columns_stay_row_per_warp(hg, ctx):
    """Auto mode samples the gather locality: with random columns the sliced traversal gathers no
    fewer lines than the row-per-warp one, so the matrix keeps the CSR kernel."""
    M = _uniform_rows(2048, 50000, 64, np.random.default_rng(17))
    assert hg.DeviceMatrix.from_any(M, ctx).spmv_form == "csr"


def test_spmv_ragged_rows_stay_csr(hg, ctx, ct64):
    A, B = ct64[0], ct64[1]
    dA = hg.DeviceMatrix.from_any(A, ctx)
    assert dA.spmv_form == "csr"  # adjacent rays drift apart entry by entry: no gather locality across lanes


def test_pixel_backprojector_uses_sliced_form(hg, ctx, ct48_unmatched):
    A, B, b, x_true = ct48_unmatched
    dB = hg.DeviceMatrix.from_any(B, ctx)
    assert dB.spmv_form == "sell32"
    u = np.random.default_rng(11).standard_normal(B.shape[1])
    assert _rel(dB.matvec(u), B @ u) < RTOL


def test_permute_bit_exact(hg, ctx, ct48_unmatched):
    from hybrid_gmres_b200.ct import tile_permutation
    A, B = ct48_unmatched[0], ct48_unmatched[1]
    q = tile_permutation(48, 4)
    rng = np.random.default_rng(12)
    r = rng.permutation(A.shape[0]).astype(np.int32)
    for M, rp, cp in ((A, None, q), (B, q, None), (A, r, q)):
        d = hg.DeviceMatrix.from_any(M, ctx).permute(rp, cp)
        ref = M.tocsr()
        if rp is not None:
            ref = ref[rp, :]
        if cp is not None:
            ref = ref[:, cp]
        ref = ref.tocsr()
        ref.sort_indices()
        indptr, indices, data = d.download()
        assert np.array_equal(indptr, ref.indptr)
        assert np.array_equal(indices, ref.indices)
        assert np.array_equal(data, ref.data)
    with pytest.raises(hg._lib.HgError):
        hg.DeviceMatrix.from_any(A, ctx).permute(None, np.zeros(A.shape[1], dtype=np.int32))


def test_permute_keep_entry_order(hg, ctx, ct48_unmatched):
    """HG_PERMUTE_KEEP_ENTRY_ORDER: entries keep their order within the row (new column labels
    only) — what the solvers' nperm path uses; products are those of A(:,q)."""
    from hybrid_gmres_b200.ct import tile_permutation
    A = ct48_unmatched[0]
    q = tile_permutation(48, 4)
    d = hg.DeviceMatrix.from_any(A, ctx).permute(None, q, sort=False)
    indptr, indices, data = d.download()
    inv = np.empty_like(q)
    inv[q] = np.arange(q.shape[0], dtype=q.dtype)
    As = A.tocsr()
    assert np.array_equal(indptr, As.indptr)
    assert np.array_equal(indices, inv[As.indices]) and np.array_equal(data, As.data)
    x = np.random.default_rng(13).standard_normal(A.shape[1])
    assert _rel(d.matvec(x[q]), A @ x) < RTOL


def test_device_buffer_cache_reuses_blocks(hg, ctx):
    """Released matrices >= 1 MB are handed to the next upload of the same size (hg_dmalloc);
    results do not depend on it and hg_ctx_trim empties the cache."""
    rng = np.random.default_rng(14)
    M = sp.random(3000, 3000, density=0.05, format="csr", random_state=rng, dtype=np.float64)
    x = rng.standard_normal(3000)
    ys = []
    for _ in range(3):
        d = hg.DeviceMatrix.from_any(M, ctx)
        ys.append(d.matvec(x))
        d.close()
    assert np.array_equal(ys[0], ys[1]) and np.array_equal(ys[0], ys[2])
    assert _rel(ys[0], M @ x) < RTOL
    ctx.trim()
    d = hg.DeviceMatrix.from_any(M, ctx)
    assert np.array_equal(d.matvec(x), ys[0])


@pytest.mark.parametrize("which", ["A", "B"])
def test_spmv_16bit_offsets_bit_identical(hg, ctx, ct64, ct48_unmatched, which):
    """col = base[group] + uint16 offset moves 10 instead of 12 bytes per entry; entry order and
    arithmetic are those of the 32-bit kernels, so the product is bit-identical whenever the launch
    shapes agree (the sliced 16-bit kernel splits slices across warps when the matrix has few rows)."""
    M = ct64[0] if which == "A" else ct48_unmatched[1]  # ragged rays (row per warp) / uniform pixel rows (sliced)
    x = np.random.default_rng(15).standard_normal(M.shape[1])
    assert hg.DeviceMatrix.from_any(M, ctx).spmv_index_bits == (32 if which == "A" else 16)  # defaults at this size
    try:
        hg.set_option("spmv_idx8", 1)  # byte offsets whatever the size (default: from 131 072 rows)
        d8 = hg.DeviceMatrix.from_any(M, ctx)
        assert d8.spmv_index_bits == (32 if which == "A" else 8)
        y8 = d8.matvec(x)
        hg.set_option("spmv_idx16", 2)  # 16-bit offsets for the row-per-warp kernel too (opt-in)
        hg.set_option("spmv_idx8", 0)
        d16 = hg.DeviceMatrix.from_any(M, ctx)
        assert d16.spmv_index_bits == 16 and d16.spmv_form == ("csr" if which == "A" else "sell32")
        y16 = d16.matvec(x)
        hg.set_option("spmv_idx16", 0)
        d32 = hg.DeviceMatrix.from_any(M, ctx)
        assert d32.spmv_index_bits == 32
        y32 = d32.matvec(x)
    finally:
        hg.set_option("spmv_idx16", 1)
        hg.set_option("spmv_idx8", -1)
    if which == "B":
        assert np.array_equal(y8, y16)  # 8- and 16-bit sliced kernels: same launch shape, same entry order
    if which == "A":
        assert np.array_equal(y16, y32)  # same kernel shape, same entry order: bit-identical
    assert _rel(y16, y32) < RTOL
    assert _rel(y16, M @ x) < RTOL


def test_spmv_16bit_offsets_fall_back_on_wide_groups(hg, ctx):
    """A group of entries spanning >= 65536 columns keeps the matrix on 32-bit indices."""
    rng = np.random.default_rng(16)
    rows, cols, per_row = 200, 300000, 64
    indptr = np.arange(rows + 1) * per_row
    indices = np.concatenate([np.sort(rng.choice(cols, per_row, replace=False)) for _ in range(rows)]).astype(np.int32)
    M = sp.csr_matrix((rng.standard_normal(indices.shape[0]), indices, indptr), shape=(rows, cols))
    d = hg.DeviceMatrix.from_any(M, ctx)
    assert d.spmv_index_bits == 32
    x = rng.standard_normal(cols)
    assert _rel(d.matvec(x), M @ x) < RTOL
    hg.set_option("spmv_mode", 1)  # the CSR kernel's 16-bit companion falls back the same way
    hg.set_option("spmv_idx16", 2)
    try:
        d1 = hg.DeviceMatrix.from_any(M, ctx)
        assert d1.spmv_form == "csr" and d1.spmv_index_bits == 32
        assert _rel(d1.matvec(x), M @ x) < RTOL
    finally:
        hg.set_option("spmv_mode", 0)
        hg.set_option("spmv_idx16", 1)


@pytest.mark.parametrize("n,k", [(40000, 24), (40000, 40), (40001, 41), (50050, 88), (38000, 89), (41003, 150),
                                 (39999, 208)])
def test_cgs_staged_stage_against_numpy(hg, ctx, n, k):
    """The one-pass CGS2 middle stage (csrc/cgs_staged.cu): `w1 = w0 - V h`, `d = V' w1` — every tile
    shape (k <= 40 / <= 88 / <= 208), n not a multiple of the tile (odd n: a row pair straddles the end),
    against NumPy and against the separate kernels, bit-identical reruns."""
    rng = np.random.default_rng(20 + k)
    V = np.asfortranarray(rng.standard_normal((n, k)))
    h = rng.standard_normal(k)
    w0 = rng.standard_normal(n)
    lib = ctx._lib
    outs = []
    for fused in (1, 1, 0):
        w1, d = np.zeros(n), np.zeros(k)
        hg._lib.check(lib.hg_cgs_mid(ctx._h, n, k, V.ctypes.data, n, h.ctypes.data, w0.ctypes.data, fused,
                                     w1.ctypes.data, d.ctypes.data))
        outs.append((w1, d))
    ref_w1 = w0 - V @ h
    ref_d = V.T @ ref_w1
    scale = np.linalg.norm(V, axis=0) * np.linalg.norm(ref_w1)
    for w1, d in outs:
        assert _rel(w1, ref_w1) < 1e-14
        assert np.max(np.abs(d - ref_d) / scale) < 1e-14
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1], equal_nan=True)
    with pytest.raises(hg._lib.HgError):  # outside its range the staged kernel refuses (callers fall back)
        hg._lib.check(lib.hg_cgs_mid(ctx._h, 1000, 8, V.ctypes.data, n, h.ctypes.data, w0.ctypes.data, 1,
                                     np.zeros(1000).ctypes.data, np.zeros(8).ctypes.data))


@pytest.mark.parametrize("n,k", [(1, 1), (63, 1), (1000, 5), (4097, 16), (20000, 17), (65536, 40), (65537, 41),
                                 (131072, 96), (131075, 97), (70001, 200), (300000, 208)])
def test_cgs2_whole_step_kernel_against_numpy(hg, ctx, n, k):
    """csrc/cgs2_step.cu — the persistent cooperative kernel that does the whole CGS2 step: every template
    shape (k <= 16 / 40 / 96 / 208), ragged n, more CTAs than tiles.  Against the same algorithm in NumPy
    (h to 1e-13 of ||w||, the new vector orthogonal to V to 1e-14) and bit-identical on a rerun."""
    rng = np.random.default_rng(n + k)
    V, _ = np.linalg.qr(rng.standard_normal((n, k)))
    V = np.asfortranarray(V)
    w0 = rng.standard_normal(n) + V @ rng.standard_normal(k) * 3.0
    lib = ctx._lib
    outs = []
    for rep in range(2):
        hcol, q = np.zeros(k + 1), np.zeros(n)
        hg._lib.check(lib.hg_cgs2_step(ctx._h, n, k, V.ctypes.data, n, w0.ctypes.data, hcol.ctypes.data, q.ctypes.data))
        outs.append((hcol, q))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1], equal_nan=True)
    hcol, q = outs[0]
    h1 = V.T @ w0
    w1 = w0 - V @ h1
    h2 = V.T @ w1
    v = w1 - V @ h2
    nw = np.linalg.norm(w0)
    assert np.max(np.abs(hcol[:k] - (h1 + h2))) < 1e-13 * nw
    assert abs(hcol[k] - np.linalg.norm(v)) < 1e-13 * nw
    if np.linalg.norm(v) > 1e-8 * nw:
        assert np.linalg.norm(q - v / np.linalg.norm(v)) < 1e-12
        assert np.max(np.abs(V.T @ q)) < 1e-13


def test_arnoldi_whole_step_kernel_matches_separate_kernels(hg, ctx):
    """The Arnoldi handle with the whole-step kernel (default for n <= 4000000) against the separate
    multi-dot / staged / update kernels (cgs_step_max_n = 0) on the 256^2 problem: H column-wise to 1e-9 over
    the well-determined columns, orthonormal basis, bit-identical reruns."""
    from hybrid_gmres_b200.ct import ct_backprojector, ct_projector, shepp_logan
    N, K, lam = 256, 60, 1e-2
    angles = np.arange(180) * 2.0
    dA = ct_projector(N, angles, None, "fan", ctx=ctx)
    dB = ct_backprojector(N, angles, None, "fan", ctx=ctx)
    b = dA.matvec(shepp_logan(N))
    res = {}
    try:
        for mode in (0, 4000000, 4000000):
            hg.set_option("cgs_step_max_n", mode)
            ar = hg.Arnoldi(dA, dB, "n", K)
            ar.set_rhs(b)
            ar.reset(lam)
            l0 = ctx.launch_count
            ar.steps(K)
            H, beta, k = ar.get()
            res.setdefault(mode, []).append((H, ctx.launch_count - l0))
            if mode:
                Q = np.column_stack([ar.q(j) for j in range(K + 1)])
            ar.close()
    finally:
        hg.set_option("cgs_step_max_n", 140000)
    (H0, l_sep), (H1, l_fused), (H2, _) = res[0][0], res[4000000][0], res[4000000][1]
    assert np.array_equal(H1, H2)
    assert l_fused <= 3 * K + 4 and l_sep > 3 * l_fused  # SpMV A, SpMV B, one CGS2 kernel per step
    for j in range(20):
        assert np.linalg.norm(H1[:j + 2, j] - H0[:j + 2, j]) <= 1e-9 * np.linalg.norm(H0[:j + 2, j]), j
    assert np.max(np.abs(Q.T @ Q - np.eye(K + 1))) < 1e-12


def test_malformed_sparse_inputs_are_rejected(hg, ctx):
    """hg_matrix_from_csr / _from_csc validate the caller's arrays on the device (pointer array monotone and
    ending at nnz, indices inside the matrix) and return HG_ERR_INVALID instead of reading or scattering out
    of bounds in every later product (the transposition of a CSC upload uses the indices as addresses)."""
    from hybrid_gmres_b200._lib import HgError
    indptr = np.array([0, 2, 3, 5], dtype=np.int64)
    indices = np.array([0, 2, 1, 0, 3], dtype=np.int32)
    data = np.arange(5, dtype=np.float64) + 1.0
    ok = hg.DeviceMatrix.from_csr(indptr, indices, data, (3, 4), ctx)
    assert np.allclose(ok.matvec(np.ones(4)), [3.0, 3.0, 9.0])
    ok.close()
    for bad_ptr, bad_idx in ((np.array([0, 3, 2, 5], dtype=np.int64), indices),            # decreasing pointer
                             (np.array([0, 2, 3, 4], dtype=np.int64), indices),            # does not end at nnz
                             (np.array([1, 2, 3, 5], dtype=np.int64), indices),            # does not start at 0
                             (indptr, np.array([0, 2, 1, 0, 4], dtype=np.int32)),          # column == cols
                             (indptr, np.array([0, -1, 1, 0, 3], dtype=np.int32))):        # negative column
        with pytest.raises(HgError) as e:
            hg.DeviceMatrix.from_csr(bad_ptr, bad_idx, data, (3, 4), ctx)
        assert e.value.status == 1 and "malformed" in str(e.value)
    # MATLAB-style CSC with 64-bit indices: a row index outside the matrix (also one that only differs above bit 31)
    jc = np.array([0, 2, 3, 5, 5], dtype=np.int64)
    for bad_ir in (np.array([0, 3, 1, 0, 2], dtype=np.int64), np.array([0, 2 ** 32 + 1, 1, 0, 2], dtype=np.int64)):
        with pytest.raises(HgError) as e:
            hg.DeviceMatrix.from_csc(jc, bad_ir, data, (3, 4), ctx)
        assert e.value.status == 1
    good = hg.DeviceMatrix.from_csc(jc, np.array([0, 2, 1, 0, 2], dtype=np.int64), data, (3, 4), ctx)
    assert np.allclose(good.matvec(np.ones(4)), [5.0, 3.0, 7.0])


@pytest.mark.parametrize("G", [2, 4, 8])
def test_spmv_row_group_form_matches_csr(hg, ctx, G):
    """csrc/spmv_group.cu — G adjacent rows interleaved per warp (the gather-bound projector A): same products
    as the row-per-warp CSR kernel up to summation order, ragged groups and a row count that is not a
    multiple of G, all epilogues; bit-identical reruns."""
    import scipy.sparse as sp
    from hybrid_gmres_b200.ct import ct_projector
    N = 256
    angles = np.arange(21) * (360.0 / 21)
    try:
        hg.set_option("spmv_group", 0)
        d0 = ct_projector(N, angles, 363, "fan", ctx=ctx)  # 7623 rows (odd), ~227 entries per row
        form0 = d0.spmv_form
        x = np.random.default_rng(5).standard_normal(d0.shape[1])
        y0 = d0.matvec(x)
        M = sp.csr_matrix(tuple(reversed(d0.download())), shape=d0.shape)
        hg.set_option("spmv_group", G)
        d1 = hg.DeviceMatrix.from_csr(M.indptr, M.indices, M.data, M.shape, ctx)
        if form0 == "csr":
            assert d1.spmv_form == "group"
            # default: the column stream is 16-bit per-lane differences (10 B per non-zero); G = 2 strides 16
            # entries of a ray per round, which can leave int16 and then keeps the 32-bit stream
            assert d1.spmv_index_bits == 16 or G == 2
        y1, y2 = d1.matvec(x), d1.matvec(x)
        assert np.array_equal(y1, y2)
        assert np.max(np.abs(y1 - y0)) <= 1e-13 * np.max(np.abs(y0))
        assert np.max(np.abs(y1 - M @ x)) <= 1e-13 * np.max(np.abs(y0))
        # several warps per group (small matrices; default by size): chunks of the rounds from column checkpoints
        ysplit = {}
        for S in (1, 2, 4):
            hg.set_option("spmv_group_split", S)
            ysplit[S] = d1.matvec(x)
            assert np.array_equal(d1.matvec(x), ysplit[S])
            assert np.max(np.abs(ysplit[S] - y0)) <= 1e-13 * np.max(np.abs(y0))
        # 32-bit column stream: same rounds, same sums as one warp per group -> identical bits
        hg.set_option("spmv_group16", 0)
        d2 = hg.DeviceMatrix.from_csr(M.indptr, M.indices, M.data, M.shape, ctx)
        if form0 == "csr":
            assert d2.spmv_form == "group" and d2.spmv_index_bits == 32
            assert np.array_equal(d2.matvec(x), ysplit[1]) or d1.spmv_index_bits == 32
        hg.set_option("spmv_group16", 1)
        hg.set_option("spmv_group_split", 0)
        # a matrix whose rows jump by more than int16 between rounds falls back to the 32-bit stream
        rng = np.random.default_rng(9)
        nr, per, stride = 2048, 256, 10000
        cols = (np.arange(per)[None, :] * stride + rng.integers(0, stride, size=(nr, per))).astype(np.int32)
        W = sp.csr_matrix((rng.standard_normal(nr * per), cols.ravel(), np.arange(nr + 1, dtype=np.int64) * per),
                          shape=(nr, per * stride))
        dW = hg.DeviceMatrix.from_csr(W.indptr, W.indices, W.data, W.shape, ctx)
        xw = rng.standard_normal(W.shape[1])
        assert dW.spmv_form == "group" and dW.spmv_index_bits == 32
        assert np.max(np.abs(dW.matvec(xw) - W @ xw)) <= 1e-13 * np.max(np.abs(W @ xw))
        # through a solver: shift epilogue, residual statistics
        import oracle
        from oracle import ct
        A, B, b, x_true = ct.make_ct_problem(128, 30, "fan", "pixel", noise=0.01)
        xs, es, rs, its = hg.hybrid_ba_gmres_rtp(A, B, b, x_true, 1e-6, 20, 1e-2, ctx=ctx, residual_mode=1, cache=False)
        xo, eo, ro, ito = oracle.hybrid_ba_gmres_rtp(A, B, b, x_true, 1e-6, 20, 1e-2)
        assert its == ito and np.max(np.abs(rs - ro) / ro) < 1e-8 and np.max(np.abs(es - eo) / eo) < 1e-8
    finally:
        hg.set_option("spmv_group16", 1)
        hg.set_option("spmv_group_split", 0)
        hg.set_option("spmv_group", -1)  # back to the default (G = 4 for matrices of >= 16 384 rows)
