"""CPU: the OpenMP C restatement (bench.py's all-cores CPU baseline) against the NumPy oracle."""
import subprocess
import os

import numpy as np
import pytest

import oracle
from oracle import cport, ct

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def built():
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    assert cport.available()


@pytest.mark.parametrize("b_kind,geometry", [("matched", "parallel"), ("pixel", "fan")])
def test_c_port_matches_numpy_oracle(b_kind, geometry):
    A, B, b, x_true = ct.make_ct_problem(24, 36, geometry, b_kind)
    # single-pass MGS loses orthogonality on this small problem after ~20 steps, where two
    # summation orders part ways transiently: strict over the first 15.
    x, err, res, it = cport.hybrid_ba_gmres_rtp(A, B, b, x_true, 1e-6, 15, 1e-2)
    xo, erro, reso, ito = oracle.hybrid_ba_gmres_rtp(A, B, b, x_true, 1e-6, 15, 1e-2)
    assert it == ito
    assert np.max(np.abs(res - reso) / reso) < 1e-9
    assert np.max(np.abs(err - erro) / erro) < 1e-9
    assert np.linalg.norm(x - xo) / np.linalg.norm(xo) < 1e-9
    assert cport.num_threads() >= 1


def test_c_port_stops_like_the_reference():
    A, B, b, x_true = ct.make_ct_problem(24, 36, "parallel", "matched")
    x, err, res, it = cport.hybrid_ba_gmres_rtp(A, B, b, x_true, 0.1, 30, 1e-2)
    xo, erro, reso, ito = oracle.hybrid_ba_gmres_rtp(A, B, b, x_true, 0.1, 30, 1e-2)
    assert it == ito and res[-1] <= 0.1


@pytest.mark.parametrize("kind", ["ab", "ba"])
@pytest.mark.parametrize("orth", ["mgs", "cgs2"])
def test_c_port_rtp_full_outputs_match_numpy_oracle(kind, orth):
    """AB and BA, MGS and CGS2: histories, H column-wise, beta and every iterate vs oracle/solvers.py
    (which is pinned to the executed reference source, tests/test_reference_golden.py)."""
    A, B, b, x_true = ct.make_ct_problem(24, 36, "fan", "pixel")
    K = 15
    exc, exo = {}, {}
    x, err, res, it = cport.hybrid_rtp(kind, A, B, b, x_true, 1e-6, K, 1e-2, orth, exc, want_X=True)
    f = oracle.hybrid_ab_gmres_rtp if kind == "ab" else oracle.hybrid_ba_gmres_rtp
    xo, erro, reso, ito = f(A, B, b, x_true, 1e-6, K, 1e-2, orth=orth, extras=exo)
    assert it == ito
    assert np.max(np.abs(res - reso) / reso) < 1e-9 and np.max(np.abs(err - erro) / erro) < 1e-9
    assert abs(exc["beta"] - exo["beta"]) < 1e-14 * exo["beta"]
    for j in range(it):
        assert np.linalg.norm(exc["H"][: j + 2, j] - exo["H"][: j + 2, j]) / np.linalg.norm(exo["H"][: j + 2, j]) < 1e-9
        assert np.linalg.norm(exc["X"][:, j] - exo["X"][:, j]) / np.linalg.norm(exo["X"][:, j]) < 1e-9
    assert len(exc["t_iter"]) == it and np.all(np.diff(exc["t_iter"]) >= 0)


def test_c_port_arnoldi_only_and_breakdown():
    A, B, b, x_true = ct.make_ct_problem(24, 36, "fan", "pixel")
    ex1, ex2 = {}, {}
    _, _, _, k = cport.hybrid_rtp("ba", A, B, b, x_true, 0.0, 12, 1e-2, "cgs2", ex1, solve=False)
    cport.hybrid_rtp("ba", A, B, b, x_true, 0.0, 12, 1e-2, "cgs2", ex2)
    assert k == 12 and np.allclose(ex1["H"], ex2["H"], rtol=1e-10, atol=1e-13)  # OpenMP reductions: order may vary
    # `== 0` breakdown at k = 1 (hybrid_ab_gmres_rtp.m:25,41-43): x unassigned, zero history entry
    import scipy.sparse as sp
    n = 6
    I, e1 = sp.identity(n, format="csr"), np.eye(n)[:, 0].copy()
    x, err, res, it = cport.hybrid_ab_gmres_rtp(I, I, e1, np.ones(n), 1e-6, 4, 1e-2)
    assert x is None and it == 1 and res[0] == 0.0 and err[0] == 0.0
    x, err, res, it = cport.hybrid_ba_gmres_rtp(I, I, e1, np.ones(n), 1e-6, 4, 1e-2)
    assert it == 1 and np.array_equal(x, np.zeros(n)) and res[0] == 0.0


def test_c_port_thread_override():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm must still use every core."""
    env = dict(os.environ, OMP_NUM_THREADS="1", PYTHONPATH=ROOT)
    out = subprocess.run(["python", "-c", "from oracle import cport; import os; a = cport.num_threads(); "
                          "b = cport.use_all_cores(); print(a, b, len(os.sched_getaffinity(0)))"],
                         env=env, capture_output=True, text=True, check=True).stdout.split()
    assert out[0] == "1" and out[1] == out[2]
