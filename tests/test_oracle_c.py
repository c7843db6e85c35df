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
