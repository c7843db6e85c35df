import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def hg():
    import hybrid_gmres_b200

    return hybrid_gmres_b200


@pytest.fixture(scope="session")
def ctx(hg):
    return hg.default_context()


@pytest.fixture(scope="session")
def ct64():
    """BASELINE config 1 shape: 64x64 parallel beam, 180 views, matched B, 1% noise."""
    from oracle import ct

    A, B, b, x_true = ct.make_ct_problem(64, 180, "parallel", "matched", noise=0.01)
    return A, B, b, x_true


@pytest.fixture(scope="session")
def ct48_unmatched():
    """Small unmatched problem: 48x48 fan beam, 90 views, pixel-driven B."""
    from oracle import ct

    return ct.make_ct_problem(48, 90, "fan", "pixel", noise=0.01)
