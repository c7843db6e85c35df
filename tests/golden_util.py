"""Loading of the committed golden fixtures (tests/golden/*.npz, made by make_golden.py)."""
import os

import numpy as np
import scipy.sparse as sp

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ["deriv2_n32", "ct16_perturbed", "ct20_fan_pixel"]


def load(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))

    def mat(prefix):
        if prefix + "_dense" in g:
            return g[prefix + "_dense"]
        shape = tuple(int(v) for v in g[prefix + "_shape"])
        return sp.csc_matrix((g[prefix + "_pr"], g[prefix + "_ir"], g[prefix + "_jc"]), shape=shape)

    return mat("A"), mat("B"), g


def strict_iters(name):
    """Iterations over which 1e-8 parity is meaningful: on deriv2 n=32 the Arnoldi / GKB
    processes break down numerically after ~5 steps."""
    return 4 if name == "deriv2_n32" else 8
