"""the .mat v5 wire format of the fixtures and oracle/replay.m.

The fixtures are exported with scipy.io.savemat, read back, and — in the build container, where the
reference sources exist — oracle/replay.m itself is executed by the mini interpreter (oracle/mlab.py):
it loads the .mat file, rebuilds the MATLAB sparse matrices from Jc/Ir/Pr, calls the UNTOUCHED reference
solvers and reports their distance to the stored oracle outputs.  (Under real MATLAB / Octave the same
script pins the oracle to MathWorks' own built-ins; that run is not possible here.)"""
import os
import re
import sys

import numpy as np
import pytest
import scipy.io as sio

from tests.golden_util import GOLDEN_DIR, load

sys.path.insert(0, GOLDEN_DIR)


def test_mat_round_trip(tmp_path):
    import export_mat
    written = export_mat.export(str(tmp_path), names=["ct16_perturbed", "deriv2_n32"])
    assert len(written) == 2
    for path in written:
        name = os.path.basename(path)[:-4]
        g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
        m = sio.loadmat(path)
        for key, val in g.items():
            got = np.asarray(m[key])
            assert got.size == np.asarray(val).size, key
            assert np.array_equal(got.reshape(-1), np.asarray(val).reshape(-1)), key
            if key.endswith(("_jc", "_ir")):
                assert got.dtype == np.int64  # what mwIndex arrays carry
        if "A_jc" in g:  # the MATLAB-sparse view of the same arrays gives back the matrix
            A, B, _ = load(name)
            jc, ir, pr = (m[k].reshape(-1) for k in ("A_jc", "A_ir", "A_pr"))
            import scipy.sparse as sp
            A2 = sp.csc_matrix((pr, ir, jc), shape=tuple(int(v) for v in m["A_shape"].reshape(-1)))
            assert (abs(A - A2)).nnz == 0


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference sources only exist in the build container")
def test_replay_script_runs_the_reference_on_the_mat_fixture(tmp_path):
    import export_mat
    from oracle import mlab
    (path,) = export_mat.export(str(tmp_path), names=["ct16_perturbed"])
    repo = os.path.dirname(GOLDEN_DIR.rstrip("/")).rsplit("/tests", 1)[0]
    ip = mlab.Interp([os.path.join(repo, "oracle"), "/root/reference"])
    ip.call("replay", [path], 0)
    text = "".join(ip.out)
    rows = dict(re.findall(r"^(\w+): iters (\d+ vs \d+)", text, flags=re.M))
    assert set(rows) == {"hybrid_ab_gmres_rtp", "hybrid_ba_gmres_rtp", "hybrid_lsqr_solver", "hybrid_lsmr_solver",
                         "lsqr_solver", "lsmr_solver"}
    assert all(v.split(" vs ")[0] == v.split(" vs ")[1] for v in rows.values())  # same stopping iteration
    for fn in ("hybrid_ab_gmres_rtp", "hybrid_ba_gmres_rtp"):
        xd, rd = re.search(fn + r": .*\|x-x_o\|/\|x_o\| = (\S+), max residual-history diff = (\S+)", text).groups()
        assert float(xd) < 1e-10 and float(rd) < 1e-10
    for t in ("ab", "ba"):
        assert float(re.search(rf"gcv_function {t}: max rel diff (\S+)", text).group(1)) < 1e-9
