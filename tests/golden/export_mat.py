"""Write the golden fixtures as MATLAB v5 .mat files for oracle/replay.m (not committed).

    python tests/golden/export_mat.py [outdir]

A and B are stored the way the npz fixtures hold them (``A_jc`` / ``A_ir`` / ``A_pr`` / ``A_shape`` —
the arrays a MEX gateway sees — or ``A_dense``); replay.m rebuilds the sparse matrices from them."""
import glob
import os
import sys

import numpy as np
import scipy.io as sio

HERE = os.path.dirname(os.path.abspath(__file__))


def export(outdir=HERE, names=None):
    written = []
    for f in sorted(glob.glob(os.path.join(HERE, "*.npz"))):
        stem = os.path.basename(f)[:-4]
        if stem.startswith("ref_") or (names is not None and stem not in names):
            continue
        out = os.path.join(outdir, stem + ".mat")
        sio.savemat(out, dict(np.load(f)), do_compression=True)
        written.append(out)
    return written


if __name__ == "__main__":
    for w in export(sys.argv[1] if len(sys.argv) > 1 else HERE):
        print("wrote", w)
