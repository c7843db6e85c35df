"""Write the golden fixtures as MATLAB v5 .mat files for oracle/replay.m (not committed)."""
import glob
import os

import numpy as np
import scipy.io as sio

HERE = os.path.dirname(os.path.abspath(__file__))
for f in glob.glob(os.path.join(HERE, "*.npz")):
    sio.savemat(f[:-4] + ".mat", dict(np.load(f)), do_compression=True)
    print("wrote", f[:-4] + ".mat")
