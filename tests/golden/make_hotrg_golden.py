"""Golden values for hotrg3dz (reference gauge2d.py:1891) on RANDOM Grassmann-even 6-leg tensors
(non-degenerate spectra, so the truncation is unambiguous -- the Z2 tensor has exact multiplets at
every small cut, see tests/test_z2_golden.py).  Run in the build container only."""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import ref_harness
from threadpoolctl import threadpool_limits
gtn = ref_harness.load_reference()
out = {}
with threadpool_limits(limits=1):
    for seed, shape, cut in ((5, (4, 4, 4, 4, 2, 2), 8), (6, (4, 4, 4, 4, 3, 3), 6)):
        np.random.seed(seed)
        T1 = gtn.random(shape, (1, 1, -1, -1, 0, 0), dtype=complex)
        T2 = gtn.random(shape, (1, 1, -1, -1, 0, 0), dtype=complex)
        T, Tn, err = gtn.gauge2d.hotrg3dz(T1.copy(), T2.copy(), cut, error_test=True)
        F = gtn.gauge2d.logZ(gtn.gauge2d.zcap(T), 'anti-periodic')
        tag = "s%d" % seed
        out[tag + "_T1"], out[tag + "_T2"] = T1.data, T2.data
        out[tag + "_res"] = np.array([Tn, err, F.real, F.imag, cut] + list(T.shape), dtype=float)
        print(tag, Tn, err, F, T.shape, flush=True)
np.savez_compressed(os.path.join(HERE, "hotrg_random.npz"), **out)
