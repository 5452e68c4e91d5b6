"""Golden for the 'HOTRG chi=64' configuration: flavour coarse-graining
hotrg3dz(T, T, 64) of the Z2 initial tensor with the REAL reference (oracle/ref_harness.py), as in
example.py:144-154 with --Nf 2 --Dcutz 64.  At Zcut = 64 = 8*8 nothing is truncated, so the result does
not depend on how degenerate multiplets are cut and is reproducible to rounding.
Records Tnorm, trace error, F = logZ(zcap(T)) + log Tnorm and the output shape."""
import os, sys, time
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import ref_harness
from threadpoolctl import threadpool_limits

gtn = ref_harness.load_reference()
z = np.load(os.path.join(HERE, "z2_initial_tensor.npz"))
T0 = gtn.dense(z["data"], statistics=tuple(int(s) for s in z["statistics"]))
cut = int(sys.argv[1]) if len(sys.argv) > 1 else 64
t0 = time.time()
with threadpool_limits(limits=1):
    T, Tn, err = gtn.gauge2d.hotrg3dz(T0.copy(), T0.copy(), cut, iternum=0, error_test=True)
    print("hotrg3dz done %.1f s" % (time.time() - t0), Tn, err, T.shape, flush=True)
    Tc = gtn.gauge2d.zcap(T)
    F = gtn.gauge2d.logZ(Tc.copy(), "anti-periodic") + np.log(Tn)
rec = np.array([Tn, err, F.real, F.imag] + list(T.shape), dtype=float)
np.savez_compressed(os.path.join(HERE, "z2_hotrg%d.npz" % cut), rec=rec, Tc_norm=float(Tc.norm))
print("saved", rec, "%.1f s" % (time.time() - t0), flush=True)
