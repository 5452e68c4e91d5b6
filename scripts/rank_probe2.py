"""Which decomposition of the chi = 64 TRG chain on the Z2 tensor breaks the reference's rank rule s_i / s_0 > 1e-14:
every call of the full Jacobi SVD prints the singular values around the numerical rank of each sector matrix next to
torch.linalg.svdvals (cuSOLVER) on the same device matrices, and the path statistics per step."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import grassmanntn_b200 as gtn
from grassmanntn_b200 import _engine as E, _ops
g = gtn.gauge2d
orig = _ops.batched_svd

def hooked(mats):
    refs = [torch.linalg.svdvals(m).cpu().numpy() for m in mats]
    usv = orig(mats)
    for b, ((U, s, Vh), r, m) in enumerate(zip(usv, refs, mats)):
        nj, nr = int(np.sum(s / (s[0] + 1e-14) > 1e-14)), int(np.sum(r / (r[0] + 1e-14) > 1e-14))
        lo = max(min(nj, nr) - 2, 0)
        print("  full SVD mat", b, tuple(m.shape), "rank jacobi", nj, "cusolver", nr, "sweeps", E.batched_svd.last_sweeps)
        print("    jacobi  ", np.array2string(s[lo:lo + 10] / s[0], precision=2))
        print("    cusolver", np.array2string(r[lo:lo + 10] / r[0], precision=2), flush=True)
    return usv

_ops.batched_svd = hooked
E.DEBUG_TRUNC = bool(int(os.environ.get("DBG", "0")))
T = g.zcap(g.load_initial_tensor().toblock())
for i in range(int(os.environ.get("STEPS", "3"))):
    before = dict(_ops.SVD_PATH_STATS)
    T, Tn = g.trg(T, 64)
    print("step", i + 1, "shape", T.effective_shape, "Tnorm", Tn, "paths", {k: _ops.SVD_PATH_STATS[k] - before[k] for k in before}, flush=True)
