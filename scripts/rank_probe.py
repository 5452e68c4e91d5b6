"""Singular values around the numerical rank of the sector matrices of the second TRG step at chi = 64 on the Z2 tensor
(exact rank 16 per sector; numpy's noise floor is 2e-16 ... 2e-15 s_0), from the full Jacobi SVD kernels, next to
torch.linalg.svdvals on the same device matrices.  Diagnoses the rank rule s_i / s_0 > 1e-14 (reference __init__.py:3939)."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import grassmanntn_b200 as gtn
from grassmanntn_b200 import _engine as E, _ops
g = gtn.gauge2d
T = g.zcap(g.load_initial_tensor()).toblock()
T, _ = g.trg(T, 64)
print("after step 1:", T.effective_shape)
T1 = gtn.einsum("ijkl->jkli", T)
ctx = _ops._decompose_prepare(T1._bt, 2, "svd")
mats = ctx["mats"]
for tol in (E.JACOBI_TOL,):
    usv = E.batched_svd([m.clone() for m in mats])
    print("sweeps", E.batched_svd.last_sweeps)
    for (U, s, Vh), M in zip(usv, mats):
        ref = torch.linalg.svdvals(M).cpu().numpy()
        print("jacobi s[12:26]/s0", np.array2string(s[12:26] / s[0], precision=2))
        print("cusolver s[12:26]/s0", np.array2string(ref[12:26] / ref[0], precision=2))
        print("count > 1e-14:", int(np.sum(s / s[0] > 1e-14)), "ref", int(np.sum(ref / ref[0] > 1e-14)),
              "max null", float(s[16:].max() / s[0]))
