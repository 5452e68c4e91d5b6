"""Phase breakdown (clock64) of the small whitening / pre-rotation kernels on a steady-state TRG step."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grassmanntn_b200 as gtn
from grassmanntn_b200 import gauge2d as g, _engine as E
from grassmanntn_b200._cabi import lib
chi = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = g.zcap(g.load_initial_tensor()).toblock()
for _ in range(5):
    T, _ = g.trg(T, chi)
torch.cuda.synchronize()
buf = (C.c_longlong * 8)()
lib.gtn_debug_phase_clocks(buf)
v = list(buf)
print("chol_whiten cycles: factor %d  inverse %d  store %d" % (v[1] - v[0], v[2] - v[1], v[3] - v[2]))
print("gram_rotate cycles: factor %d  jacobi %d  store %d" % (v[5] - v[4], v[6] - v[5], v[7] - v[6]))
for plan in E._trunc_plans.values():
    print("L", plan.L_, "rot sweeps", plan.rot_sweeps.cpu().tolist(), "jacobi sweeps", E.batched_svd.last_sweeps)

import ctypes
lib._handle  # noqa
f = getattr(lib, "gtn_debug_clocks16", None)
if f is not None:
    b16 = (C.c_longlong * 16)()
    f.argtypes = [C.c_void_p]; f.restype = C.c_int
    f(b16)
    w = list(b16)
    print("pivot step k=5: search %d rsqrt %d publish %d barrier %d update %d diag %d store %d" % tuple(w[i + 1] - w[i] for i in range(7)))
