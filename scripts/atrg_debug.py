import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grassmanntn_b200 as gtn
from grassmanntn_b200 import gauge2d as g, _engine as E, _ops
T = g.zcap(g.load_initial_tensor()).toblock()
for _ in range(2):
    T, _ = g.trg(T, 32)
X = T
for _ in range(3):
    X = g.atrg2dy(X, X, 32)[0]
torch.cuda.synchronize()
E.DEBUG_TRUNC = True
for i in range(2):
    print("==== step", i, flush=True)
    b = dict(_ops.SVD_PATH_STATS)
    t0 = time.perf_counter()
    X = g.atrg2dy(X, X, 32)[0]
    torch.cuda.synchronize()
    print("ms", (time.perf_counter() - t0) * 1e3, {k: _ops.SVD_PATH_STATS[k] - b[k] for k in b}, flush=True)
