import os, sys, json
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "oracle"))
import numpy as np, torch
import grassmanntn_b200 as gtn
from grassmanntn_b200 import _engine as E, _ops
g = gtn.gauge2d
T = g.zcap(g.load_initial_tensor()).toblock()
for _ in range(2):
    T, _ = g.trg(T, 32)
X = T
for i in range(3):
    X = (g.atrg2dx if i % 2 else g.atrg2dy)(X, X, 32)[0]
torch.cuda.synchronize()
_ops.SVD_PATH_STATS.update(truncated=0, truncated_rejected=0, full=0)
E.PROF.start()
X = T
import time; t0 = time.perf_counter()
for i in range(4):
    X = (g.atrg2dx if i % 2 else g.atrg2dy)(X, X, 32)[0]
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 4 * 1e3
pr = E.PROF.stop()
print('ms/step', dt, _ops.SVD_PATH_STATS, 'sweeps', E.batched_svd.last_sweeps)
for k, v in sorted(pr.items(), key=lambda kv: -kv[1]['ms']):
    print('  %-18s %8.2f ms/step  %6.1f launches/step' % (k, v['ms'] / 4, v['launches'] / 4))
