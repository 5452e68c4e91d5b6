"""find what an ATRG step with a >50 ms spike was doing (engine debug lines of that step only)"""
import io, os, sys, time, contextlib
os.environ["GTN_DEBUG_TRUNC"] = "1"
sys.path.insert(0, os.getcwd())
import torch
import grassmanntn_b200 as gtn
from grassmanntn_b200 import _engine as E, _ops
g = gtn.gauge2d
T0 = g.zcap(g.load_initial_tensor()).toblock()
for _ in range(2):
    T0, _ = g.trg(T0, 32)
X = T0
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 50):
    buf = io.StringIO()
    p0 = dict(_ops.SVD_PATH_STATS); g0 = dict(g.STEP_GRAPH_STATS)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with contextlib.redirect_stdout(buf):
        X, n = (g.atrg2dx if i % 2 == 0 else g.atrg2dy)(X, X, 32)[:2]
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    if dt > 40 and i > 8:
        print("=== step", i, "%.1f ms" % dt, {k: _ops.SVD_PATH_STATS[k] - p0[k] for k in p0},
              {k: g.STEP_GRAPH_STATS[k] - g0[k] for k in g0 if isinstance(g0[k], int)})
        for line in buf.getvalue().splitlines():
            print("   ", line[:260])
