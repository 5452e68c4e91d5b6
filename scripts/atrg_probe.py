"""per-step wall time, speculation outcome and iteration hints of an ATRG chain (spec off / on)"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
import grassmanntn_b200 as gtn
from grassmanntn_b200 import _engine as E, _ops
g = gtn.gauge2d
T0 = g.zcap(g.load_initial_tensor()).toblock()
for _ in range(2):
    T0, _ = g.trg(T0, 32)
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 14
for spec in (False, True):
    g.SPECULATE = spec
    E._trunc_iters_hint.clear(); E._trunc_rate.clear(); E._trunc_fail.clear()
    X = T0
    print("=== speculate", spec)
    for i in range(nsteps):
        s0 = dict(g.SPEC_STATS); p0 = dict(_ops.SVD_PATH_STATS)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        X, n = (g.atrg2dx if i % 2 == 0 else g.atrg2dy)(X, X, 32)[:2]
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
        hints = {k[1][1:] if isinstance(k[1], tuple) and k[1][0] == "atrg" else "other": v for k, v in E._trunc_iters_hint.items()}
        print("step %2d %7.2f ms  spec+%d fail+%d  paths %s  hints %s  Tnorm %.12g" % (
            i, dt, g.SPEC_STATS["speculated"] - s0["speculated"], g.SPEC_STATS["failed"] - s0["failed"],
            {k: _ops.SVD_PATH_STATS[k] - p0[k] for k in p0}, hints, n), flush=True)
