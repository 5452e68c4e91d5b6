"""TRG / ATRG steady-state step time and truncated-SVD iteration counts (engine defaults or env overrides)"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
import grassmanntn_b200 as gtn
from grassmanntn_b200 import _engine as E, _ops
g = gtn.gauge2d
T0 = g.zcap(g.load_initial_tensor()).toblock()
for _ in range(2):
    T0, _ = g.trg(T0, 32)
for algo, nsteps in (("trg", 60), ("atrg", 80)):
    X = T0
    ts = []
    for i in range(nsteps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        if algo == "trg":
            X, n = g.trg(X, 32)[:2]
        else:
            X, n = (g.atrg2dx if i % 2 == 0 else g.atrg2dy)(X, X, 32)[:2]
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    tail = ts[-6:]
    print(algo, "last6 mean %.3f ms" % (sum(tail) / len(tail)), "all:", " ".join("%.2f" % t for t in ts))
    print("   hints", {str(k[1]): v for k, v in E._trunc_iters_hint.items()}, "paths", _ops.SVD_PATH_STATS,
          "graph", g.STEP_GRAPH_STATS, "spec", g.SPEC_STATS, "probe", {str(k[1]): v["every"] for k, v in E._trunc_probe.items()}, "Tnorm %.13g" % n, flush=True)
