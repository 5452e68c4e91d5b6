"""per-step wall time of TRG / ATRG chains with the whole-step graph off / on"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
import grassmanntn_b200 as gtn
from grassmanntn_b200 import _engine as E, _ops
g = gtn.gauge2d
T0 = g.zcap(g.load_initial_tensor()).toblock()
for _ in range(2):
    T0, _ = g.trg(T0, 32)
for algo, nsteps in (("trg", 14), ("atrg", 28)):
    for graph in (False, True):
        g.STEP_GRAPH = graph
        E._trunc_iters_hint.clear(); E._trunc_rate.clear(); E._trunc_fail.clear()
        g._step_graphs.clear(); g._steady.clear()
        for k in list(g.STEP_GRAPH_STATS):
            g.STEP_GRAPH_STATS[k] = 0
        X = T0
        ts = []
        for i in range(nsteps):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            if algo == "trg":
                X, n = g.trg(X, 32)[:2]
            else:
                X, n = (g.atrg2dx if i % 2 == 0 else g.atrg2dy)(X, X, 32)[:2]
            torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
        print(algo, "graph", graph, "ms/step:", " ".join("%.2f" % t for t in ts))
        print("   stats", g.STEP_GRAPH_STATS, "spec", g.SPEC_STATS, "Tnorm %.13g" % n, flush=True)
