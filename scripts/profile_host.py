import os, sys, cProfile, pstats
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "oracle"))
import numpy as np, torch
import grassmanntn_b200 as gtn
g = gtn.gauge2d
T = g.zcap(g.load_initial_tensor()).toblock()
for _ in range(2):
    T, _ = g.trg(T, 32)
X = T
for _ in range(3):
    X, _ = g.trg(X, 32)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
X = T
for _ in range(10):
    X, n = g.trg(X, 32)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(45)
