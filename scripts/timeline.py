"""Kernel timeline of steady-state TRG steps (torch.profiler / CUPTI): true GPU busy time per kernel,
including the kernels inside the replayed CUDA graph.  Run on the GPU box."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import grassmanntn_b200 as gtn
from grassmanntn_b200 import gauge2d as g

ap = argparse.ArgumentParser()
ap.add_argument("--chi", type=int, default=32)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--method", default="trg")
args = ap.parse_args()
T = g.zcap(g.load_initial_tensor()).toblock()
for _ in range(2):
    T, _ = g.trg(T, args.chi)
def step(X):
    if args.method == "trg":
        return g.trg(X, args.chi)[0]
    return g.atrg2dy(X, X, args.chi)[0]
for _ in range(30):      # reach the steady state on the bench's step (same tensor every time): counts settled, graph recorded
    step(T)
g.freeze(True)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(args.steps):
        X = step(T)
    torch.cuda.synchronize()
print("steps", args.steps, "step graph", g.STEP_GRAPH_STATS)
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
