"""Target for ncu (scripts, not product): steady-state TRG steps at chi = 32 on the Z2 tensor (the bench workload).
Warm-up outside the profiled range (cudaProfilerStart/Stop; run ncu with --profile-from-start off), then N steps
inside it -- the whole-step CUDA graph is replayed there, ncu profiles its kernel nodes one by one."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import grassmanntn_b200 as gtn
g = gtn.gauge2d
T = g.zcap(g.load_initial_tensor()).toblock()
for _ in range(2):
    T, _ = g.trg(T, 32)
for _ in range(int(os.environ.get("WARM", "30"))):     # the bench's step: the same site tensor every time
    g.trg(T, 32)
g.freeze(True)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    X, n = g.trg(T, 32)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("Tnorm", n, "step graph", g.STEP_GRAPH_STATS)
