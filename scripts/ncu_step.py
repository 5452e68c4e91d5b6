"""Target for ncu (scripts, not product): steady-state TRG steps at chi = $CHI (default 32) on the Z2 tensor (the bench
workload: the first tensor of the chain with all legs at chi, the same tensor every step).  Warm-up outside the
profiled range (cudaProfilerStart/Stop; run ncu with --profile-from-start off), then N steps inside it -- at
chi <= 64 the whole-step CUDA graph is replayed there (ncu profiles its kernel nodes one by one), at chi = 128 the
step is enqueued eagerly behind the replayed truncated-SVD schedule."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import grassmanntn_b200 as gtn
g = gtn.gauge2d
chi = int(os.environ.get("CHI", "32"))
T = g.zcap(g.load_initial_tensor()).toblock()
if chi == 32:
    for _ in range(2):
        T, _ = g.trg(T, 32)
else:
    while tuple(T.effective_shape) != (chi,) * 4:
        T, _ = g.trg(T, chi)
for _ in range(int(os.environ.get("WARM", "30"))):     # the bench's step: the same site tensor every time
    g.trg(T, chi)
g.freeze(True)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    X, n = g.trg(T, chi)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("Tnorm", n, "step graph", g.STEP_GRAPH_STATS)
