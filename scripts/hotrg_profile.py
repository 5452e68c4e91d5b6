"""hotrg3dz at Zcut = 64 on the Z2 tensor (BASELINE.json configs[2]): wall time per step, CUPTI kernel timeline, host
profile (cProfile) and launch count -- where the step's time goes.  Run on the GPU box."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import grassmanntn_b200 as gtn
from grassmanntn_b200 import gauge2d as g, _ops
cut = int(os.environ.get("ZCUT", "64"))
T6 = g.load_initial_tensor()
print("input", T6.shape, T6.statistics)
for _ in range(3):
    g.hotrg3dz(T6, T6, cut)
torch.cuda.synchronize()
n0 = gtn.launch_count()
t0 = time.perf_counter()
for _ in range(5):
    out = g.hotrg3dz(T6, T6, cut)
torch.cuda.synchronize()
print("ms/step %.2f" % ((time.perf_counter() - t0) / 5 * 1e3), "launches/step", (gtn.launch_count() - n0) / 5, "paths", _ops.SVD_PATH_STATS)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        g.hotrg3dz(T6, T6, cut)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
pr = cProfile.Profile(); pr.enable()
for _ in range(3):
    g.hotrg3dz(T6, T6, cut)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(40)
