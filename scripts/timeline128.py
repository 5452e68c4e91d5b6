import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import grassmanntn_b200 as gtn
from grassmanntn_b200 import gauge2d as g
chi = int(sys.argv[1]) if len(sys.argv) > 1 else 128
T = g.zcap(g.load_initial_tensor()).toblock()
for _ in range(4):
    T, _ = g.trg(T, chi)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        T, _ = g.trg(T, chi)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
