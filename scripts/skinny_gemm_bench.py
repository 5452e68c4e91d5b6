"""Panel products of the subspace iteration: nb problems of (l x p) . (p x q), both tile configurations."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from grassmanntn_b200 import _engine as E
dev = torch.device("cuda")
for (l, p, q, nb) in ((48, 512, 512, 4), (64, 2048, 2048, 4), (128, 8192, 8192, 4), (64, 2048, 2048, 2)):
    a = torch.randn(nb * l * p, dtype=torch.complex128, device=dev)
    b = torch.randn(nb * p * q, dtype=torch.complex128, device=dev)
    c = torch.empty(nb * l * q, dtype=torch.complex128, device=dev)
    for smax in (48, 256):
        E.SKINNY_MAX = smax
        groups = [dict(a_off=i * l * p, b_off=i * p * q, c_off=i * l * q, lda=p, ldb=q, ldc=q, m=l, n=q, k=p) for i in range(nb)]
        plan = E.GemmPlan(groups, torch.complex128)
        for _ in range(3):
            plan.run(a, b, c)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            plan.run(a, b, c)
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 10
        print("l=%d p=%d q=%d nb=%d config=%d tiles=%d: %.3f ms  %.1f TFLOP/s" % (l, p, q, nb, plan.config, plan.tiles, ms, plan.flops / ms / 1e9))
