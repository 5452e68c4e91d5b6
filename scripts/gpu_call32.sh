#!/bin/bash
mkdir -p gpurun_out
timeout 600 python - > gpurun_out/c32_atrg_time.log 2>&1 <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
import grassmanntn_b200 as gtn
from grassmanntn_b200 import _ops, _engine as E
g = gtn.gauge2d
chi = 128
T = g.zcap(g.load_initial_tensor()).toblock()
while tuple(T.effective_shape) != (chi,) * 4:
    T, _ = g.trg(T, chi)
X = T
for i in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    X, tn = (g.atrg2dy if i % 2 == 0 else g.atrg2dx)(X, X, chi)[:2]
    torch.cuda.synchronize()
    print("atrg step", i, "%.1f ms" % ((time.perf_counter() - t0) * 1e3), "Tnorm %.15g" % tn, _ops.SVD_PATH_STATS, E.RANK_CHECK_STATS, flush=True)
PY
echo "rc=$?"; grep "atrg step" gpurun_out/c32_atrg_time.log
timeout 600 python scripts/atrg_launch_probe.py > gpurun_out/c32_atrg_probe.log 2>&1; grep -E "^step|total" gpurun_out/c32_atrg_probe.log | head -24
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/c32_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c32_pytest.log | cut -c1-250
