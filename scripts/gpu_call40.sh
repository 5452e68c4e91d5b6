#!/bin/bash
# round 2, call 40 (4 GPUs): sharded tests at the final state; ATRG chi = 192 sharded over 4 GPUs (l = 192 subspaces: the packed
# Cholesky on the global scratch inside a real run)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x > gpurun_out/c40_sharded.log 2>&1; echo "sharded test rc=$?"; tail -1 gpurun_out/c40_sharded.log
GTN_DEBUG_TRUNC=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29591 scripts/atrg_sharded.py --chi 192 --steps 4 --out gpurun_out/r2h_atrg_sharded_chi192_n4.json > gpurun_out/c40_atrg192.log 2>&1; echo "atrg192 rc=$?"; grep -E "^\{\"step|Error|trunc sharded" gpurun_out/c40_atrg192.log | cut -c1-240 | tail -30
