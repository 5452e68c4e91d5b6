#!/bin/bash
# round 2, call 8 (2 GPUs): whole GPU suite incl. the sharded tests, rank-rule probe with the rank certificate,
# sharded ATRG at chi = 128 (with single-GPU check) with the truncated-SVD trace
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/c8_pytest.log 2>&1; echo "pytest rc=$?"; tail -22 gpurun_out/c8_pytest.log | cut -c1-220
STEPS=4 timeout 300 python scripts/rank_probe2.py > gpurun_out/c8_rank_probe2.log 2>&1; echo "probe rc=$?"; grep -E "^step|rank jacobi" gpurun_out/c8_rank_probe2.log | cut -c1-200
GTN_DEBUG_TRUNC=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 scripts/atrg_sharded.py --chi 128 --steps 4 --check --out gpurun_out/r2_atrg_sharded_chi128_n2.json > gpurun_out/c8_atrg128.log 2>&1; echo "atrg128 rc=$?"; grep -E "^\{|trunc sharded|Error" gpurun_out/c8_atrg128.log | cut -c1-400 | tail -40
