#!/bin/bash
# round 2, call 24 (2 GPUs): whole GPU suite, sharded ATRG chi = 128 with the shifted-CholQR robust mode (with single-GPU check)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/c24_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/c24_pytest.log | cut -c1-300
GTN_DEBUG_TRUNC=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 scripts/atrg_sharded.py --chi 128 --steps 6 --check --out gpurun_out/r2f_atrg_sharded_chi128_n2.json > gpurun_out/c24_atrg128.log 2>&1; echo "atrg128 rc=$?"; grep -E "^\{\"step|trunc sharded.*robust True|Error" gpurun_out/c24_atrg128.log | cut -c1-330 | tail -30
