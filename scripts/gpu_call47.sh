#!/bin/bash
# round 2, call 47 (2 GPUs): sharded ATRG chi = 192 on TWO GPUs (4 before: 44.5 GiB per rank) with the unwritten permutations
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29597 scripts/atrg_sharded.py --chi 192 --steps 3 --out gpurun_out/r2i_atrg_sharded_chi192_n2.json > gpurun_out/c47_atrg192.log 2>&1; echo "atrg192 rc=$?"; grep -E "^\{\"step|Error" gpurun_out/c47_atrg192.log | cut -c1-240 | tail -8
