#!/bin/bash
# round 2, call 45 (2 GPUs): sharded ATRG chi = 128 with the unwritten-permutation path: time and peak memory per rank
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29595 scripts/atrg_sharded.py --chi 128 --steps 6 --out gpurun_out/r2i_atrg_sharded_chi128_n2.json > gpurun_out/c45_atrg128.log 2>&1; echo "atrg128 rc=$?"; grep -E "^\{\"step|Error" gpurun_out/c45_atrg128.log | cut -c1-240 | tail -8
