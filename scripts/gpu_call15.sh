#!/bin/bash
# round 2, call 15 (8 GPUs): BASELINE config 4: ATRG chi = 256 sharded over 8 GPUs; whitening kernel probe
N=8
mkdir -p gpurun_out
GTN_DEBUG_TRUNC=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 scripts/atrg_sharded.py --chi 256 --steps 4 --out gpurun_out/r2c_atrg_sharded_chi256_n$N.json > gpurun_out/c15_atrg256_n$N.log 2>&1; echo "atrg256 rc=$?"; grep -E "^\{\"step|Error|trunc sharded|rank cert|one-call" gpurun_out/c15_atrg256_n$N.log | cut -c1-300 | tail -50
timeout 120 python scripts/whiten_probe.py > gpurun_out/c15_whiten_probe.log 2>&1; cat gpurun_out/c15_whiten_probe.log | tail -6
