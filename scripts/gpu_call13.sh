#!/bin/bash
# round 2, call 13 (8 GPUs): sharded bench (TRG chi = 128, strong scaling) and BASELINE config 4: ATRG chi = 256 sharded over 8 GPUs
N=8
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 5 --warmup 3 ) > gpurun_out/c13_bench_n$N.json 2> gpurun_out/c13_bench_n$N.err; echo "bench n$N rc=$?"; tail -c 600 gpurun_out/c13_bench_n$N.json; tail -3 gpurun_out/c13_bench_n$N.err
GTN_DEBUG_TRUNC=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 scripts/atrg_sharded.py --chi 256 --steps 4 --out gpurun_out/r2c_atrg_sharded_chi256_n$N.json > gpurun_out/c13_atrg256_n$N.log 2>&1; echo "atrg256 rc=$?"; grep -E "^\{\"step|Error|trunc sharded|rank cert|one-call" gpurun_out/c13_atrg256_n$N.log | cut -c1-300 | tail -60
