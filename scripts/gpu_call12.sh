#!/bin/bash
# round 2, call 12 (N GPUs, N = $1): sharded bench (TRG chi = 128, strong scaling) and the sharded ATRG chain at chi = 128
N=${1:-4}
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 5 --warmup 3 ) > gpurun_out/c12_bench_n$N.json 2> gpurun_out/c12_bench_n$N.err; echo "bench n$N rc=$?"; tail -c 1500 gpurun_out/c12_bench_n$N.json; tail -3 gpurun_out/c12_bench_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 scripts/atrg_sharded.py --chi 128 --steps 6 --out gpurun_out/r2c_atrg_sharded_chi128_n$N.json > gpurun_out/c12_atrg128_n$N.log 2>&1; echo "atrg128 rc=$?"; grep -E "^\{\"step|Error" gpurun_out/c12_atrg128_n$N.log | cut -c1-330 | tail -8
