#!/bin/bash
# round 2, call 43 (2 GPUs): sharded tests and the sharded bench with the unwritten-permutation path
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x > gpurun_out/c43_sharded.log 2>&1; echo "sharded test rc=$?"; tail -1 gpurun_out/c43_sharded.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29593 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2i_bench_chi128_n2.json 2> gpurun_out/c43_bench_n2.err; echo "bench rc=$?"; tail -1 gpurun_out/r2i_bench_chi128_n2.json | cut -c1-900
