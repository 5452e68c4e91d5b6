#!/bin/bash
# round 2, call 4 (2 GPUs): sharded chain parity (NCCL) and the sharded bench line at chi = 128 and chi = 64
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_sharded.py -q -s > gpurun_out/c4_sharded_tests.log 2>&1; echo "sharded tests rc=$?"; tail -12 gpurun_out/c4_sharded_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_chi128_n2.json 2> gpurun_out/c4_bench_n2.err; echo "bench n2 rc=$?"; cut -c1-3500 gpurun_out/r2_bench_chi128_n2.json; tail -8 gpurun_out/c4_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --chi 64 --steps 10 --warmup 3 > gpurun_out/r2_bench_chi64_n2.json 2> gpurun_out/c4_bench64_n2.err; echo "bench64 n2 rc=$?"; cut -c1-2500 gpurun_out/r2_bench_chi64_n2.json; tail -5 gpurun_out/c4_bench64_n2.err
