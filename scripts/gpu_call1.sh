#!/bin/bash
# round 2, call 1: TMA GEMM correctness + tile-configuration bench + whole GPU suite + chi=128 chain
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt
timeout 600 python -m pytest tests/test_gpu_gemm_tma.py -x -q > gpurun_out/c1_tma_tests.log 2>&1; echo "tma tests rc=$?"; tail -5 gpurun_out/c1_tma_tests.log
timeout 400 python scripts/gemm_bench.py > gpurun_out/c1_gemm_bench.log 2>&1; echo "gemm bench rc=$?"; tail -12 gpurun_out/c1_gemm_bench.log | cut -c1-600
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c1_gpu_tests.log 2>&1; echo "gpu suite rc=$?"; tail -8 gpurun_out/c1_gpu_tests.log
timeout 400 python scripts/big_chi.py --chi 128 --steps 8 --out gpurun_out/r2_trg_chi128_chain.json > gpurun_out/c1_chi128.log 2>&1; echo "chi128 rc=$?"; tail -8 gpurun_out/c1_chi128.log | cut -c1-400
