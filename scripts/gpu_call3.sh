#!/bin/bash
# round 2, call 3 (1 GPU): whole GPU suite (no -x), full ncu capture of the chi = 128 main contraction (TMA GEMM)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c3_gpu_tests.log 2>&1; echo "gpu suite rc=$?"; tail -25 gpurun_out/c3_gpu_tests.log
CHI=128 WARM=4 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"grouped_gemm_tma_kernel<1, 128" -s 4 -c 1 -o gpurun_out/r2_tma_contraction_chi128 python scripts/ncu_step.py 1 > gpurun_out/c3_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/c3_ncu_full.log
