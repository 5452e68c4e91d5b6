#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/launches_r1c.csv python scripts/ncu_step.py 3 > gpurun_out/c_ncu_list.log 2>&1; tail -1 gpurun_out/c_ncu_list.log
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:"jacobi_persistent|gram_rotate|chol_whiten|grouped_gemm" -c 14 -o gpurun_out/r1c_step python scripts/ncu_step.py 1 > gpurun_out/c_ncu_full.log 2>&1; tail -1 gpurun_out/c_ncu_full.log
timeout 300 python scripts/timeline.py --steps 5 > gpurun_out/r1c_trg_chi32_timeline.txt 2>&1; tail -3 gpurun_out/r1c_trg_chi32_timeline.txt
