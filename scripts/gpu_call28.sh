#!/bin/bash
# round 2, call 28 (1 GPU): final ncu evidence of the chi = 128 step: launch list (shares) and a full capture of the
# 32x32 panel GEMM of the subspace iteration (the second family of the step)
mkdir -p gpurun_out
CHI=128 WARM=4 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2g_trg_chi128_launches.csv python scripts/ncu_step.py 2 > gpurun_out/c28_ncu_list.log 2>&1; echo "ncu list rc=$?"; tail -2 gpurun_out/c28_ncu_list.log
CHI=128 WARM=4 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:grouped_gemm_kernel -s 4 -c 1 -o gpurun_out/r2g_panel_gemm_chi128 -f python scripts/ncu_step.py 1 > gpurun_out/c28_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/c28_ncu_full.log
