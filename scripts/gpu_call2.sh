#!/bin/bash
# round 2, call 2 (1 GPU): whole GPU suite with the new chain tests, bench at chi = 128 (headline) and chi = 32,
# ncu launch list of the chi = 128 step and a full capture of its main contraction (TMA GEMM)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c2_gpu_tests.log 2>&1; echo "gpu suite rc=$?"; tail -15 gpurun_out/c2_gpu_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_chi128.json 2> gpurun_out/c2_bench128.err; echo "bench128 rc=$?"; tail -c 3000 gpurun_out/r2_bench_chi128.json; tail -5 gpurun_out/c2_bench128.err
timeout 600 python bench.py --chi 32 --steps 20 --warmup 5 --no-micro > gpurun_out/r2_bench_chi32.json 2> gpurun_out/c2_bench32.err; echo "bench32 rc=$?"; cut -c1-1500 gpurun_out/r2_bench_chi32.json; tail -5 gpurun_out/c2_bench32.err
CHI=128 WARM=6 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_trg_chi128_launches.csv python scripts/ncu_step.py 2 > gpurun_out/c2_ncu_list.log 2>&1; echo "ncu list rc=$?"; tail -3 gpurun_out/c2_ncu_list.log
CHI=128 WARM=4 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:grouped_gemm_tma -c 6 -o gpurun_out/r2_tma_gemm_chi128 python scripts/ncu_step.py 1 > gpurun_out/c2_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/c2_ncu_full.log
ls -la gpurun_out | tail -20
