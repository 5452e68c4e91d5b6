#!/bin/bash
# round 2, call 5 (1 GPU): whitening variants, rank-rule probe, chi=128 bench with the packed whitening, main-contraction ncu
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_whiten.py -q > gpurun_out/c5_whiten.log 2>&1; echo "whiten rc=$?"; tail -8 gpurun_out/c5_whiten.log
timeout 300 python scripts/rank_probe.py > gpurun_out/c5_rank_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/c5_rank_probe.log | cut -c1-300
timeout 900 python bench.py --steps 10 --warmup 3 --no-micro --no-cpu > gpurun_out/c5_bench_chi128.json 2> gpurun_out/c5_bench128.err; echo "bench128 rc=$?"; cut -c1-600 gpurun_out/c5_bench_chi128.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/c5_bench_chi128.json'))
print({k:(round(v['ms_per_step'],2)) for k,v in d['extra']['kernel_shares'].items()})
PY
CHI=128 WARM=4 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:grouped_gemm_tma_kernel -s 6 -c 1 -o gpurun_out/r2_tma_contraction_chi128 python scripts/ncu_step.py 1 > gpurun_out/c5_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/c5_ncu_full.log
