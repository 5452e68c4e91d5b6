#!/bin/bash
# round 2, call 30 (1 GPU): the round-end gate as the driver runs it (pytest -x, smoke, examples, bench, reference arm) + the
# CUPTI kernel timeline of the chi = 128 step
bash scripts/gpu_round_check.sh
( time timeout 600 python bench.py --impl reference ) > gpurun_out/check_bench_reference.json 2> gpurun_out/check_bench_reference.err; echo "reference rc=$?"; tail -c 400 gpurun_out/check_bench_reference.json
timeout 600 python scripts/timeline.py --chi 128 --steps 2 > gpurun_out/r2g_trg_chi128_timeline.txt 2>&1; sed -n 3,14p gpurun_out/r2g_trg_chi128_timeline.txt | cut -c1-70,150-240
python - <<'PY'
import json
d=json.loads(open('gpurun_out/check_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['cpu_baseline'])
print(d['extra'].get('trg_block_moving_chain_chi128_ms'), d['extra'].get('atrg_block_chi128_ms_per_step'))
PY
