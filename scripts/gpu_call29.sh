#!/bin/bash
# round 2, call 29 (1 GPU): packed Cholesky with physical pivoting: whitening tests, probe, whole GPU suite, bench
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_whiten.py -m gpu -q ) > gpurun_out/c29_whiten.log 2>&1; echo "whiten rc=$?"; tail -5 gpurun_out/c29_whiten.log | cut -c1-250
timeout 120 python scripts/whiten_probe.py 96 128 > gpurun_out/c29_whiten_probe.log 2>&1; tail -3 gpurun_out/c29_whiten_probe.log
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/c29_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/c29_pytest.log | cut -c1-250
( time timeout 600 python bench.py --no-micro ) > gpurun_out/c29_bench_chi128.json 2> gpurun_out/c29_bench_chi128.err; echo "bench128 rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c29_bench_chi128.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'])
print({k: round(v['ms_per_step'],2) for k,v in d['extra']['kernel_shares'].items()})
PY
