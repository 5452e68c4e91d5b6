#!/bin/bash
# round 2, call 19 (1 GPU): reproducibility + streaming tests, chi = 32 kernel timeline (regression check), bench chi = 128
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_edges.py tests/test_checkpoint.py tests/test_gpu_parity.py -m gpu -q ) > gpurun_out/c19_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/c19_pytest.log | cut -c1-250
timeout 300 python scripts/timeline.py --chi 32 > gpurun_out/c19_timeline_chi32.log 2>&1; sed -n 3,16p gpurun_out/c19_timeline_chi32.log | cut -c1-60,150-240
( time timeout 600 python bench.py --no-micro ) > gpurun_out/c19_bench_chi128.json 2> gpurun_out/c19_bench_chi128.err; echo "bench128 rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c19_bench_chi128.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['serial_value'])
PY
