#!/bin/bash
# round 2, call 17 (1 GPU): packed Cholesky whitening after the rewrite (tests, probe), truncated-SVD tests, bench chi = 128
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_whiten.py tests/test_chains.py tests/test_gpu_sector_onecall.py tests/test_gpu_edges.py -m gpu -q ) > gpurun_out/c17_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/c17_pytest.log | cut -c1-220
timeout 120 python scripts/whiten_probe.py > gpurun_out/c17_whiten_probe.log 2>&1; tail -6 gpurun_out/c17_whiten_probe.log
( time timeout 600 python bench.py --no-micro ) > gpurun_out/c17_bench_chi128.json 2> gpurun_out/c17_bench_chi128.err; echo "bench128 rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c17_bench_chi128.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['extra']['whole_step_vs_fp64_yardstick']['frac_of_yardstick'])
for k,v in d['extra']['kernel_shares'].items(): print(' ', k, round(v['ms_per_step'],2), v['launches_per_step'])
PY
