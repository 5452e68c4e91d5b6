#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/e_pytest.log 2>&1; tail -3 gpurun_out/e_pytest.log | head -1
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
( time timeout 600 python bench.py --chi 64 --no-micro ) > gpurun_out/e_bench_chi64.json 2> gpurun_out/e_bench_chi64.err; tail -4 gpurun_out/e_bench_chi64.err
python - <<EOF
import json
d=json.load(open('gpurun_out/e_bench_chi64.json'))
print({k:d[k] for k in ['value','ms_per_step','gpu_launches']}, d['e2e']['value'], d['cpu_baseline'], d['extra']['step_graph'], d['extra']['speculation'], d['extra']['svd_paths'])
EOF
