#!/bin/bash
# round 2, call 37 (1 GPU): final confirmation of the committed state: GPU suite (-x), smoke, bench at chi = 64 and the default bench
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final_pytest.log | cut -c1-200
timeout 300 python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/final_smoke.log
( time timeout 600 python bench.py --chi 64 --no-micro ) > gpurun_out/final_bench_chi64.json 2> gpurun_out/final_bench_chi64.err; echo "bench64 rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/final_bench_chi64.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'], d['cpu_baseline'])
PY
( time timeout 900 python bench.py ) > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['gpu_launches'])
print(d['extra'].get('trg_block_moving_chain_chi128_ms'), d['extra'].get('atrg_block_chi128_ms_per_step'))
PY
