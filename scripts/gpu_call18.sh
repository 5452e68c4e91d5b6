#!/bin/bash
# round 2, call 18 (2 GPUs): whole GPU suite, default bench (streamed e2e, 512-wide sweep), reference arm, bench --gpus 2
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/c18_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/c18_pytest.log | cut -c1-220
( time timeout 900 python bench.py ) > gpurun_out/c18_bench.json 2> gpurun_out/c18_bench.err; echo "bench rc=$?"; tail -4 gpurun_out/c18_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c18_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['cpu_baseline'])
print(json.dumps(d['extra']['other_workloads']['einsum_sweep'])[:1800])
print({k: v for k, v in d['extra']['other_workloads'].items() if k.endswith('_ms')})
PY
( time timeout 600 python bench.py --impl reference ) > gpurun_out/c18_bench_reference.json 2> gpurun_out/c18_bench_reference.err; echo "reference rc=$?"; tail -c 700 gpurun_out/c18_bench_reference.json
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/c18_bench_n2.json 2> gpurun_out/c18_bench_n2.err; echo "bench n2 rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c18_bench_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['sharded']['tnorm_rel_diff_vs_single_gpu'])
PY
