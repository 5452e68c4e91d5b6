#!/bin/bash
# round 2, call 14 (2 GPUs): whole GPU suite (compact packing, wide-subspace pre-rotation, gram_rotate early stop),
# per-launch probe of a sharded step, N = 1 benches at chi = 128 and chi = 32
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/c14_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/c14_pytest.log | cut -c1-220
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/sharded_launch_probe.py > gpurun_out/c14_launch_probe_n2.log 2>&1; echo "probe rc=$?"; grep -v "^\[\|Warning\|warn\|^\*\|OMP" gpurun_out/c14_launch_probe_n2.log | head -60
( time timeout 600 python bench.py --no-micro ) > gpurun_out/c14_bench_chi128.json 2> gpurun_out/c14_bench_chi128.err; echo "bench128 rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c14_bench_chi128.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['extra']['whole_step_vs_fp64_yardstick']['frac_of_yardstick'])
for k,v in d['extra']['kernel_shares'].items(): print(' ', k, round(v['ms_per_step'],2), v['launches_per_step'])
PY
( time timeout 600 python bench.py --chi 32 --no-micro ) > gpurun_out/c14_bench_chi32.json 2> gpurun_out/c14_bench_chi32.err; echo "bench32 rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c14_bench_chi32.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'])
for k,v in d['extra']['kernel_shares'].items(): print(' ', k, round(v['ms_per_step'],3), v['launches_per_step'])
print({k: v for k, v in d['extra'].get('other_workloads', {}).items() if k.endswith('_ms')})
PY
