#!/bin/bash
# round 2, call 27 (4 GPUs): bench --gpus 4 (defaults of the round end)
mkdir -p gpurun_out
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 4 --steps 10 --warmup 3 ) > gpurun_out/c27_bench_n4.json 2> gpurun_out/c27_bench_n4.err; echo "bench n4 rc=$?"; tail -2 gpurun_out/c27_bench_n4.err; python - <<PY
import json
d=json.loads(open('gpurun_out/c27_bench_n4.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['serial_value'], d['sharded']['tnorm_rel_diff_vs_single_gpu'], d['sharded']['collective_ms_per_step'])
print({k: round(v['ms_per_step'],2) for k,v in d['extra']['kernel_shares'].items()})
PY
