#!/bin/bash
# round 2, call 33 (8 GPUs): bench --gpus 8 at the round's final state
mkdir -p gpurun_out
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --steps 10 --warmup 3 ) > gpurun_out/c33_bench_n8.json 2> gpurun_out/c33_bench_n8.err; echo "bench n8 rc=$?"; tail -2 gpurun_out/c33_bench_n8.err; python - <<PY
import json
d=json.loads(open('gpurun_out/c33_bench_n8.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['serial_value'], d['sharded']['tnorm_rel_diff_vs_single_gpu'], d['sharded']['collective_ms_per_step'])
print({k: round(v['ms_per_step'],2) for k,v in d['extra']['kernel_shares'].items()})
PY
