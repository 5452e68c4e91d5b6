#!/bin/bash
# round 2, call 25 (8 GPUs): ATRG chi = 256 sharded over 8 GPUs with the shifted-Cholesky-QR robust mode, bench --gpus 8 (defaults)
N=8
mkdir -p gpurun_out
GTN_DEBUG_TRUNC=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 scripts/atrg_sharded.py --chi 256 --steps 5 --out gpurun_out/r2f_atrg_sharded_chi256_n$N.json > gpurun_out/c25_atrg256_n$N.log 2>&1; echo "atrg256 rc=$?"; grep -E "^\{\"step|Error|trunc sharded" gpurun_out/c25_atrg256_n$N.log | cut -c1-260 | tail -40
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 10 --warmup 3 ) > gpurun_out/c25_bench_n$N.json 2> gpurun_out/c25_bench_n$N.err; echo "bench n$N rc=$?"; tail -2 gpurun_out/c25_bench_n$N.err; python - <<PY
import json
d=json.loads(open('gpurun_out/c25_bench_n$N.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['serial_value'], d['sharded']['tnorm_rel_diff_vs_single_gpu'], d['sharded']['collective_ms_per_step'])
PY
