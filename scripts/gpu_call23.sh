#!/bin/bash
# round 2, call 23 (8 GPUs): bench --gpus 8 with the fused GEMM + all-reduce over peer memory and with NCCL all-reduces
mkdir -p gpurun_out
for F in 1 0; do
( time GTN_FUSED_ALLREDUCE=$F timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2956$F bench.py --gpus 8 --steps 8 --warmup 3 ) > gpurun_out/c23_bench_n8_fused$F.json 2> gpurun_out/c23_bench_n8_fused$F.err; echo "bench n8 fused=$F rc=$?"; tail -2 gpurun_out/c23_bench_n8_fused$F.err; python - <<PY
import json
d=json.loads(open('gpurun_out/c23_bench_n8_fused$F.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['sharded']['tnorm_rel_diff_vs_single_gpu'], d['sharded']['collective_ms_per_step'])
print({k: round(v['ms_per_step'],2) for k,v in d['extra']['kernel_shares'].items()})
PY
done
