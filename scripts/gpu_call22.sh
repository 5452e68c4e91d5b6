#!/bin/bash
# round 2, call 22 (2 GPUs): fused GEMM + all-reduce with the library's own peer barrier: tests, bench --gpus 2 with / without it
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q -s ) > gpurun_out/c22_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c22_pytest.log | cut -c1-300
for F in 1 0; do
( time GTN_FUSED_ALLREDUCE=$F timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2955$F bench.py --gpus 2 --steps 8 --warmup 3 ) > gpurun_out/c22_bench_n2_fused$F.json 2> gpurun_out/c22_bench_n2_fused$F.err; echo "bench n2 fused=$F rc=$?"; tail -2 gpurun_out/c22_bench_n2_fused$F.err; python - <<PY
import json
d=json.loads(open('gpurun_out/c22_bench_n2_fused$F.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['sharded']['tnorm_rel_diff_vs_single_gpu'], d['sharded']['collective_ms_per_step'])
print({k: round(v['ms_per_step'],2) for k,v in d['extra']['kernel_shares'].items()})
PY
done
