#!/bin/bash
# round 2, call 21 (2 GPUs): sharded tests with the fused GEMM + all-reduce over peer memory, bench --gpus 2, bench N = 1 (streamed e2e)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_checkpoint.py -m gpu -q -s ) > gpurun_out/c21_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/c21_pytest.log | cut -c1-600
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 8 --warmup 3 ) > gpurun_out/c21_bench_n2.json 2> gpurun_out/c21_bench_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/c21_bench_n2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c21_bench_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['serial_value'], d['sharded'])
for k,v in d['extra']['kernel_shares'].items(): print(' ', k, round(v['ms_per_step'],2), v['launches_per_step'])
PY
( time timeout 600 python bench.py --no-micro --steps 10 ) > gpurun_out/c21_bench_chi128.json 2> gpurun_out/c21_bench_chi128.err; echo "bench128 rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c21_bench_chi128.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['serial_value'])
PY
