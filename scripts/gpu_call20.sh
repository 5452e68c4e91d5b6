#!/bin/bash
# round 2, call 20 (2 GPUs): sharded tests on the C-ABI collectives (gtn_comm_*), bench --gpus 2, chi = 32 timeline after the revert
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_checkpoint.py -m gpu -q -s ) > gpurun_out/c20_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/c20_pytest.log | cut -c1-400
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 8 --warmup 3 ) > gpurun_out/c20_bench_n2.json 2> gpurun_out/c20_bench_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/c20_bench_n2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c20_bench_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['serial_value'], d['sharded']['tnorm_rel_diff_vs_single_gpu'], d['sharded']['collective_ms_per_step'])
PY
timeout 300 python scripts/timeline.py --chi 32 > gpurun_out/c20_timeline_chi32.log 2>&1; sed -n 3,12p gpurun_out/c20_timeline_chi32.log | cut -c1-60,150-240
