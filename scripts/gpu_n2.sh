#!/bin/bash
mkdir -p gpurun_out
( time timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/c_bench_n2.json 2> gpurun_out/c_bench_n2.err
tail -5 gpurun_out/c_bench_n2.err | cut -c1-300
python - <<EOF
import json
line=[l for l in open('gpurun_out/c_bench_n2.json') if l.startswith('{')][-1]
d=json.loads(line)
print({k:d[k] for k in ['value','ms_per_step','n_gpus','gpu_launches']}, d['e2e']['value'], d['extra']['sharded_contraction'], d['extra']['step_graph'])
EOF
