#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/h_pytest.log 2>&1; tail -2 gpurun_out/h_pytest.log | cut -c1-300
for es in 0 1e-10; do
GTN_JACOBI_EARLY_STOP=$es timeout 300 python bench.py --no-micro > gpurun_out/h_bench_es$es.json 2>/dev/null
python - <<EOF
import json
d=json.load(open('gpurun_out/h_bench_es$es.json'))
ks=d['extra']['kernel_shares']
print("EARLY_STOP=$es", round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['extra']['speculation'], d['extra']['jacobi_sweeps_last'],
      {k:round(ks[k]['ms_per_step'],3) for k in ('gram_rotate','jacobi_persistent')})
EOF
done
