#!/bin/bash
mkdir -p gpurun_out
for tol in 1e-12 1e-9 1e-6; do
GTN_ROTATE_TOL=$tol timeout 300 python bench.py --no-micro > gpurun_out/f_bench_rot$tol.json 2>/dev/null
python - <<EOF
import json
d=json.load(open('gpurun_out/f_bench_rot$tol.json'))
ks=d['extra']['kernel_shares']
print("ROTATE_TOL=$tol", round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['extra']['speculation'], d['extra']['jacobi_sweeps_last'],
      {k:round(ks[k]['ms_per_step'],3) for k in ('gram_rotate','jacobi_persistent','chol_whiten')})
EOF
done
