#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_edges.py -m gpu -x -q -k conj_transposed 2>&1 | tail -3 | cut -c1-300
GTN_GEMM_CONJ_TRANS=1 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/g_pytest_ct.log 2>&1; tail -3 gpurun_out/g_pytest_ct.log | cut -c1-300
for ct in 0 1; do
GTN_GEMM_CONJ_TRANS=$ct timeout 300 python bench.py --no-micro > gpurun_out/g_bench_ct$ct.json 2>/dev/null
python - <<EOF
import json
d=json.load(open('gpurun_out/g_bench_ct$ct.json'))
ks=d['extra']['kernel_shares']
print("CONJ_TRANS=$ct", round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['gpu_launches'], d['extra']['speculation'],
      {k:(round(ks[k]['ms_per_step'],3), ks[k]['launches_per_step']) for k in ('grouped_gemm','sign_permute')})
EOF
done
