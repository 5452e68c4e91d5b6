#!/bin/bash
mkdir -p gpurun_out
for ov in 32 24; do
GTN_TRUNC_OVERSAMPLE=$ov timeout 300 python bench.py --no-micro > gpurun_out/d_bench_ov$ov.json 2>/dev/null
python - <<EOF
import json
d=json.load(open('gpurun_out/d_bench_ov$ov.json'))
print("OVERSAMPLE=$ov", d['value'], d['ms_per_step'], d['e2e']['value'], d['extra']['step_graph'], d['extra']['speculation'], d['extra']['trunc_refinements_last'])
print({k:(round(v['ms_per_step'],3), v['launches_per_step']) for k,v in d['extra']['kernel_shares'].items()})
EOF
done
