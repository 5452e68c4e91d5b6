#!/bin/bash
# round 2, call 26 (1 GPU): compute-sanitizer memcheck over the kernels added / rewritten this round (small tests), then the
# one-call tests and the default bench
mkdir -p gpurun_out
( time timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_whiten.py "tests/test_gpu_sector_onecall.py::test_onecall_svd_vs_lapack" "tests/test_gpu_sector_onecall.py::test_onecall_steep_spectrum_needs_the_robust_mode" "tests/test_chains.py::test_gpu_einsum_sweep_large_dims_vs_oracle" tests/test_tensor_prep.py -m gpu -q -x ) > gpurun_out/c26_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|error" gpurun_out/c26_memcheck.log | head -12
( time timeout 900 python -m pytest tests/test_gpu_sector_onecall.py tests/test_gpu_edges.py -m gpu -q ) > gpurun_out/c26_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c26_pytest.log | cut -c1-200
( time timeout 900 python bench.py ) > gpurun_out/c26_bench.json 2> gpurun_out/c26_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/c26_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c26_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['extra'].get('atrg_block_chi128_ms_per_step'))
print({k: round(v,2) for k, v in d['extra']['other_workloads'].items() if k.endswith('_ms')})
PY
