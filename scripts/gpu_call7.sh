#!/bin/bash
# round 2, call 7 (1 GPU): rank-rule probe on the chi = 64 chain, the whole GPU suite (no -x), default bench
mkdir -p gpurun_out
timeout 300 python scripts/rank_probe2.py > gpurun_out/c7_rank_probe2.log 2>&1; echo "probe rc=$?"; tail -30 gpurun_out/c7_rank_probe2.log | cut -c1-200
( time timeout 1500 python -m pytest tests -m gpu -q --durations=10 ) > gpurun_out/c7_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/c7_pytest.log | cut -c1-220
( time timeout 600 python bench.py ) > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/c7_bench.json; tail -5 gpurun_out/c7_bench.err
