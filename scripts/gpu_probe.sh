#!/bin/bash
mkdir -p gpurun_out
echo "== default (fast iteration on)"; timeout 300 python scripts/step_probe2.py 2>&1 | tail -4 | cut -c1-900
echo "== GTN_FAST_ITER=0"; GTN_FAST_ITER=0 timeout 300 python scripts/step_probe2.py 2>&1 | tail -4 | cut -c1-900
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/b_pytest.log 2>&1; tail -8 gpurun_out/b_pytest.log | cut -c1-400
