#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/b_pytest.log 2>&1; tail -12 gpurun_out/b_pytest.log | cut -c1-400
for r in 1 2 3 4; do
timeout 600 python -m pytest tests/test_gpu_edges.py -m gpu -x -q -k "whole_step or speculative" > gpurun_out/b_pytest_$r.log 2>&1; echo "run $r"; tail -2 gpurun_out/b_pytest_$r.log | cut -c1-300
done
timeout 300 python scripts/step_probe.py > gpurun_out/step_probe.log 2>&1
tail -12 gpurun_out/step_probe.log
