#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_edges.py -m gpu -x -q -k "whole_step_graph or unwritten or speculative or reproducible" ) > gpurun_out/c42_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c42_pytest.log | cut -c1-300
CHIS=32,64 timeout 400 python scripts/lazy_ab.py > gpurun_out/c42_lazy_ab.log 2>&1; echo "ab rc=$?"; tail -4 gpurun_out/c42_lazy_ab.log | cut -c1-1800
