#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/c41_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c41_pytest.log | cut -c1-300
timeout 400 python scripts/lazy_ab.py > gpurun_out/c41_lazy_ab.log 2>&1; echo "ab rc=$?"; tail -5 gpurun_out/c41_lazy_ab.log | cut -c1-1500
