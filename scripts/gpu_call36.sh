#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_edges.py tests/test_chains.py tests/test_z2_golden.py -m gpu -q > gpurun_out/c36_tests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/c36_tests.log | cut -c1-300
