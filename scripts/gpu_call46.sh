#!/bin/bash
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_edges.py -m gpu -x -q -k "unwritten" ) > gpurun_out/c46_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c46_pytest.log | cut -c1-300
