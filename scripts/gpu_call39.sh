#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/final2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final2_pytest.log | cut -c1-200
timeout 300 python __graft_entry__.py smoke > gpurun_out/final2_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/final2_smoke.log
