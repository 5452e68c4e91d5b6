#!/bin/bash
# round 2, call 11 (1 GPU): whole GPU suite (split-K GEMM, fused hotrg3dz chains, Jacobi kernel changes), hotrg3dz profile
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/c11_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/c11_pytest.log | cut -c1-220
timeout 300 python scripts/hotrg_profile.py > gpurun_out/c11_hotrg_profile.log 2>&1; echo "hotrg rc=$?"; grep -E "ms/step|input" gpurun_out/c11_hotrg_profile.log; sed -n 5,16p gpurun_out/c11_hotrg_profile.log | cut -c1-60,150-230
