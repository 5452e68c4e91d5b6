#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_whiten.py -m gpu -q > gpurun_out/c38_whiten.log 2>&1; echo "whiten rc=$?"; tail -4 gpurun_out/c38_whiten.log | cut -c1-300
timeout 200 python scripts/whiten_probe.py 128 160 256 > gpurun_out/c38_whiten_probe.log 2>&1; tail -4 gpurun_out/c38_whiten_probe.log
