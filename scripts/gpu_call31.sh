#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/atrg_launch_probe.py > gpurun_out/c31_atrg_probe.log 2>&1; echo "rc=$?"; grep -v "Warning\|warn" gpurun_out/c31_atrg_probe.log | head -90
