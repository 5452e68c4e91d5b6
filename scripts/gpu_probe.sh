#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/step_probe2.py 2>&1 | tail -4 | cut -c1-1300
