#!/bin/bash
mkdir -p gpurun_out
timeout 40 python bench.py > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err; echo "rc $?"; tail -2 gpurun_out/j_bench.err | cut -c1-300
