#!/bin/bash
# round 2, call 48: bench.py at HEAD (the permute micro-benchmarks force their result to be written)
mkdir -p gpurun_out
( time timeout 300 python bench.py ) > gpurun_out/final4_bench.json 2> gpurun_out/final4_bench.err; echo "bench rc=$?"; head -1 gpurun_out/final4_bench.json | cut -c1-300
