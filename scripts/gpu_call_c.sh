#!/bin/bash
mkdir -p gpurun_out
( time timeout 300 python bench.py ) > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err; tail -3 gpurun_out/i_bench.err
