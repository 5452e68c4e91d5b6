#!/bin/bash
# round 2, call 35 (2 GPUs): new rank-certificate test; the 2-GPU sharded test three times (flakiness of the multiplet-cut comparison)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_edges.py -m gpu -q -x > gpurun_out/c35_edges.log 2>&1; echo "edges rc=$?"; tail -3 gpurun_out/c35_edges.log | cut -c1-250
for i in 1 2 3; do
  timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x -s > gpurun_out/c35_sharded_$i.log 2>&1; echo "sharded run $i rc=$?"; grep -o "z2 chi32 \[[^]]*\]" gpurun_out/c35_sharded_$i.log | cut -c1-300; tail -1 gpurun_out/c35_sharded_$i.log
done
