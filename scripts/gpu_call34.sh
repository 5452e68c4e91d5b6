#!/bin/bash
# round 2, call 34 (1 GPU): the GPU suite four times in a row (flakiness check at the round's final state)
mkdir -p gpurun_out
for i in 1 2 3 4; do
  timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c34_pytest_$i.log 2>&1; echo "run $i rc=$?"; tail -2 gpurun_out/c34_pytest_$i.log | cut -c1-200
done
