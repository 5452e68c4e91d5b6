#!/bin/bash
# one GPU-box call: parity tests, smoke, examples, bench (results under gpurun_out/)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/check_gpu.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=12 ) > gpurun_out/check_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/check_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/check_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/check_smoke.log
( time timeout 200 python examples/example.py --cgsteps 6 ) > gpurun_out/check_example_atrg.log 2>&1
( time timeout 200 python examples/example_block.py --trg --cgsteps 6 --log gpurun_out/check_run.jsonl --checkpoint gpurun_out/check_ckpt ) > gpurun_out/check_example_block_trg.log 2>&1
rm -rf gpurun_out/check_ckpt
( time timeout 600 python bench.py ) > gpurun_out/check_bench.json 2> gpurun_out/check_bench.err
echo "bench exit $?" >> gpurun_out/check_bench.err
tail -3 gpurun_out/check_pytest.log; tail -2 gpurun_out/check_smoke.log; tail -c 600 gpurun_out/check_bench.err
