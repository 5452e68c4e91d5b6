#!/bin/bash
# round 2, call 9 (2 GPUs): sharded tests, chi = 64 chain test, sharded ATRG chi = 128 with the rank-certificate trace, bench --gpus 2
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_chains.py tests/test_gpu_whiten.py -m gpu -q ) > gpurun_out/c9_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/c9_pytest.log | cut -c1-220
GTN_DEBUG_TRUNC=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 scripts/atrg_sharded.py --chi 128 --steps 4 --check --out gpurun_out/r2_atrg_sharded_chi128_n2.json > gpurun_out/c9_atrg128.log 2>&1; echo "atrg128 rc=$?"; grep -E "^\{\"step|trunc|Error" gpurun_out/c9_atrg128.log | cut -c1-330 | tail -60
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/c9_bench_n2.json 2> gpurun_out/c9_bench_n2.err; echo "bench n2 rc=$?"; tail -c 2500 gpurun_out/c9_bench_n2.json; tail -3 gpurun_out/c9_bench_n2.err
