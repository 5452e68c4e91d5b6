#!/bin/bash
# round 2, call 10 (2 GPUs): one-call C-ABI tests, whole GPU suite, sharded ATRG chi = 128 with the robust path, default bench
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_sector_onecall.py -m gpu -q -x ) > gpurun_out/c10_onecall.log 2>&1; echo "onecall rc=$?"; tail -15 gpurun_out/c10_onecall.log | cut -c1-250
( time timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_sector_onecall.py ) > gpurun_out/c10_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/c10_pytest.log | cut -c1-220
GTN_DEBUG_TRUNC=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 scripts/atrg_sharded.py --chi 128 --steps 6 --check --out gpurun_out/r2_atrg_sharded_chi128_n2.json > gpurun_out/c10_atrg128.log 2>&1; echo "atrg128 rc=$?"; grep -E "^\{\"step|trunc sharded|rank cert|one-call|Error" gpurun_out/c10_atrg128.log | cut -c1-330 | tail -50
( time GTN_DEBUG_TRUNC=1 timeout 600 python bench.py --no-micro ) > gpurun_out/c10_bench.json 2> gpurun_out/c10_bench.err; echo "bench rc=$?"; grep -v "^\[trunc" gpurun_out/c10_bench.json | tail -c 1800; grep -c "one-call" gpurun_out/c10_bench.json; grep "one-call" gpurun_out/c10_bench.json | tail -4; tail -3 gpurun_out/c10_bench.err
timeout 300 python scripts/hotrg_profile.py > gpurun_out/c10_hotrg_profile.log 2>&1; echo "hotrg rc=$?"; grep -E "ms/step|input" gpurun_out/c10_hotrg_profile.log
