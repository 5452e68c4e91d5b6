#!/bin/bash
# round 2, call 6 (2 GPUs): sharded TRG + ATRG parity (NCCL), whitening variants, chi=64 chain test, sharded ATRG at chi = 64 / 128 with check
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_whiten.py tests/test_chains.py -q > gpurun_out/c6_tests.log 2>&1; echo "whiten+chains rc=$?"; tail -6 gpurun_out/c6_tests.log
timeout 900 python -m pytest tests/test_gpu_sharded.py -q -s > gpurun_out/c6_sharded_tests.log 2>&1; echo "sharded tests rc=$?"; tail -6 gpurun_out/c6_sharded_tests.log | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/atrg_sharded.py --chi 64 --steps 4 --check --out gpurun_out/r2_atrg_sharded_chi64_n2.json > gpurun_out/c6_atrg64.log 2>&1; echo "atrg64 rc=$?"; tail -6 gpurun_out/c6_atrg64.log | cut -c1-900
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 scripts/atrg_sharded.py --chi 128 --steps 4 --check --out gpurun_out/r2_atrg_sharded_chi128_n2.json > gpurun_out/c6_atrg128.log 2>&1; echo "atrg128 rc=$?"; tail -6 gpurun_out/c6_atrg128.log | cut -c1-900
