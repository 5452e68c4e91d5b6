#!/bin/bash
mkdir -p gpurun_out
timeout 400 python scripts/big_chi.py --chi 128 --steps 9 --out gpurun_out/r1c_trg_chi128_chain.json 2>&1 | cut -c1-330 | tail -9
