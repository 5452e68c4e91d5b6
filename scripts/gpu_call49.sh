#!/bin/bash
# round 2, call 49: the example scripts at HEAD (ATRG dense, TRG block with log + checkpoint + resume)
mkdir -p gpurun_out
timeout 30 python examples/example.py --cgsteps 4 > gpurun_out/c49_example_atrg.log 2>&1; echo "example rc=$?"; tail -3 gpurun_out/c49_example_atrg.log | cut -c1-200
timeout 30 python examples/example_block.py --trg --cgsteps 4 --log gpurun_out/c49_run.jsonl --checkpoint gpurun_out/c49_ckpt > gpurun_out/c49_example_block.log 2>&1; echo "block rc=$?"; tail -2 gpurun_out/c49_example_block.log | cut -c1-200
