#!/bin/bash
# round 2, call 44: the round-end sequence at HEAD (GPU suite, smoke, bench, hotrg profile)
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/final3_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/final3_pytest.log | cut -c1-200
timeout 300 python __graft_entry__.py smoke > gpurun_out/final3_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/final3_smoke.log | cut -c1-300
( time timeout 400 python bench.py ) > gpurun_out/final3_bench.json 2> gpurun_out/final3_bench.err; echo "bench rc=$?"; head -1 gpurun_out/final3_bench.json | cut -c1-400
timeout 120 python scripts/hotrg_profile.py > gpurun_out/r2i_hotrg3dz_zcut64_profile_unwritten.txt 2>&1; echo "hotrg rc=$?"; head -3 gpurun_out/r2i_hotrg3dz_zcut64_profile_unwritten.txt | cut -c1-300
