#!/bin/bash
# ncu --set full captures of the current top kernels (one launch each), after a plain run of each script
mkdir -p gpurun_out
NCU="ncu --set full --import-source on --clock-control none --launch-count 1 -f"
python scripts/prof_pairs.py > /dev/null 2>&1 && timeout 600 $NCU -k regex:kFakePairs$ --launch-skip 3 -o gpurun_out/prof_pairs_v12 python scripts/prof_pairs.py > gpurun_out/ncu_pairs.log 2>&1; tail -1 gpurun_out/ncu_pairs.log
python scripts/prof_stream.py > /dev/null 2>&1 && timeout 600 $NCU -k regex:kFakeStream --launch-skip 3 -o gpurun_out/prof_stream python scripts/prof_stream.py > gpurun_out/ncu_stream.log 2>&1; tail -1 gpurun_out/ncu_stream.log
python scripts/prof_hmc.py > /dev/null 2>&1 && timeout 600 $NCU -k regex:kHmcExxtUpdate --launch-skip 1 -o gpurun_out/prof_exxt python scripts/prof_hmc.py > gpurun_out/ncu_exxt.log 2>&1; tail -1 gpurun_out/ncu_exxt.log
C3_POOLED=16 C3_STEPS=12 python scripts/prof_c3.py > /dev/null 2>&1 && C3_POOLED=16 C3_STEPS=12 timeout 600 $NCU -k regex:kProposePooledTile --launch-skip 8 -o gpurun_out/prof_pooled_tile python scripts/prof_c3.py > gpurun_out/ncu_ptile.log 2>&1; tail -1 gpurun_out/ncu_ptile.log
C3_STEPS=40 python scripts/prof_c3.py > /dev/null 2>&1 && C3_STEPS=40 timeout 600 $NCU -k regex:kProposeStaged --launch-skip 35 -o gpurun_out/prof_propose_staged python scripts/prof_c3.py > gpurun_out/ncu_staged.log 2>&1; tail -1 gpurun_out/ncu_staged.log
ls -la gpurun_out/*.ncu-rep | tail -8
