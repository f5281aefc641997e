#!/bin/bash
# ncu --set full of the pooled-step kernels after the v14/v15 changes (one launch each)
mkdir -p gpurun_out
NCU="ncu --set full --import-source on --clock-control none --launch-count 1 -f"
export C3_POOLED=16 C3_STEPS=12
python scripts/prof_c3.py > /dev/null 2>&1 || exit 1
for k in kProposePooledTile kPoolAccumulateDmma kAcceptLocal; do
  timeout 600 $NCU -k regex:$k --launch-skip 8 -o gpurun_out/prof_$k python scripts/prof_c3.py > gpurun_out/ncu_$k.log 2>&1; tail -1 gpurun_out/ncu_$k.log
done
