#!/bin/bash
# same-box comparison of the C3 per-chain step: round-1 tree (_r01) against the current one
mkdir -p gpurun_out
O=gpurun_out
for rep in 1 2; do
  (cd _r01 && timeout 300 python scripts/configs_bench.py c3 2>/dev/null | grep '"mode": "per-chain"\|pooled' | cut -c1-230 | sed 's/^/r01: /')
  timeout 300 python scripts/configs_bench.py c3 2>/dev/null | grep '"mode": "per-chain"\|pooled' | cut -c1-230 | sed 's/^/r02: /'
done > $O/r02_o_c3_ab.txt 2>&1
cat $O/r02_o_c3_ab.txt
(cd _r01 && C3_STEPS=12 timeout 300 ncu --set full --clock-control none --launch-count 1 -f -k regex:kProposeStaged --launch-skip 6 -o ../$O/r02_o_r01_staged python scripts/prof_c3.py > /dev/null 2>&1; ncu -i ../$O/r02_o_r01_staged.ncu-rep --page raw --csv > ../$O/r02_o_r01_staged.raw.csv; rm -f ../$O/r02_o_r01_staged.ncu-rep)
C3_STEPS=12 timeout 300 ncu --set full --clock-control none --launch-count 1 -f -k regex:kProposeStaged --launch-skip 6 -o $O/r02_o_r02_staged python scripts/prof_c3.py > /dev/null 2>&1; ncu -i $O/r02_o_r02_staged.ncu-rep --page raw --csv > $O/r02_o_r02_staged.raw.csv; rm -f $O/r02_o_r02_staged.ncu-rep
python scripts/ncu_kernel_summary.py $O/r02_o_r01_staged.raw.csv kProposeStaged | cut -c1-500
python scripts/ncu_kernel_summary.py $O/r02_o_r02_staged.raw.csv kProposeStaged | cut -c1-500
