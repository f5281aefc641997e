#!/bin/bash
# kFakePairs with provisional counting (no compare / select per event): parity, timing, ncu
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_fake.py tests/test_gpu_fullsize.py tests/test_gpu_fake2.py tests/test_gpu_small_sizes.py tests/test_gpu_multi.py tests/test_gpu_chains.py -q -x -p no:cacheprovider > $O/q_pytest.log 2>&1; tail -5 $O/q_pytest.log
LIBS="default ${LIBS}" bash scripts/gpu_ab_libs.sh
if [ -n "$NCU_PAIRS" ]; then
NCU="ncu --set full --import-source on --clock-control none --launch-count 1 -f"
timeout 300 $NCU -k regex:^kFakePairs$ --launch-skip 3 -o $O/q_kFakePairs python scripts/prof_pairs.py > $O/q_ncu_pairs.log 2>&1; tail -1 $O/q_ncu_pairs.log
ncu -i $O/q_kFakePairs.ncu-rep --page raw --csv > $O/q_kFakePairs.raw.csv 2>/dev/null
ncu -i $O/q_kFakePairs.ncu-rep --page source --csv > $O/q_kFakePairs.source.csv 2>/dev/null
fi
