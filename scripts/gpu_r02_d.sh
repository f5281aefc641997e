#!/bin/bash
# round 2, call D: fused leap-frog stage with the row-coalesced epilogue; streaming step launch list
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider -k "hmc or baseline_shapes or chains or resident" > $O/r02_d_pytest.log 2>&1; tail -6 $O/r02_d_pytest.log
timeout 600 python bench.py --config c4 --no-cpu-baseline > $O/r02_d_bench_c4.json 2> $O/r02_d_bench_c4.err; tail -2 $O/r02_d_bench_c4.err; head -c 300 $O/r02_d_bench_c4.json; echo
SMCMC_HMC_NO_FUSE=1 timeout 600 python bench.py --config c4 --no-cpu-baseline > $O/r02_d_bench_c4_nofuse.json 2> /dev/null; head -c 300 $O/r02_d_bench_c4_nofuse.json; echo
HMC_STEPS=6 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_d_launches_hmc.csv python scripts/prof_hmc.py > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_d_launches_stream.csv python scripts/prof_stream.py > /dev/null 2>&1
HMC_STEPS=4 timeout 300 ncu --set full --import-source on --clock-control none --launch-count 1 -f -k regex:kHmcLeapDmma --launch-skip 20 -o $O/r02_d_kHmcLeapDmma python scripts/prof_hmc.py > $O/r02_d_ncu_leap.log 2>&1; tail -1 $O/r02_d_ncu_leap.log
