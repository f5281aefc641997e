#!/bin/bash
# round 2, call C: the fused leap-frog stage + 3-stage DMMA pipeline, staged draw variants
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider -k "hmc or pooled or baseline_shapes or staged or chains or user_functor" > $O/r02_c_pytest.log 2>&1; tail -8 $O/r02_c_pytest.log
timeout 600 python bench.py --config c4 --no-cpu-baseline > $O/r02_c_bench_c4.json 2> $O/r02_c_bench_c4.err; tail -2 $O/r02_c_bench_c4.err; head -c 300 $O/r02_c_bench_c4.json; echo
SMCMC_HMC_NO_FUSE=1 timeout 600 python bench.py --config c4 --no-cpu-baseline > $O/r02_c_bench_c4_nofuse.json 2> /dev/null; head -c 300 $O/r02_c_bench_c4_nofuse.json; echo
for d in 0 1 2; do
  SMCMC_STAGED_DRAW=$d timeout 300 python bench.py --config c3 --no-cpu-baseline > $O/r02_c_bench_c3_draw$d.json 2>/dev/null; head -c 260 $O/r02_c_bench_c3_draw$d.json; echo
done
timeout 600 python scripts/configs_bench.py c3 > $O/r02_c_configs_c3.jsonl 2>&1; cut -c1-330 $O/r02_c_configs_c3.jsonl
HMC_STEPS=6 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_c_launches_hmc.csv python scripts/prof_hmc.py > /dev/null 2>&1
HMC_STEPS=4 timeout 300 ncu --set full --import-source on --clock-control none --launch-count 1 -f -k regex:kHmcLeapDmma --launch-skip 20 -o $O/r02_c_kHmcLeapDmma python scripts/prof_hmc.py > $O/r02_c_ncu_leap.log 2>&1; tail -1 $O/r02_c_ncu_leap.log
