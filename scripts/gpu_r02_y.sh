#!/bin/bash
# C4: kHmcBegin with the momenta loaded ahead of the draws and the kinetic-energy sums one lane per chain
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_hmc.py tests/test_gpu_baseline_shapes.py tests/test_cpp_facade.py -q -x -p no:cacheprovider > $O/y_pytest.log 2>&1; tail -3 $O/y_pytest.log
timeout 600 python bench.py --config c4 --no-cpu-baseline > $O/y_bench_c4.json 2> $O/y_bench_c4.err; tail -2 $O/y_bench_c4.err
python -c "
import json; d=json.load(open('$O/y_bench_c4.json')); r=d['roofline']
print('steady: ms/step %.3f steps/s %.4g L %.1f frac %.3f evals/s %.4g launches %d'%(d['ms_per_step'], d['value'], d['mean_trajectory_length'], r['frac'], d['likelihood_evals_per_s'], d['gpu_launches']))
print('transient:', d['tuning_transient'])"
timeout 300 python scripts/hmc_ab.py 2>&1 | tail -6
HMC_STEPS=6 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/y_launches_hmc.csv python scripts/prof_hmc.py > /dev/null 2>&1
python scripts/summarize_launches.py $O/y_launches_hmc.csv 2>/dev/null | head -9
