#!/bin/bash
# C4: the first gradient of a step taken from the previous step (kHmcLeapCached)
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_hmc.py tests/test_gpu_baseline_shapes.py tests/test_cpp_facade.py -q -x -p no:cacheprovider > $O/y_pytest.log 2>&1; tail -5 $O/y_pytest.log
for v in 0 1; do
  [ $v = 1 ] && export SMCMC_HMC_NO_GRADIENT_CACHE=1
  echo "== SMCMC_HMC_NO_GRADIENT_CACHE=$v"
  timeout 600 python bench.py --config c4 --no-cpu-baseline > $O/y_bench_c4_$v.json 2> $O/y_bench_c4_$v.err; tail -2 $O/y_bench_c4_$v.err
  python -c "
import json; d=json.load(open('$O/y_bench_c4_$v.json')); r=d['roofline']
print('steady: ms/step %.3f steps/s %.4g L %.1f frac %.3f executed %.3f evals/s %.4g launches %d'%(d['ms_per_step'], d['value'], d['mean_trajectory_length'], r['frac'], r['executed']['frac'], d['likelihood_evals_per_s'], d['gpu_launches']))
print('transient:', d['tuning_transient'])"
  timeout 300 python scripts/hmc_ab.py 2>&1 | grep "rep [02] fused"
done
