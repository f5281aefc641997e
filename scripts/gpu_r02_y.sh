#!/bin/bash
# C4: ordered fused stage with its own gather loader; A/B of the fixed-length case against the library before the ordering
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_hmc.py tests/test_gpu_baseline_shapes.py tests/test_cpp_facade.py -q -x -p no:cacheprovider > $O/y_pytest.log 2>&1; tail -3 $O/y_pytest.log
timeout 600 python bench.py --config c4 --no-cpu-baseline > $O/y_bench_c4.json 2> $O/y_bench_c4.err; tail -2 $O/y_bench_c4.err
python -c "
import json; d=json.load(open('$O/y_bench_c4.json')); r=d['roofline']
print('steady: ms/step %.3f steps/s %.4g L %.1f frac %.3f evals/s %.4g launches %d'%(d['ms_per_step'], d['value'], d['mean_trajectory_length'], r['frac'], d['likelihood_evals_per_s'], d['gpu_launches']))
print('transient:', d['tuning_transient'])"
for lib in "" root-simple-mcmc_b200/smcmc_b200/variants/libsmcmc_pre.so ""; do
  echo "== lib: ${lib:-default}"
  SMCMC_B200_LIB=$lib timeout 300 python scripts/hmc_ab.py 2>&1 | grep "rep [02] fused"
done
