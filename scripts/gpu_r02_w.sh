#!/bin/bash
# C3 per chain: deferred covariance update in kProposeStaged (SMCMC_COV_DEFER=0: every step)
mkdir -p gpurun_out; O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_staged.py tests/test_gpu_chains.py tests/test_gpu_resident.py tests/test_gpu_baseline_shapes.py tests/test_gpu_small_sizes.py tests/test_gpu_graph.py tests/test_gpu_debug_modes.py tests/test_cpp_facade.py tests/test_gpu_accept_local.py -q -x -p no:cacheprovider > $O/w_pytest.log 2>&1; tail -6 $O/w_pytest.log
for k in 16 0 8 32; do
  echo "== SMCMC_COV_DEFER=$k"
  SMCMC_COV_DEFER=$k timeout 300 python bench.py --config c3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('per-chain ms/step %.4f frac %.3f'%(d['ms_per_step'], r['frac']), ' pooled ms/step %.4f'%r['pooled']['ms_per_step'], 'e2e %.4g'%d['e2e']['value'])"
done
