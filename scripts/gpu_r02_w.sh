#!/bin/bash
# C3 per chain: kProposeStaged at 8 CTAs per SM (64 registers)
mkdir -p gpurun_out; O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_staged.py tests/test_gpu_chains.py -q -x -p no:cacheprovider > $O/w_pytest.log 2>&1; tail -3 $O/w_pytest.log
for r in 1 2; do
timeout 300 python bench.py --config c3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('per-chain ms/step %.4f frac %.3f'%(d['ms_per_step'], r['frac']), ' pooled ms/step %.4f'%r['pooled']['ms_per_step'], 'e2e %.4g'%d['e2e']['value'])"
done
