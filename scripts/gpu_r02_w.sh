#!/bin/bash
# C3 pooled: kProposePooledTile with division-free row loops and a pointer-walked L.z loop
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_pooled.py tests/test_gpu_baseline_shapes.py tests/test_gpu_multi.py -q -x -p no:cacheprovider > $O/w_pytest.log 2>&1; tail -3 $O/w_pytest.log
timeout 300 python bench.py --config c3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('per-chain ms/step %.4f frac %.3f'%(d['ms_per_step'], r['frac']), ' pooled ms/step %.4f'%r['pooled']['ms_per_step'])"
C3_POOLED=16 C3_STEPS=20 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/w_launches_c3_pooled.csv python scripts/prof_c3.py > /dev/null 2>&1
python scripts/summarize_launches.py $O/w_launches_c3_pooled.csv 2>/dev/null | head -7
