#!/bin/bash
# same-box A/B of library builds on the headline bench: LIBS="path1 path2 ..." ("default" = the in-tree library)
for v in $LIBS; do
  [ "$v" = default ] && lib="" || lib=$v
  SMCMC_B200_LIB=$lib timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-cpp-tree --no-multi-leg 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$v', 'ms/step %.4f'%d['ms_per_step'], 'kernel ms %.4f'%r['launch_ms'], 'sfu frac %.4f'%r['frac'])"
done
