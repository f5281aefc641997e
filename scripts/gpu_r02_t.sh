#!/bin/bash
# source-level ncu capture of one kernel: K=<kernel regex> P=<profiling script> [env for the script]
mkdir -p gpurun_out; O=gpurun_out
NCU="ncu --set full --import-source on --clock-control none --launch-count 1 -f"
timeout 300 $NCU -k regex:$K --launch-skip ${SKIP:-6} -o $O/t_$K python scripts/$P > $O/t_ncu_$K.log 2>&1; tail -1 $O/t_ncu_$K.log
ncu -i $O/t_$K.ncu-rep --page raw --csv > $O/t_$K.raw.csv 2>/dev/null
ncu -i $O/t_$K.ncu-rep --page source --csv > $O/t_$K.source.csv 2>/dev/null
rm -f $O/t_$K.ncu-rep
