#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python scripts/hmc_ab.py > $O/r02_g_hmc_ab.txt 2>&1; cat $O/r02_g_hmc_ab.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02_g_launches_pooled_large.csv python scripts/prof_pooled_large.py > /dev/null 2>&1
python scripts/summarize_launches.py $O/r02_g_launches_pooled_large.csv | head -14
