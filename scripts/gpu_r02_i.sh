#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > $O/r02_i_pytest.log 2>&1; tail -8 $O/r02_i_pytest.log
timeout 300 python scripts/pooled_bench.py > $O/r02_i_pooled_large.txt 2>&1; cat $O/r02_i_pooled_large.txt
POOLED_STEPS=34 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_i_launches_pooled_large.csv python scripts/prof_pooled_large.py > /dev/null 2>&1
python scripts/summarize_launches.py $O/r02_i_launches_pooled_large.csv | head -12
timeout 300 python scripts/hmc_ab.py > $O/r02_i_hmc_ab.txt 2>&1; cat $O/r02_i_hmc_ab.txt
