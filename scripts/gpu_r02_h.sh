#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider -k "pooled or hmc or diagnostics" > $O/r02_h_pytest.log 2>&1; tail -6 $O/r02_h_pytest.log
timeout 300 python scripts/pooled_bench.py > $O/r02_h_pooled_large.txt 2>&1; cat $O/r02_h_pooled_large.txt
timeout 300 python scripts/hmc_ab.py > $O/r02_h_hmc_ab.txt 2>&1; cat $O/r02_h_hmc_ab.txt
POOLED_STEPS=34 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_h_launches_pooled_large.csv python scripts/prof_pooled_large.py > /dev/null 2>&1
python scripts/summarize_launches.py $O/r02_h_launches_pooled_large.csv | head -14
timeout 600 python bench.py --config c4 --no-cpu-baseline > $O/r02_h_bench_c4.json 2> /dev/null; head -c 300 $O/r02_h_bench_c4.json; echo
timeout 600 python bench.py --config c3 --no-cpu-baseline > $O/r02_h_bench_c3.json 2> /dev/null; head -c 300 $O/r02_h_bench_c3.json; echo
