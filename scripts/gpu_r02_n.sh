#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider -k "fake or small or stream or fullsize or pooled or hmc or unbinned" > $O/r02_n_pytest.log 2>&1; tail -6 $O/r02_n_pytest.log
timeout 300 python scripts/configs_bench.py stream > $O/r02_n_stream.jsonl 2>&1; cut -c1-200 $O/r02_n_stream.jsonl
timeout 300 python scripts/hmc_ab.py > $O/r02_n_hmc_ab.txt 2>&1; cat $O/r02_n_hmc_ab.txt
timeout 300 python scripts/pooled_bench.py 2>&1 | tail -2
HMC_STEPS=4 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02_n_launches_hmc.csv python scripts/prof_hmc.py > /dev/null 2>&1
python scripts/summarize_launches.py $O/r02_n_launches_hmc.csv | head -9
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02_n_launches_stream.csv python scripts/prof_stream.py > /dev/null 2>&1
python scripts/summarize_launches.py $O/r02_n_launches_stream.csv | grep -v "cub::\|Gather\|SortKeys\|CountClasses\|PadEvents\|InitState\|StoreStart" | head -8
