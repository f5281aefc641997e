#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider -k "hmc or fake or small or fullsize" > $O/r02_e_pytest.log 2>&1; tail -6 $O/r02_e_pytest.log
timeout 600 python bench.py --config c4 --no-cpu-baseline > $O/r02_e_bench_c4.json 2> $O/r02_e_bench_c4.err; tail -2 $O/r02_e_bench_c4.err; head -c 300 $O/r02_e_bench_c4.json; echo
HMC_STEPS=6 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_e_launches_hmc.csv python scripts/prof_hmc.py > /dev/null 2>&1
timeout 300 python scripts/configs_bench.py stream > $O/r02_e_stream_noprefetch.jsonl 2>&1; cut -c1-400 $O/r02_e_stream_noprefetch.jsonl
SMCMC_STREAM_PREFETCH=1 timeout 300 python scripts/configs_bench.py stream > $O/r02_e_stream_prefetch.jsonl 2>&1; cut -c1-400 $O/r02_e_stream_prefetch.jsonl
timeout 300 python scripts/pooled_bench.py > $O/r02_e_pooled_large.txt 2>&1; cat $O/r02_e_pooled_large.txt
