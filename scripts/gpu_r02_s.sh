#!/bin/bash
# streaming regime: prepare kernel with its transcendental functions spread over warps, single-CTA finish
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_fake.py tests/test_gpu_fake2.py tests/test_gpu_small_sizes.py tests/test_gpu_graph.py tests/test_gpu_chains.py tests/test_cpp_facade.py -q -x -p no:cacheprovider > $O/s_pytest.log 2>&1; tail -3 $O/s_pytest.log
timeout 300 python scripts/configs_bench.py stream 2>/dev/null | cut -c1-330
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/s_launches_stream.csv python scripts/prof_stream.py > /dev/null 2>&1
python scripts/summarize_launches.py $O/s_launches_stream.csv 2>/dev/null | grep -v "cub::\|Gather\|SortKeys\|CountClasses\|PadEvents\|InitState\|StoreStart" | head -10
