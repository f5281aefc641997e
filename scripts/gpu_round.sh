#!/bin/bash
# round check: gpu tests, smoke, both bench arms, the ncu launch list of the bench command,
# the per-config figures (C1, C3, C4, example2) and the C3 launch list
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/gputests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
python scripts/configs_bench.py c1 c3 c4 ex2 2>&1 | tee gpurun_out/configs.jsonl
C3_STEPS=12 python scripts/prof_c3.py > gpurun_out/c3_plain.log 2>&1 && \
C3_STEPS=12 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c3_launches.csv \
    python scripts/prof_c3.py > gpurun_out/ncu_c3.log 2>&1
tail -2 gpurun_out/ncu_c3.log
