#!/bin/bash
# launch list of the default bench command + ncu --set full of kStepsResident and kFakeFinish
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 > gpurun_out/b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_launch.log 2>&1
NCU="ncu --set full --import-source on --clock-control none --launch-count 1 -f"
python scripts/prof_resident.py 2>&1 | tail -1 | tee gpurun_out/resident.txt && timeout 600 $NCU -k regex:kStepsResident --launch-skip 1 -o gpurun_out/prof_resident python scripts/prof_resident.py > gpurun_out/ncu_resident.log 2>&1; tail -1 gpurun_out/ncu_resident.log
RES_CHAINS=1 RES_STEPS=5000 python scripts/prof_resident.py 2>&1 | tail -1 | tee -a gpurun_out/resident.txt
python scripts/prof_pairs.py > /dev/null 2>&1 && timeout 600 $NCU -k regex:kFakeFinish --launch-skip 3 -o gpurun_out/prof_finish python scripts/prof_pairs.py > gpurun_out/ncu_finish.log 2>&1; tail -1 gpurun_out/ncu_finish.log
ls -la gpurun_out/*.ncu-rep | tail -3
