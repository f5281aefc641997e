#!/bin/bash
# staged proposal kernel: parity tests, C3 timing (staged vs global-memory kernel), ncu capture
mkdir -p gpurun_out
python -m pytest tests/test_gpu_staged.py tests/test_gpu_chains.py -x -q 2>&1 | tail -15 | tee gpurun_out/staged_tests.log
python scripts/configs_bench.py c3 2>&1 | tee gpurun_out/c3_staged.jsonl
SMCMC_PROPOSE_GENERIC=1 python scripts/configs_bench.py c3 2>&1 | tee gpurun_out/c3_generic.jsonl
timeout 600 ncu --set full --import-source on --clock-control none -k regex:kProposeStaged --launch-skip 35 --launch-count 1 \
  -o gpurun_out/prof_propose_staged -f python scripts/prof_c3.py > gpurun_out/ncu_staged.log 2>&1
tail -3 gpurun_out/ncu_staged.log
