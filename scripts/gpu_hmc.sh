#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_hmc.py -x -q 2>&1 | tail -5 | tee gpurun_out/hmc_tests.log
python scripts/configs_bench.py c4 2>&1 | tee gpurun_out/c4.jsonl
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/gputests.log
