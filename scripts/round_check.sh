#!/bin/bash
# what the driver runs at round end: gpu tests, smoke, reference arm, bench
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/gputests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
