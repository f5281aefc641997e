#!/bin/bash
# 8-GPU check: the bench line at N=8 and the C5 sweeps (chain-sharded and event-sharded)
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$T --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench8.json 2> gpurun_out/err8.log; tail -2 gpurun_out/err8.log; cut -c1-300 gpurun_out/bench8.json
$T --master-port 29522 scripts/configs_bench.py c5 --steps 2 2> gpurun_out/c5a.err | tee gpurun_out/c5_8gpu.jsonl
$T --master-port 29523 scripts/configs_bench.py c5 --steps 2 --event-group 8 2> gpurun_out/c5b.err | tee -a gpurun_out/c5_8gpu.jsonl
