#!/bin/bash
# end of round 2, eight GPUs, the tree as committed: the bench line at N = 8 with its multi_gpu leg, and C5 at its own size
# (262 144 chains x 16.8 M events) chain-sharded (8 x 1), mixed (4 x 2) and event-sharded (1 x 8)
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29811 bench.py --gpus 8 --steps 10 --warmup 3 > $O/r02_z_bench_8gpu.json 2> $O/r02_z_bench_8gpu.err; tail -3 $O/r02_z_bench_8gpu.err; head -c 300 $O/r02_z_bench_8gpu.json; echo
rm -f $O/r02_z_c5_8gpu.jsonl
for eg in 1 8 2; do
  timeout 900 $TR --master-port 2982$eg scripts/configs_bench.py c5 --event-group $eg --steps 3 2>/dev/null | grep '^{' >> $O/r02_z_c5_8gpu.jsonl
done
cut -c1-420 $O/r02_z_c5_8gpu.jsonl
