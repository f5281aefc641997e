#!/bin/bash
# end of round 2, two GPUs, the tree as committed: the NCCL tests, the new pair-kernel tests, the bench
# line at N = 2 (and its reference arm), the C5 shape chain- vs event-sharded
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_fake.py -m gpu -q -p no:cacheprovider > $O/r02_v_pytest_multi.log 2>&1; tail -5 $O/r02_v_pytest_multi.log
timeout 600 $TR --master-port 29711 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r02_v_bench_2gpu.json 2> $O/r02_v_bench_2gpu.err; tail -3 $O/r02_v_bench_2gpu.err; head -c 400 $O/r02_v_bench_2gpu.json; echo
grep -c "NCCL INFO" $O/r02_v_bench_2gpu.err
rm -f $O/r02_v_c5_2gpu.jsonl
for eg in 1 2; do
  timeout 600 $TR --master-port 2972$eg scripts/configs_bench.py c5 --event-group $eg --chains 131072 --events 8388608 --steps 3 2>/dev/null | grep '^{' >> $O/r02_v_c5_2gpu.jsonl
done
cut -c1-500 $O/r02_v_c5_2gpu.jsonl
