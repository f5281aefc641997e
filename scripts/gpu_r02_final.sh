#!/bin/bash
# round 2, the state of the tree at the end: whole GPU suite, smoke, every bench line (both arms), the ncu
# launch list of the default bench command, ncu --set full captures of the kernels DESIGN.md quotes
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-r02_final}
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > $O/${T}_pytest.log 2>&1; tail -6 $O/${T}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; tail -1 $O/${T}_smoke.log
timeout 600 python bench.py --impl reference --steps 6 --warmup 1 > $O/${T}_bench_reference_arm.json 2> /dev/null; head -c 200 $O/${T}_bench_reference_arm.json; echo
timeout 600 python bench.py --steps 20 --warmup 3 > $O/${T}_bench_c2.json 2> $O/${T}_bench_c2.err; tail -2 $O/${T}_bench_c2.err; head -c 300 $O/${T}_bench_c2.json; echo
for c in c3 c4 c1; do
  timeout 600 python bench.py --config $c > $O/${T}_bench_$c.json 2> $O/${T}_bench_$c.err; tail -2 $O/${T}_bench_$c.err; head -c 300 $O/${T}_bench_$c.json; echo
done
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-cpp-tree > $O/${T}_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/${T}_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-cpp-tree > $O/${T}_ncu_launch.log 2>&1
python scripts/summarize_launches.py $O/${T}_launches.csv > $O/${T}_launches.txt 2>&1; head -12 $O/${T}_launches.txt
timeout 300 python scripts/hmc_ab.py > $O/${T}_hmc_ab.txt 2>&1; cat $O/${T}_hmc_ab.txt
timeout 300 python scripts/pooled_bench.py > $O/${T}_pooled_large.txt 2>&1; cat $O/${T}_pooled_large.txt
timeout 300 python scripts/configs_bench.py stream > $O/${T}_stream.jsonl 2>&1; cut -c1-200 $O/${T}_stream.jsonl
NCU="ncu --set full --import-source on --clock-control none --launch-count 1 -f"
# gpurun brings back at most 64 MiB: every report is exported to its raw CSV page on the box and
# removed, except the headline kernel's
keep() { ncu -i $O/${T}_$1.ncu-rep --page raw --csv > $O/${T}_$1.raw.csv 2>/dev/null; [ "$1" = kFakePairs ] || rm -f $O/${T}_$1.ncu-rep; }
timeout 300 $NCU -k regex:^kFakePairs$ --launch-skip 3 -o $O/${T}_kFakePairs env PAIRS_FILTER_CHECK=1 python scripts/prof_pairs.py > $O/${T}_ncu_pairs.log 2>&1; tail -1 $O/${T}_ncu_pairs.log; keep kFakePairs
export HMC_STEPS=4
for k in kHmcLeapDmma kDummyContractDmma kPoolGramDmma; do
  timeout 400 $NCU -k regex:$k --launch-skip 3 -o $O/${T}_$k python scripts/prof_hmc.py > $O/${T}_ncu_$k.log 2>&1; tail -1 $O/${T}_ncu_$k.log; keep $k
done
C3_STEPS=12 timeout 300 $NCU -k regex:kProposeStaged --launch-skip 6 -o $O/${T}_kProposeStaged python scripts/prof_c3.py > $O/${T}_ncu_staged.log 2>&1; tail -1 $O/${T}_ncu_staged.log; keep kProposeStaged
C3_POOLED=16 C3_STEPS=12 timeout 300 $NCU -k regex:kProposePooledTile --launch-skip 6 -o $O/${T}_kProposePooledTile python scripts/prof_c3.py > $O/${T}_ncu_pooledtile.log 2>&1; tail -1 $O/${T}_ncu_pooledtile.log; keep kProposePooledTile
timeout 300 $NCU -k regex:kFakeStream --launch-skip 2 -o $O/${T}_kFakeStream python scripts/prof_stream.py > $O/${T}_ncu_stream.log 2>&1; tail -1 $O/${T}_ncu_stream.log; keep kFakeStream
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/${T}_launches_stream.csv python scripts/prof_stream.py > /dev/null 2>&1
python scripts/summarize_launches.py $O/${T}_launches_stream.csv | grep -v "cub::\|Gather\|SortKeys\|CountClasses\|PadEvents\|InitState\|StoreStart" | head -10
C3_POOLED=16 C3_STEPS=20 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/${T}_launches_c3_pooled.csv python scripts/prof_c3.py > /dev/null 2>&1
python scripts/summarize_launches.py $O/${T}_launches_c3_pooled.csv | head -8
HMC_STEPS=6 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_hmc.csv python scripts/prof_hmc.py > /dev/null 2>&1
python scripts/summarize_launches.py $O/${T}_launches_hmc.csv | head -14
# same-box A/B against the round-1 tree (a git worktree at _r01, present only while measuring)
if [ -d _r01 ]; then
  for d in _r01 .; do
    ( cd $d; timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$d', 'C2 ms/step %.4f'%d['ms_per_step'], 'kFakePairs ms %.4f'%r['launch_ms'], 'sfu frac %.4f'%r['frac'], 'e2e %.4g'%d['e2e']['value'])" )
  done > $O/${T}_ab_r01.txt 2>&1; cat $O/${T}_ab_r01.txt
fi
