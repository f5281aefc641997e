#!/bin/bash
# round 2, call A: the whole GPU suite, the bench lines of every config, and the ncu captures
# behind the kernel work of this round (stall breakdown of kFakePairs, the C4 kernels, kProposeStaged)
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > $O/r02_a_pytest.log 2>&1; tail -15 $O/r02_a_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_a_smoke.log 2>&1; tail -2 $O/r02_a_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/r02_a_bench_c2.json 2> $O/r02_a_bench_c2.err; tail -3 $O/r02_a_bench_c2.err; head -c 600 $O/r02_a_bench_c2.json; echo
for c in c3 c4 c1; do
  timeout 600 python bench.py --config $c > $O/r02_a_bench_$c.json 2> $O/r02_a_bench_$c.err; tail -2 $O/r02_a_bench_$c.err; head -c 400 $O/r02_a_bench_$c.json; echo
done
NCU="ncu --set full --import-source on --clock-control none --launch-count 1 -f"
timeout 300 $NCU -k regex:kFakePairs --launch-skip 3 -o $O/r02_a_kFakePairs python scripts/prof_pairs.py > $O/r02_a_ncu_pairs.log 2>&1; tail -1 $O/r02_a_ncu_pairs.log
export HMC_STEPS=18
for k in kDummyContractDmma kHmcExxtFlush kHmcKickDrift kHmcPost; do
  timeout 400 $NCU -k regex:$k --launch-skip 4 -o $O/r02_a_$k python scripts/prof_hmc.py > $O/r02_a_ncu_$k.log 2>&1; tail -1 $O/r02_a_ncu_$k.log
done
C3_STEPS=12 timeout 300 $NCU -k regex:kProposeStaged --launch-skip 6 -o $O/r02_a_kProposeStaged python scripts/prof_c3.py > $O/r02_a_ncu_staged.log 2>&1; tail -1 $O/r02_a_ncu_staged.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_a_launches_hmc.csv python scripts/prof_hmc.py > $O/r02_a_ncu_hmc_list.log 2>&1; tail -1 $O/r02_a_ncu_hmc_list.log
ls -la $O | tail -30
