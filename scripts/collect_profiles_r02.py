#!/usr/bin/env python
"""gpurun_out/<tag>_* of scripts/gpu_r02_final.sh -> the tracked profiles/r02_* files.

    python scripts/collect_profiles_r02.py r02_final
"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02_final"
G = os.path.join(ROOT, "gpurun_out", tag + "_")
P = os.path.join(ROOT, "profiles")

plain = {"bench_c1.json": "r02_bench_c1.json", "bench_c2.json": "r02_bench_c2.json", "bench_c3.json": "r02_bench_c3.json",
         "bench_c4.json": "r02_bench_c4.json", "bench_reference_arm.json": "r02_bench_reference_arm.json",
         "launches.csv": "r02_launches.csv", "launches.txt": "r02_launches.txt", "hmc_ab.txt": "r02_hmc_ab.txt",
         "pooled_large.txt": "r02_pooled_large.txt", "stream.jsonl": "r02_stream.jsonl", "ab_r01.txt": "r02_ab_r01.txt",
         "pytest.log": "r02_pytest_gpu.log", "smoke.log": "r02_smoke.log"}
for src, dst in plain.items():
    if os.path.exists(G + src):
        shutil.copyfile(G + src, os.path.join(P, dst))
    else:
        sys.stderr.write("missing %s\n" % (G + src))
for name in ("stream", "c3_pooled", "hmc"):
    src = G + "launches_%s.csv" % name
    if os.path.exists(src):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "summarize_launches.py"), src],
                             capture_output=True, text=True).stdout
        open(os.path.join(P, "r02_launches_%s.txt" % name), "w").write(out)

kernels = ["kFakePairs", "kFakeStream", "kProposeStaged", "kProposePooledTile", "kHmcLeapDmma", "kDummyContractDmma", "kPoolGramDmma"]
out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_reports_to_jsonl.py"), G] + kernels,
                     capture_output=True, text=True)
sys.stderr.write(out.stderr)
# a kernel that was not captured this time (kDummyContractDmma no longer runs inside an HMC step: the potential
# comes out of the last gradient launch) keeps its previous summary
have = {json.loads(l)["kernel"].split("<")[0].split()[-1] for l in out.stdout.splitlines() if l.strip()}
kept = ""
try:
    for l in open(os.path.join(P, "r02_kernels.jsonl")):
        if l.strip() and json.loads(l)["kernel"].split("<")[0].split()[-1] not in have:
            kept += l
except OSError:
    pass
open(os.path.join(P, "r02_kernels.jsonl"), "w").write(out.stdout + kept)

# the pair kernel's profile, in the form bench.py reads
rows = list(csv.reader(open(G + "kFakePairs.raw.csv")))
d = dict(zip(rows[0], rows[2]))
f = lambda k: float(d[k].replace(",", ""))
stalls = {h[33:]: f(h) for h in rows[0] if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")
          and d[h] not in ("", "n/a")}
tot = sum(stalls.values())
E, N = 4096, 1000000
prof = {
    "kernel": "smcmc::kFakePairs", "launches_profiled": 1, "workload": {"chains": E, "events": N},
    "report": "gpurun_out/%s_kFakePairs.ncu-rep (ncu --set full --clock-control none, round 2, end of round)" % tag,
    "dram_bytes_per_launch": f("dram__bytes_read.sum") + f("dram__bytes_write.sum"),
    "warp_instructions_per_launch": f("smsp__inst_executed.sum"),
    "warp_instructions_per_pair_warp": f("smsp__inst_executed.sum") / (E * N / 32.0),
    "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "warps_eligible_per_scheduler": f("smsp__warps_eligible.avg.per_cycle_active"),
    "duration_ms_under_ncu": f("gpu__time_duration.sum"),
    "registers_per_thread": f("launch__registers_per_thread"),
    "occupancy_limit_ctas_per_sm": {"registers": f("launch__occupancy_limit_registers"),
                                    "shared_memory": f("launch__occupancy_limit_shared_mem")},
    "pipes_pct": {k: f("sm__inst_executed_pipe_%s.avg.pct_of_peak_sustained_active" % k) for k in ("alu", "fma", "fp64", "lsu", "xu")},
    "top_stalls_pct": {k: round(100 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:10]},
    "stall_note": "smsp__pcsamp_warps_issue_stalled_* of the launch.  By instruction (source page of the same report): "
                  "about 46 % of all samples sit on MUFU.EX2, half of them mio_throttle -- MUFU, LDS and the shared-memory "
                  "atomics leave the sub-partition through one in-order queue; see DESIGN.md 4.1 for the take-out experiments",
    "grid": f("launch__grid_size"), "block": f("launch__block_size"),
    "dyn_smem": f("launch__shared_mem_per_block_dynamic"),
}
units = dict(zip(rows[0], rows[1]))
if units.get("gpu__time_duration.sum", "").startswith("us"):
    prof["duration_ms_under_ncu"] /= 1000.0
json.dump(prof, open(os.path.join(P, "pair_kernel_profile.json"), "w"), indent=1)
log = G + "ncu_pairs.log"
if os.path.exists(log):
    for line in open(log):
        if line.startswith("filter check"):
            open(os.path.join(P, "r02_filter_check.txt"), "w").write(line)
print("profiles refreshed from", G)
