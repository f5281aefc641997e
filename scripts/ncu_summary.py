#!/usr/bin/env python
"""Turn an `ncu --set full` report of the pair kernel into the small JSON that
bench.py reads (profiles/pair_kernel_profile.json) plus a readable summary.

    ncu -i gpurun_out/prof_pairs.ncu-rep --page raw --csv > raw.csv
    python scripts/ncu_summary.py raw.csv profiles/r01_pair_kernel
"""
import csv
import json
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "launch__registers_per_thread": "registers",
    "launch__occupancy_limit_shared_mem": "occupancy_limit_smem_blocks",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "pipe_fp64_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "pipe_xu_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active": "pipe_tma_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "sm__cycles_elapsed.max": "sm_cycles",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "dyn_smem",
}
STALLS = "smsp__average_warps_issue_stalled_"
SCALE = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0,
         "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}


def main(raw, out_prefix, chains, events):
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    launches = []
    for r in data:
        rec = {"kernel": r[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in KEYS:
                v = float(r[i].replace(",", "")) if r[i] not in ("", "n/a") else None
                if v is not None and units[i] in SCALE:
                    v *= SCALE[units[i]]
                rec[KEYS[h]] = v
            elif h.startswith(STALLS) and h.endswith("_per_warp_active.pct"):
                rec.setdefault("stall_pct", {})[h[len(STALLS):-len("_per_warp_active.pct")]] = float(r[i])
        launches.append(rec)
    # the dominant kernel only (the regex of the capture also matches kFakePairsGeneric)
    launches = [l for l in launches if "kFakePairs(" in l["kernel"] or l["kernel"].endswith("kFakePairs")]
    n = len(launches)
    avg = lambda k: sum(l[k] for l in launches) / n
    pair_warps = chains * events / 32.0
    summary = {
        "kernel": "smcmc::kFakePairs", "launches_profiled": n,
        "workload": {"chains": chains, "events": events},
        "dram_bytes_per_launch": avg("dram_read") + avg("dram_write"),
        "warp_instructions_per_launch": avg("warp_instructions"),
        "warp_instructions_per_pair_warp": avg("warp_instructions") / pair_warps,
        "issue_active_pct": avg("issue_active_pct"),
        "warps_active_pct": avg("warps_active_pct"),
        "duration_ms_under_ncu": avg("duration") * 1e3,
        "registers_per_thread": launches[0].get("registers"),
        "pipes_pct": {k[5:-4]: avg(k) for k in launches[0] if k.startswith("pipe_") and launches[0][k] is not None},
        "smem_bank_conflicts_per_wavefront": avg("smem_bank_conflicts") / avg("smem_wavefronts"),
        "top_stalls_pct": dict(sorted(launches[0].get("stall_pct", {}).items(), key=lambda kv: -kv[1])[:6]),
        "grid": launches[0].get("grid"), "block": launches[0].get("block"), "dyn_smem": launches[0].get("dyn_smem"),
    }
    json.dump(summary, open(out_prefix + ".json", "w"), indent=1)
    print(json.dumps(summary, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 4096,
         int(sys.argv[4]) if len(sys.argv) > 4 else 1000000)
