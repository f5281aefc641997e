#!/usr/bin/env python
"""Key figures of one kernel out of `ncu -i X.ncu-rep --page raw --csv`:
    python scripts/ncu_kernel_summary.py raw.csv <kernel name substring> [label]
Prints a JSON object (duration, executed instructions, pipe utilisations, DRAM
bytes, occupancy limits, registers)."""
import csv
import json
import sys

KEYS = {
    "gpu__time_duration.sum": "duration_ms",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "pipe_fp64_pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "pipe_xu_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active": "pipe_dmma_pct",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active": "inst_dmma_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_wavefronts_pct",
}
SCALE = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3,
         "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6, "second": 1e3}


def main(raw, needle, label=None):
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = []
    for r in data:
        name = r[hdr.index("Kernel Name")]
        if needle not in name:
            continue
        rec = {"kernel": name.split("(")[0]}
        for i, h in enumerate(hdr):
            if h in KEYS and r[i] not in ("", "n/a"):
                v = float(r[i].replace(",", ""))
                if units[i] in SCALE:
                    v *= SCALE[units[i]]
                rec[KEYS[h]] = v
        out.append(rec)
    if not out:
        raise SystemExit("no launch of %s in %s" % (needle, raw))
    best = max(out, key=lambda x: x.get("duration_ms", 0.0))       # the full-size launch
    if label:
        best["workload"] = label
    print(json.dumps(best))


if __name__ == "__main__":
    main(*sys.argv[1:])
