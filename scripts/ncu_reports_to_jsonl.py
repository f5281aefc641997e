#!/usr/bin/env python
"""ncu --set full reports -> one JSON object per kernel with the figures DESIGN.md quotes and the
warp-stall breakdown (smsp__pcsamp_warps_issue_stalled_*).

    python scripts/ncu_reports_to_jsonl.py gpurun_out/r02_l_ kFakePairs kHmcLeapDmma ... > profiles/r02_kernels.jsonl
(reads <prefix><kernel>.ncu-rep, exports the raw page with `ncu -i ... --page raw --csv`)"""
import csv
import io
import json
import subprocess
import sys

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
from ncu_kernel_summary import KEYS, SCALE  # noqa: E402

EXTRA = {
    "sm__warps_active.avg.per_cycle_active": "warps_active_per_sm",
    "smsp__warps_eligible.avg.per_cycle_active": "warps_eligible_per_scheduler",
    "launch__occupancy_limit_registers": "occupancy_limit_registers_ctas",
    "launch__occupancy_limit_shared_mem": "occupancy_limit_smem_ctas",
    "launch__occupancy_limit_warps": "occupancy_limit_warps_ctas",
    "launch__shared_mem_per_block_dynamic": "dynamic_smem_bytes",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
}


def main(prefix, names):
    for name in names:
        rep = prefix + name + ".ncu-rep"
        try:
            if __import__("os").path.exists(prefix + name + ".raw.csv"):
                raw = open(prefix + name + ".raw.csv").read()
            else:
                raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        except Exception as exc:
            sys.stderr.write("%s: %s\n" % (rep, exc))
            continue
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        dur = hdr.index("gpu__time_duration.sum")
        r = max(data, key=lambda x: float(x[dur].replace(",", "")))
        rec = {"kernel": r[hdr.index("Kernel Name")].split("(")[0], "report": rep.split("/")[-1]}
        stalls = {}
        for i, h in enumerate(hdr):
            if r[i] in ("", "n/a"):
                continue
            if h in KEYS or h in EXTRA:
                v = float(r[i].replace(",", ""))
                if units[i] in SCALE:
                    v *= SCALE[units[i]]
                rec[KEYS.get(h) or EXTRA[h]] = v
            elif h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
                stalls[h.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(r[i].replace(",", ""))
        tot = sum(stalls.values()) or 1.0
        rec["top_stalls_pct"] = {k: round(100.0 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:7]}
        print(json.dumps(rec))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
