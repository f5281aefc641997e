#!/usr/bin/env python
"""profiles/r02_sass_loops.txt: the inner loops of kFakePairs, kDummyContractDmma and kHmcLeapDmma as
cuobjdump -sass shows them in the shipped library.   python scripts/sass_excerpts.py > profiles/r02_sass_loops.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "root-simple-mcmc_b200", "smcmc_b200", "libsmcmc_b200.so")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout.split("\n")


def fn_ins(name):
    st = [i for i, l in enumerate(out) if "Function : " + name in l and l.strip().endswith(name)][0]
    ins = []
    for l in out[st + 1:]:
        if "Function :" in l:
            break
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def loops(ins, op):
    res = []
    for a, t in ins:
        if "BRA" in t:
            m = re.search(r"0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a:
                tgt = int(m.group(1), 16)
                body = [x for x in ins if tgt <= x[0] <= a]
                if any(op in x[1] for x in body):
                    res.append(body)
    return res


print("Inner loops of the shipped library (cuobjdump -sass libsmcmc_b200.so, sm_100a), extracted by")
print("scripts/sass_excerpts.py.  Addresses are offsets inside the kernel.\n")
ins = fn_ins("_ZN5smcmc10kFakePairsENS_10PairLaunchE")
lp = sorted(loops(ins, "MUFU.EX2"), key=len)[0]
print("== smcmc::kFakePairs, the loop over a tile's events where no separation test is needed (8 events per iteration).")
print("   LDS.128 = four events' worth of one field, broadcast; FMUL2 / FFMA2 / FADD2 = packed FP32x2; MUFU.EX2 = the two")
print("   exponentials per pair; FFMA2.RZ = floor(q(1+m)), floor(q(1-m)); VIMNMX + IMAD + ATOMS = the provisional count;")
print("   LOP3 = hi ^ lo accumulated; the branch at the end of the hot part skips the block that settles undecided pairs.\n")
hot = []
for a, t in lp:
    hot.append((a, t))
    if re.match(r"@!?P\d\s+BRA", t) and len(hot) > 60:
        break
for a, t in hot:
    print("    /*%04x*/  %s" % (a, t))
print("    ... (%d instructions: take the provisional counts of the undecided events back, queue them) ..." % (len(lp) - len(hot) - 4))
for a, t in lp[-4:]:
    print("    /*%04x*/  %s" % (a, t))
print("\n   hot path: %d instructions per 8 events = %.2f per event\n" % (len(hot) + 4, (len(hot) + 4) / 8.0))
for name, label in (("_ZN5smcmc18kDummyContractDmmaILb1EEEvPKdS2_PdPKiiiii", "smcmc::kDummyContractDmma<true>"),
                    ("_ZN5smcmc12kHmcLeapDmmaILb1ELb0EEEvPKdNS_9LeapFusedEiii", "smcmc::kHmcLeapDmma<true,false>")):
    ins = fn_ins(name)
    lp = sorted(loops(ins, "DMMA"), key=len)[0]
    h = collections.Counter(re.sub(r"^@!?U?P\d\s+", "", t).split()[0].split(".")[0] for a, t in lp)
    print("== %s, the K loop (one step of 16: DMMA = mma.sync.m8n8k4.f64, LDGSTS = 16-byte cp.async of the step after next):" % label)
    print("   %d instructions: %s\n" % (len(lp), dict(h.most_common())))
    for a, t in lp[:44]:
        print("    /*%04x*/  %s" % (a, t))
    print("    ... (%d more) ...\n" % (len(lp) - 44))
