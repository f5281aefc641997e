"""Inner loops of one kernel in a built library: instruction count and mnemonic histogram of every
backward-branch loop that contains a given mnemonic (default MUFU.EX2).
usage: sass_loops.py lib.so mangled_name [print] [--op DMMA] [--max 400]"""
import re, subprocess, sys, collections
args = sys.argv[1:]
op, limit = "MUFU.EX2", 400
if "--op" in args:
    op = args[args.index("--op") + 1]
    del args[args.index("--op"):args.index("--op") + 2]
if "--max" in args:
    limit = int(args[args.index("--max") + 1])
    del args[args.index("--max"):args.index("--max") + 2]
sys.argv = [sys.argv[0]] + args
lib, fn = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(out) if "Function : " + fn in l and l.strip().endswith(fn)][0]
ins = []
for l in out[start + 1:]:
    if "Function :" in l:
        break
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
print(len(ins), "instructions")
for a, t in ins:
    if "BRA" in t:
        m2 = re.search(r"0x([0-9a-f]+)", t)
        if m2 and int(m2.group(1), 16) < a:
            tgt = int(m2.group(1), 16)
            body = [x[1] for x in ins if tgt <= x[0] <= a]
            nm = sum(op in x for x in body)
            if nm and len(body) < limit:
                h = collections.Counter(re.sub(r"^@!?U?P\d\s+", "", x).split()[0].split(".")[0] for x in body)
                print("loop %#x..%#x: %d instructions, %d %s%s" % (tgt, a, len(body), nm, op,
                                                                    " (%.2f per event)" % (len(body) / (nm / 2.0)) if op == "MUFU.EX2" else ""),
                      dict(h.most_common()))
                if len(sys.argv) > 3:
                    print("\n".join("    " + x for x in body))
