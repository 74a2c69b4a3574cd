"""dev tool: per-kernel SASS summary of libmergenet_b200.so (cuobjdump -sass): instruction count and the mnemonics
that show how a kernel touches memory (TMA bulk copies, mbarriers, vector widths, atomics, barriers, fp64).
usage: python tools/sass_summary.py [lib.so] > profiles/rNN_sass_summary.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "mergenet_b200/libmergenet_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEEP = re.compile(r"^(UBLKCP|UTMA|SYNCS|LDG|STG|LDS|STS|ATOM|ATOMS|ATOMG|RED|BAR|DADD|DMUL|DFMA|MUFU|LDGSTS|CCTL|MEMBAR|ERRBAR|FENCE|WARPSYNC|SHFL|VOTE|MATCH|REDUX|LDL|STL|BSSY)")
cur = None
hist = collections.OrderedDict()
first = {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        first[cur] = {}
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if cur and m:
        op = m.group(1)
        hist[cur]["_total"] += 1
        if KEEP.match(op):
            hist[cur][op] += 1
            first[cur].setdefault(op.split(".")[0], line.strip()[:150])
own = [k for k in hist if "mn_" in k and "cub" not in k]
print("SASS summary of", lib, "(sm_100a; cuobjdump -sass); own kernels only\n")
for k in own:
    h = hist[k]
    name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip().split("(")[0]
    print("== %s: %d instructions" % (name, h["_total"]))
    groups = collections.Counter()
    for op, n in h.items():
        if op != "_total":
            groups[op] += n
    print("   " + ", ".join("%s %d" % (op, n) for op, n in sorted(groups.items(), key=lambda kv: (-kv[1], kv[0]))[:28]))
    for g in ("UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDG", "STG", "ATOMG", "BAR"):
        if g in first[k]:
            print("   e.g. " + first[k][g])
    print()
