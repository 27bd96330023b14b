#!/usr/bin/env python3
"""Per-kernel SASS opcode counts of the shipped library (evidence that the hot kernels are what DESIGN.md says:
tcgen05 MMA = UTCHMMA, TMEM loads = LDTM, TMA = UTMALDG, 128-bit / 64-bit streaming loads = LDG.E.128 / LDG.E.64, ...).
  python profiles/sass_opcodes.py [vectorlite_b200/libvectorlite_cuda.so] > profiles/r02_sass_opcodes.txt
Runs on the CPU (cuobjdump only)."""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "vectorlite_b200/libvectorlite_cuda.so"
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMAPF", "SYNCS", "LDGSTS", "LDG.E.128", "LDG.E.64",
         "LDG.E", "STG.E", "ATOMG", "RED.E", "REDUX", "SHFL", "DADD", "DMUL", "DFMA", "FFMA", "FADD", "HMMA", "BAR.SYNC",
         "ACQBULK", "ERRBAR", "MEMBAR", "CCTL", "ELECT", "FMNMX3"]
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
demangle = {}
names = re.findall(r"Function : (\S+)", out)
if names:
    d = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    demangle = dict(zip(names, d))
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.search(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        if op.startswith("LDG."):   # size modifiers may follow cache hints (LDG.E.EF.64.CONSTANT ...)
            counts[cur]["LDG.E.128" if ".128" in op else "LDG.E.64" if ".64" in op else "LDG.E"] += 1
            continue
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                counts[cur][w] += 1
                break
print(f"# SASS opcode counts per kernel of {so} (cuobjdump -sass, sm_100a); instructions, not dynamic executions")
print(f"# columns: total instructions | " + "watched opcodes with a non-zero count")
for fn, c in sorted(counts.items(), key=lambda kv: demangle.get(kv[0], kv[0])):
    name = demangle.get(fn, fn)
    name = re.sub(r"\(.*", "", name)[:110]
    if not total[fn]:
        continue
    print(f"{name}\n    {total[fn]:6d} | " + "  ".join(f"{k}={v}" for k, v in c.items()))
agg = collections.Counter()
for c in counts.values():
    agg.update(c)
print("# library totals: " + "  ".join(f"{k}={agg[k]}" for k in WATCH if agg[k]))
