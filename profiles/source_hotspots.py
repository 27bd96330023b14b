#!/usr/bin/env python3
"""SASS-level view of an ncu report captured with --import-source on (runs on the CPU: `ncu -i … --page source`).
Prints, for one kernel of the report, the executed-instruction and stall-sample totals per opcode and the most
sampled SASS lines with their dominant stall reasons — the view behind the r02 findings (the GPU-scope fence in front of
a remote mbarrier arrive, the epilogue's share of the first filtered stage, warp 0's pool maintenance in the HNSW kernel).
  python profiles/source_hotspots.py gpurun_out/x.ncu-rep [kernel_index=-1] [first_line last_line]
ncu lists every kernel twice on this page; index 0, 2, 4 … are the distinct launches in capture order."""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else -1
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
blocks, cur = [], None
for line in out.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = [line]
        blocks.append(cur)
    elif cur is not None:
        cur.append(line)
b = blocks[which]
rows = list(csv.DictReader(io.StringIO("\n".join(b[1:]))))
tot_inst = sum(int(r["Instructions Executed"]) for r in rows)
tot_samp = sum(int(r["# Samples"]) for r in rows)
print(b[0][:140])
print("instructions", tot_inst, "samples", tot_samp)
ops, samp = collections.Counter(), collections.Counter()
for r in rows:
    src = r["Source"].split()
    op = src[1] if src and src[0].startswith("@") else (src[0] if src else "?")
    ops[op] += int(r["Instructions Executed"])
    samp[op] += int(r["# Samples"])
print("-- by opcode (instructions executed, stall samples)")
for op, c in ops.most_common(25):
    print(f"{op:30s} {c:12d} {100 * c / max(tot_inst, 1):5.1f}%   samples {samp[op]:8d} {100 * samp[op] / max(tot_samp, 1):5.1f}%")


def stalls(r, n):
    st = {k[6:]: int(v) for k, v in r.items() if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0")}
    return sorted(st.items(), key=lambda x: -x[1])[:n]


print("-- most sampled lines")
for i, r in sorted(enumerate(rows), key=lambda x: -int(x[1]["# Samples"]))[:30]:
    print(i, r["Source"].strip()[:72].ljust(72), r["# Samples"].rjust(6), r["Instructions Executed"].rjust(10), stalls(r, 3))
if len(sys.argv) > 4:
    print("-- lines", sys.argv[3], "…", sys.argv[4])
    for i in range(int(sys.argv[3]), int(sys.argv[4])):
        r = rows[i]
        print(i, r["Source"].strip()[:90].ljust(90), r["# Samples"].rjust(6), r["Instructions Executed"].rjust(10), stalls(r, 2))
