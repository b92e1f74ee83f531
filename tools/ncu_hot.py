#!/usr/bin/env python
"""Read an .ncu-rep (source page) and print where a kernel's time goes: executed warp instructions by opcode, and the SASS
instructions with the most stall samples together with their dominant stall reasons.
    python tools/ncu_hot.py <report.ncu-rep> [kernel regex] [launch index] [top N]"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else "."
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern, "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
isrc, iex, ismp = col["Source"], col["Instructions Executed"], col["# Samples"]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_i = sum(int(r[iex]) for r in data)
tot_s = sum(int(r[ismp]) for r in data)
print(f"{rows[0][1][:80]}: {tot_i / 1e6:.1f} M warp instructions, {tot_s} samples")
ops = Counter()
for r in data:
    t = r[isrc].split()
    op = t[1] if t and t[0].startswith("@") else (t[0] if t else "?")
    ops[op] += int(r[iex])
print("opcodes:", ", ".join(f"{k} {v / 1e6:.1f}M" for k, v in ops.most_common(18)))
tot_st = Counter()
for r in data:
    for h in stalls:
        tot_st[h] += int(r[col[h]] or 0)
print("stall samples:", ", ".join(f"{k[6:]} {v}" for k, v in tot_st.most_common(10)))
print(f"top {top} instructions by samples:")
order = sorted(range(len(data)), key=lambda i: -int(data[i][ismp]))[:top]
for i in sorted(order):
    r = data[i]
    st = sorted(((int(r[col[h]] or 0), h[6:]) for h in stalls), reverse=True)[:3]
    print(f"{i:5d} {r[isrc][:78]:78s} exec {int(r[iex]):9d} smp {int(r[ismp]):6d}  " + " ".join(f"{n}:{c}" for c, n in st if c))
