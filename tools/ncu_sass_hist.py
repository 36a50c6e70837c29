#!/usr/bin/env python
"""Per-opcode executed-instruction histogram and stall summary from an ncu report's source page."""
import csv, subprocess, sys, io, collections, re
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); tot = 0
stall = collections.Counter(); samples = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    src = r[ci["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", src)
    op = m.group(2) if m else src[:10]
    try:
        n = int(r[ci["Instructions Executed"]] or 0)
    except ValueError:
        continue                                   # a second kernel's header row
    ops[op] += n; tot += n
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            stall[h] += int(r[ci[h]] or 0)
    samples += int(r[ci["# Samples"]] or 0)
unit = float(sys.argv[2]) if len(sys.argv) > 2 else None   # warp-level genotype rows, to normalise
print(f"total warp instructions executed: {tot:,}")
for op, n in ops.most_common(40):
    extra = f"  {n/unit:8.2f} /row" if unit else ""
    print(f"  {op:14s} {n:15,d} {100*n/tot:6.2f}%{extra}")
print("stall samples:", samples)
for h, n in stall.most_common(12):
    print(f"  {h:28s} {100*n/max(samples,1):6.2f}%")
