#!/usr/bin/env python
"""Static per-opcode instruction counts of one kernel in a cuobjdump -sass listing, optionally
restricted to an address range: sass_count.py file.sass <mangled-name substring> [lo hi]."""
import re, sys, collections
txt = open(sys.argv[1]).read()
pat = sys.argv[2]
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
funcs = re.split(r"\n\s*Function : ", txt)
for f in funcs[1:]:
    name = f.split("\n", 1)[0]
    if pat not in name: continue
    ops = collections.Counter(); n = 0
    for m in re.finditer(r"/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", f):
        addr = int(m.group(1), 16)
        if addr < lo or addr >= hi: continue
        op = m.group(3).split(".")[0]
        if m.group(3).startswith("IMAD.WIDE"): op = "IMAD.WIDE"
        if m.group(3).startswith("IMAD.HI"): op = "IMAD.HI"
        if m.group(3).startswith("IMAD.MOV"): op = "IMAD.MOV"
        if m.group(3).startswith("FFMA.SAT"): op = "FFMA.SAT"
        ops[op] += 1; n += 1
    print(name[:100], "total", n)
    for op, c in ops.most_common(): print(f"  {op:12s} {c:6d}  {c/8:7.2f}/genotype")
