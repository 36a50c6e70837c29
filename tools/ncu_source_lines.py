#!/usr/bin/env python
"""Executed instructions and stall samples per SOURCE LINE of an ncu report captured with --import-source on
(read here, without a GPU):  python tools/ncu_source_lines.py rep.ncu-rep [top N]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
cur = fn = hdr = None
agg, samp, src = collections.Counter(), collections.Counter(), {}
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        fn = r[1]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None:
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    d = dict(zip(hdr, r))
    key = (fn, cur, ln)
    def num(v):
        try:
            return int(v)
        except (TypeError, ValueError):
            return 0
    agg[key] += num(d.get("Instructions Executed"))
    samp[key] += num(d.get("# Samples"))
    src[key] = r[1]
for f in sorted({k[0] for k in agg}):
    tot = sum(v for k, v in agg.items() if k[0] == f); ts = sum(v for k, v in samp.items() if k[0] == f)
    print(f"== {f}: {tot} warp instructions, {ts} stall samples")
    for k, v in sorted(((k, v) for k, v in agg.items() if k[0] == f), key=lambda kv: -kv[1])[:top]:
        print(f"  {k[1]}:{k[2]:<5d} inst {100.0 * v / max(tot, 1):5.1f}%  samples {100.0 * samp[k] / max(ts, 1):5.1f}%  | {src[k].strip()[:120]}")
