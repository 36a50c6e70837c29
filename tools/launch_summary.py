#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections, csv, sys
for f in sys.argv[1:]:
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        v = float(r[vi].replace(',', ''))
        v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(r[ui], 1.0)
        name = r[ki].split('(')[0][:70]
        agg[name][0] += 1; agg[name][1] += v
    tot = sum(t for _, t in agg.values())
    print(f, "total %.3f ms" % tot)
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f"  {k:70s} n={n:4d} total={t:9.3f} ms avg={t/n:8.4f} ms share={100*t/tot:5.1f}%")
