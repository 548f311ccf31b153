#!/usr/bin/env python
"""Per-kernel totals and shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
agg = collections.defaultdict(list)
for r in rows:
    agg[r[4].split("(")[0]].append(float(r[-1]))
tot = sum(sum(v) for v in agg.values())
print(f"{len(rows)} launches, {tot / 1e3:.1f} us of device time (cold-cache, serialised: compare shares, not absolutes)")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:60s} n={len(v):4d} mean={sum(v) / len(v) / 1000:9.2f} us  share={sum(v) / tot * 100:5.1f}%")
