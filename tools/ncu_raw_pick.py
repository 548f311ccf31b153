#!/usr/bin/env python
"""Pick metrics (regex) out of `ncu --page raw --csv`, one block per profiled launch."""
import csv, re, sys
pat = re.compile(sys.argv[1])
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("--", r[hdr.index("Kernel Name")][:90], " grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
    for i, h in enumerate(hdr):
        if pat.fullmatch(h):
            print(f"   {h:70s} {r[i]:>18s} {units[i]}")
