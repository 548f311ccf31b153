#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: stall-reason totals and the hottest SASS instructions.
usage: ncu -i rep.ncu-rep --page source --csv | python tools/ncu_source_summary.py [top_n]"""
import csv, sys
top = int(sys.argv[1]) if len(sys.argv) > 1 else 25
rows = list(csv.reader(sys.stdin))
# first kernel only: rows[0] = kernel name, rows[1] = header
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = []
for r in rows[2:]:
    if len(r) < len(hdr) or r[0] == "Address" or r[0] == "Kernel Name":
        if r and r[0] == "Kernel Name":
            break
        continue
    body.append(r)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: sum(float(r[ix[s]] or 0) for r in body) for s in stalls}
all_s = sum(tot.values()) or 1
print("kernel:", rows[0][1][:100])
print("instructions executed (warp):", sum(int(r[ix["Instructions Executed"]] or 0) for r in body), " SASS lines:", len(body))
print("stall samples:", ", ".join(f"{k[6:]}={v / all_s * 100:.1f}%" for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v / all_s > 0.01))
key = "# Samples"
body2 = sorted(enumerate(body), key=lambda ir: -float(ir[1][ix[key]] or 0))[:top]
ns = sum(float(r[ix[key]] or 0) for r in body) or 1
for i, r in body2:
    rs = {s: float(r[ix[s]] or 0) for s in stalls}
    dom = max(rs, key=rs.get)
    print(f"{i:5d} {float(r[ix[key]]) / ns * 100:5.1f}%  exec={r[ix['Instructions Executed']]:>9s}  {dom[6:]:<12s} {r[ix['Source']].strip()[:90]}")
