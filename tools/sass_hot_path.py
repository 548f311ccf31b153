#!/usr/bin/env python3
"""Static instruction count of the integrate kernels' step loop on its hot path (rare blocks — out-of-line calls, the
negative-voxel recount — are skipped), from the SASS of the built object. No GPU needed.
usage: tools/sass_hot_path.py <substring of the mangled kernel name> [full]   e.g. integrate_kernel_r1ILb1ELb0ELb1ELb1ELi4"""
import re, signal, subprocess, sys
signal.signal(signal.SIGPIPE, signal.SIG_DFL)      # `| head` closes the pipe early
import os
obj = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "voxel-hashing-sdf_b200", "build", "vh_integrate.o")
syms = sorted(set(re.findall(r"_ZN2vh[0-9A-Za-z_]*", subprocess.run(["cuobjdump", "-elf", obj], capture_output=True, text=True).stdout)))
name = [x for x in syms if sys.argv[1] in x][0]
sass = subprocess.run(["cuobjdump", "-sass", "-fun", name, obj], capture_output=True, text=True).stdout
ins = []
for l in sass.splitlines():
    m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
addr2i = {a: i for i, (a, _) in enumerate(ins)}
# the step loop: last STG.E.128 -> next backward conditional branch
last_stg = max(i for i, (_, t) in enumerate(ins) if "STG.E.128" in t)
end = next(i for i in range(last_stg, len(ins)) if re.search(r"BRA\s+(P\d, )?0x", ins[i][1]) and int(re.search(r"0x([0-9a-f]+)", ins[i][1]).group(1), 16) < ins[i][0])
start = addr2i[int(re.search(r"0x([0-9a-f]+)", ins[end][1]).group(1), 16)]
i, hot, skipped = start, [], 0
while i <= end:
    a, t = ins[i]
    m = re.search(r"BRA\s+(?:!?P\d, )?0x([0-9a-f]+)", t)
    if m and t.startswith("@"):
        tgt = addr2i.get(int(m.group(1), 16))
        if tgt is not None and i < tgt <= end:
            region = [x for _, x in ins[i + 1:tgt]]
            big = any("STG.E.128" in x or "LDG.E.128" in x or "CCTL" in x for x in region)
            rare = (any("CALL" in x for x in region) and not big) or (sum(x.startswith("FSETP.GEU.AND") and "RZ" in x for x in region) >= 6 and not big)
            if rare:
                hot.append(t)
                skipped += tgt - i - 1
                i = tgt
                continue
    hot.append(t)
    i += 1
from collections import Counter
ops = Counter((x.split()[1] if x.startswith("@") else x.split()[0]).split(".")[0] for x in hot)
print(f"{name[:60]}...: loop {end - start + 1} static, {len(hot)} on the hot path ({skipped} in rare blocks)")
print("  " + ", ".join(f"{k} {v}" for k, v in ops.most_common(24)))
if len(sys.argv) > 2:
    for x in hot: print("   ", x)
