#!/usr/bin/env bash
# BASELINE configs 1 and 3 on one GPU with the shipped defaults
TAG="${1:-r02o}"; OUT=gpurun_out; mkdir -p $OUT
for C in C3 C1; do
  timeout 100 python bench.py --no-cpu-baseline --no-c4 --no-ref-cuda --config $C --steps 2 --warmup 1 > $OUT/bench_${TAG}_$C.log 2>&1
  python - <<PY
import json
for l in open("$OUT/bench_${TAG}_$C.log"):
    if l.startswith("{"):
        d=json.loads(l); print("$C", round(d["value"]), round(d["e2e"]["value"]), round(d["e2e"].get("async_value") or 0), {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["per_frame"].items() if k.startswith("ms_")}, round(d["roofline"]["frac"],3))
PY
done
