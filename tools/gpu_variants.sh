#!/usr/bin/env bash
# parity tests once, then the bench for each VH_INTEGRATE_CTAS variant. usage: tools/gpu_variants.sh <tag>
TAG="${1:-v}"; OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_$TAG.log
tail -4 $OUT/pytest_gpu_$TAG.log
for c in 3 2 4; do
  VH_INTEGRATE_CTAS=$c timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_${TAG}_c$c.log 2>&1
  python - <<PY
import json
for l in open("$OUT/bench_${TAG}_c$c.log"):
    if l.startswith("{"):
        d=json.loads(l); print("ctas=$c", round(d["value"]), round(d["e2e"]["value"]), {k:round(v,4) for k,v in d["per_frame"].items() if k.startswith("ms_")}, round(d["roofline"]["frac"],3))
PY
done
