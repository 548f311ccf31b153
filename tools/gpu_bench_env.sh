#!/usr/bin/env bash
# bench once per environment setting. usage: tools/gpu_bench_env.sh <tag> "ENV1=a ENV2=b" "ENV1=c" ...
TAG="$1"; shift; OUT=gpurun_out; mkdir -p $OUT; n=0
for envs in "$@"; do
  n=$((n+1))
  env $envs timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_${TAG}_$n.log 2>&1
  python - <<PY
import json
for l in open("$OUT/bench_${TAG}_$n.log"):
    if l.startswith("{"):
        d=json.loads(l); print("$envs", round(d["value"]), round(d["e2e"]["value"]), {k:round(v,4) for k,v in d["per_frame"].items() if k.startswith("ms_")}, round(d["roofline"]["frac"],3))
PY
done
