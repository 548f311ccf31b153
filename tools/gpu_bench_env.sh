#!/usr/bin/env bash
# bench once per environment setting. usage: tools/gpu_bench_env.sh <tag> "ENV1=a ENV2=b" "ENV1=c" ...
# prints: value, e2e (sync), e2e async, e2e u16 async, per-stage ms, roofline frac.  e.g. tools/gpu_bench_env.sh r02a VH_STATUS_PUBLISH=0 VH_STATUS_PUBLISH=1
TAG="$1"; shift; OUT=gpurun_out; mkdir -p $OUT; n=0
for envs in "$@"; do
  n=$((n+1))
  env $envs timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_${TAG}_$n.log 2>&1
  python - <<PY
import json
for l in open("$OUT/bench_${TAG}_$n.log"):
    if l.startswith("{"):
        d=json.loads(l); print("$envs", round(d["value"]), round(d["e2e"]["value"]), round(d["e2e"].get("async_value") or 0), round((d["e2e"].get("u16_async") or {}).get("value") or 0), {k:round(v,4) for k,v in d["per_frame"].items() if k.startswith("ms_")}, round(d["roofline"]["frac"],3))
PY
done
