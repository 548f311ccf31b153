#!/usr/bin/env bash
# A/B of the marching-cubes mesh kernel's colour path on configs 2 and 4 (after the pop-ahead was reverted)
TAG="${1:-r02k}"; OUT=gpurun_out; mkdir -p $OUT
run() { local label="$1"; shift; local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 900 python bench.py --no-cpu-baseline --no-c4 --no-ref-cuda "$@" > $OUT/bench_${TAG}_$label.log 2>&1
  python - <<PY
import json
for l in open("$OUT/bench_${TAG}_$label.log"):
    if l.startswith("{"):
        d=json.loads(l); print("$label", round(d["value"]), round(d["e2e"]["value"]), round(d["e2e"].get("async_value") or 0), {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["per_frame"].items() if k.startswith("ms_")}, round(d["roofline"]["frac"],3))
PY
}
run c2_mc_gathers VH_MC_COLOR_TILE=0 -- --steps 10 --warmup 3
run c2_mc_colour_tile VH_MC_COLOR_TILE=1 -- --steps 10 --warmup 3
run c4_mc_gathers VH_MC_COLOR_TILE=0 -- --config C4 --steps 4 --warmup 1
run c4_mc_colour_tile VH_MC_COLOR_TILE=1 -- --config C4 --steps 4 --warmup 1
timeout 300 python -m pytest tests/test_gpu_mc_kernels.py -m gpu -q > $OUT/pytest_mc_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_mc_$TAG.log; tail -2 $OUT/pytest_mc_$TAG.log
