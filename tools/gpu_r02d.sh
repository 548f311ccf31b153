#!/usr/bin/env bash
# Round-2 fourth GPU call: the whole (un-gated) GPU suite with the work-list integrate kernels, then the bench per kernel variant.
TAG="${1:-r02d}"; OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_$TAG.log
tail -14 $OUT/pytest_gpu_$TAG.log
BASE="VH_MC_REV=1 VH_MC_FILTER_CTAS=8 VH_STATUS_PUBLISH=1"
run() { local label="$1"; shift; local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 900 python bench.py --no-cpu-baseline "$@" > $OUT/bench_${TAG}_$label.log 2>&1
  python - <<PY
import json
for l in open("$OUT/bench_${TAG}_$label.log"):
    if l.startswith("{"):
        d=json.loads(l); print("$label", round(d["value"]), round(d["e2e"]["value"]), round(d["e2e"].get("async_value") or 0), {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["per_frame"].items()}, round(d["roofline"]["frac"],3), d["roofline"].get("frac_per_frame"))
PY
}
run c2_direct3 $BASE VH_INTEGRATE_REV=1 VH_BENCH_DUMP=$OUT/per_frame_c2_direct_$TAG.csv -- --steps 10 --warmup 3
run c2_direct4 $BASE VH_INTEGRATE_REV=1 VH_INTEGRATE_CTAS=4 -- --steps 10 --warmup 3
run c2_staged_ns2 $BASE VH_INTEGRATE_REV=2 VH_INTEGRATE_TWO_STEPS=1 VH_BENCH_DUMP=$OUT/per_frame_c2_staged_$TAG.csv -- --steps 10 --warmup 3
run c2_staged_ns1 $BASE VH_INTEGRATE_REV=2 VH_INTEGRATE_TWO_STEPS=0 -- --steps 10 --warmup 3
run c2_staged3_ns2 $BASE VH_INTEGRATE_REV=2 VH_INTEGRATE_TWO_STEPS=1 VH_INTEGRATE_CTAS=3 -- --steps 10 --warmup 3
C4="--config C4 --ray-steps 1100 --pool-blocks 16777216 --frames-per-step 4 --steps 4 --warmup 1"
run c4_direct3 $BASE VH_INTEGRATE_REV=1 -- $C4
run c4_staged $BASE VH_INTEGRATE_REV=2 VH_INTEGRATE_TWO_STEPS=1 -- $C4
env $BASE VH_INTEGRATE_REV=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:integrate_kernel_direct -s 120 -c 2 -f -o $OUT/prof_direct_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_direct_$TAG.log 2>&1
env $BASE VH_INTEGRATE_REV=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:cull_list_kernel -s 120 -c 2 -f -o $OUT/prof_cull_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_cull_$TAG.log 2>&1
