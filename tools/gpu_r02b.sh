#!/usr/bin/env bash
# Round-2 second GPU call: candidate defaults on the headline config (per-frame dump), the other BASELINE configs on one GPU
# (config 4 with the ray step cap scaled to 1100), ncu of the allocation kernels and of integrate_kernel_r1 at 3 CTAs/SM.
TAG="${1:-r02b}"; OUT=gpurun_out; mkdir -p $OUT
NEW="VH_INTEGRATE_REV=1 VH_INTEGRATE_CTAS=3 VH_MC_REV=1 VH_MC_FILTER_CTAS=8 VH_STATUS_PUBLISH=1"
run() { local label="$1"; shift; local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 900 python bench.py --no-cpu-baseline "$@" > $OUT/bench_${TAG}_$label.log 2>&1
  python - <<PY
import json
for l in open("$OUT/bench_${TAG}_$label.log"):
    if l.startswith("{"):
        d=json.loads(l); print("$label", round(d["value"]), round(d["e2e"]["value"]), round(d["e2e"].get("async_value") or 0), {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["per_frame"].items()}, round(d["roofline"]["frac"],3), d["roofline"].get("frac_per_frame"))
PY
}
run c2_new $NEW VH_BENCH_DUMP=$OUT/per_frame_c2_$TAG.csv -- --steps 10 --warmup 3
run c2_new_allocr1 $NEW VH_ALLOC_REV=1 -- --steps 10 --warmup 3
run c4_old -- --config C4 --ray-steps 1100 --pool-blocks 8388608 --steps 2 --warmup 1
run c4_new $NEW VH_BENCH_DUMP=$OUT/per_frame_c4_$TAG.csv -- --config C4 --ray-steps 1100 --pool-blocks 8388608 --steps 2 --warmup 1
run c4_new_allocr1 $NEW VH_ALLOC_REV=1 -- --config C4 --ray-steps 1100 --pool-blocks 8388608 --steps 2 --warmup 1
run c3_new $NEW -- --config C3 --steps 2 --warmup 1
run c1_new $NEW -- --config C1 --steps 2 --warmup 1
env $NEW timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/bench_under_ncu_$TAG.log 2>&1
env $NEW timeout 600 ncu --set full --clock-control none --import-source on -k regex:integrate_kernel_r1 -s 120 -c 2 -f -o $OUT/prof_integrate_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_integrate_$TAG.log 2>&1
env $NEW VH_ALLOC_REV=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:alloc_visible_kernel_r1 -s 120 -c 2 -f -o $OUT/prof_allocr1_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_allocr1_$TAG.log 2>&1
env $NEW timeout 600 ncu --set full --clock-control none --import-source on -k regex:alloc_visible_kernel -s 120 -c 2 -f -o $OUT/prof_alloc_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_alloc_$TAG.log 2>&1
env $NEW timeout 600 ncu --set full --clock-control none --import-source on -k regex:mc_filter_kernel -s 120 -c 2 -f -o $OUT/prof_mcfilter_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_mcfilter_$TAG.log 2>&1
