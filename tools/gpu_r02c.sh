#!/usr/bin/env bash
# Round-2 third GPU call: integrate_kernel_r2 (bulk-copy staging): parity tests, memcheck, bench sweep, ncu; config 4 on one GPU
# (16 frames at the 1100-step ray cap: the map grows ~0.9 M blocks per frame, 100 frames do not fit one GPU).
TAG="${1:-r02c}"; OUT=gpurun_out; mkdir -p $OUT
VH_TEST_REV1=1 timeout 900 python -m pytest tests/test_gpu_integrate_rev1.py -q -k "r2" --durations=5 > $OUT/pytest_r2_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_r2_$TAG.log
tail -4 $OUT/pytest_r2_$TAG.log
VH_INTEGRATE_REV=2 VH_INTEGRATE_TWO_STEPS=1 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/sanitizer_memcheck_r2_$TAG.log 2>&1; echo "memcheck rc=$?" >> $OUT/sanitizer_memcheck_r2_$TAG.log
tail -3 $OUT/sanitizer_memcheck_r2_$TAG.log
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
run c2_r2_4_ns2 $BASE VH_INTEGRATE_REV=2 VH_INTEGRATE_TWO_STEPS=1 VH_BENCH_DUMP=$OUT/per_frame_c2_r2_$TAG.csv -- --steps 10 --warmup 3
run c2_r2_4_ns1 $BASE VH_INTEGRATE_REV=2 VH_INTEGRATE_TWO_STEPS=0 -- --steps 10 --warmup 3
run c2_r2_3_ns2 $BASE VH_INTEGRATE_REV=2 VH_INTEGRATE_TWO_STEPS=1 VH_INTEGRATE_CTAS=3 -- --steps 10 --warmup 3
run c2_r1_3 $BASE VH_INTEGRATE_REV=1 VH_INTEGRATE_CTAS=3 -- --steps 10 --warmup 3
C4="--config C4 --ray-steps 1100 --pool-blocks 16777216 --frames-per-step 4 --steps 4 --warmup 1"
run c4_r1 $BASE VH_INTEGRATE_REV=1 VH_INTEGRATE_CTAS=3 VH_BENCH_DUMP=$OUT/per_frame_c4_$TAG.csv -- $C4
run c4_r1_allocr1 $BASE VH_INTEGRATE_REV=1 VH_INTEGRATE_CTAS=3 VH_ALLOC_REV=1 -- $C4
run c4_r2_allocr1 $BASE VH_INTEGRATE_REV=2 VH_INTEGRATE_TWO_STEPS=1 VH_ALLOC_REV=1 -- $C4
env $BASE VH_INTEGRATE_REV=2 VH_INTEGRATE_TWO_STEPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:integrate_kernel_r2 -s 120 -c 2 -f -o $OUT/prof_integrate_r2_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_integrate_r2_$TAG.log 2>&1
env $BASE timeout 600 ncu --set full --clock-control none --import-source on -k regex:mc_mesh_kernel -s 120 -c 2 -f -o $OUT/prof_mcmesh_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_mcmesh_$TAG.log 2>&1
