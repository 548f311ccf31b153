#!/usr/bin/env bash
# Round-2 fifth GPU call: new defaults (staged integrate, CTA-aggregated work list, merged MC revisions, status published by a kernel);
# allocation revision 2 (ray_keys_kernel + insert_keys_kernel) against revision 0: parity tests, memcheck, bench on configs 2 and 4.
TAG="${1:-r02e}"; OUT=gpurun_out; mkdir -p $OUT
VH_ALLOC_REV=2 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs34.py tests/test_gpu_stream.py tests/test_gpu_hashmap.py -m gpu -q --durations=5 > $OUT/pytest_alloc2_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_alloc2_$TAG.log
tail -5 $OUT/pytest_alloc2_$TAG.log
VH_ALLOC_REV=2 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/sanitizer_memcheck_alloc2_$TAG.log 2>&1; echo "memcheck rc=$?" >> $OUT/sanitizer_memcheck_alloc2_$TAG.log
tail -3 $OUT/sanitizer_memcheck_alloc2_$TAG.log
VH_ALLOC_REV=2 timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/sanitizer_racecheck_alloc2_$TAG.log 2>&1; echo "racecheck rc=$?" >> $OUT/sanitizer_racecheck_alloc2_$TAG.log
tail -3 $OUT/sanitizer_racecheck_alloc2_$TAG.log
run() { local label="$1"; shift; local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 900 python bench.py --no-cpu-baseline "$@" > $OUT/bench_${TAG}_$label.log 2>&1
  python - <<PY
import json
for l in open("$OUT/bench_${TAG}_$label.log"):
    if l.startswith("{"):
        d=json.loads(l); print("$label", round(d["value"]), round(d["e2e"]["value"]), round(d["e2e"].get("async_value") or 0), {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["per_frame"].items()}, round(d["roofline"]["frac"],3), d["roofline"].get("frac_per_frame"))
PY
}
run c2_alloc0 VH_ALLOC_REV=0 VH_BENCH_DUMP=$OUT/per_frame_c2_$TAG.csv -- --steps 10 --warmup 3
run c2_alloc2 VH_ALLOC_REV=2 -- --steps 10 --warmup 3
run c2_alloc2_direct VH_ALLOC_REV=2 VH_INTEGRATE_REV=1 -- --steps 10 --warmup 3
C4="--config C4 --ray-steps 1100 --pool-blocks 16777216 --frames-per-step 4 --steps 4 --warmup 1"
run c4_alloc0 VH_ALLOC_REV=0 -- $C4
run c4_alloc2 VH_ALLOC_REV=2 -- $C4
run c3_alloc2 VH_ALLOC_REV=2 -- --config C3 --steps 2 --warmup 1
run c1_alloc2 VH_ALLOC_REV=2 -- --config C1 --steps 2 --warmup 1
VH_ALLOC_REV=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 420 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/bench_under_ncu_$TAG.log 2>&1
for k in ray_keys_kernel insert_keys_kernel mc_mesh_kernel; do
VH_ALLOC_REV=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 120 -c 2 -f -o $OUT/prof_${k}_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_${k}_$TAG.log 2>&1
done
