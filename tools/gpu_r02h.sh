#!/usr/bin/env bash
# Round-2 single-GPU call: staged integrate with half-block work items at several launch shapes, changed parity tests, reference-PLY test,
# the driver's default bench command (all side keys on), ncu of the staged kernel.
TAG="${1:-r02h}"; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_reference_ply.py tests/test_gpu_parity.py tests/test_gpu_integrate_kernels.py tests/test_gpu_alloc_kernels.py -m gpu -q -k "not direct" --durations=5 > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_$TAG.log
tail -9 $OUT/pytest_gpu_$TAG.log
run() { local label="$1"; shift; local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 900 python bench.py --no-cpu-baseline --no-c4 --no-ref-cuda "$@" > $OUT/bench_${TAG}_$label.log 2>&1
  python - <<PY
import json
for l in open("$OUT/bench_${TAG}_$label.log"):
    if l.startswith("{"):
        d=json.loads(l); print("$label", round(d["value"]), round(d["e2e"]["value"]), round(d["e2e"].get("async_value") or 0), {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["per_frame"].items() if k.startswith("ms_")}, round(d["roofline"]["frac"],3), {k: round(v,3) for k,v in (d["roofline"].get("frac_per_frame") or {}).items()})
PY
}
run c2_ns2_m4 VH_BENCH_DUMP=$OUT/per_frame_c2_$TAG.csv -- --steps 10 --warmup 3
run c2_ns2_m5 VH_INTEGRATE_CTAS=5 -- --steps 10 --warmup 3
run c2_ns1_m6 VH_INTEGRATE_CTAS=6 VH_INTEGRATE_TWO_STEPS=0 -- --steps 10 --warmup 3
run c2_direct VH_INTEGRATE_REV=1 -- --steps 10 --warmup 3
C4="--config C4 --steps 4 --warmup 1"
run c4_ns2_m4 -- $C4
run c4_ns1_m6 VH_INTEGRATE_CTAS=6 VH_INTEGRATE_TWO_STEPS=0 -- $C4
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $OUT/bench_default_$TAG.log 2>&1; echo "bench rc=$?" >> $OUT/bench_default_$TAG.log
python - <<PY
import json
for l in open("$OUT/bench_default_$TAG.log"):
    if l.startswith("{"):
        d=json.loads(l); print("default", round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"],3), "room_scale", {k: (round(v,1) if isinstance(v,float) else v) for k,v in (d.get("room_scale") or {}).items() if k in ("frames_per_sec","voxel_updates_per_sec","error")}, "ref_cuda", {k: ({kk: vv for kk, vv in v.items() if kk in ("reference_frames_per_sec","ours_frames_per_sec","speedup","faces_reference","faces_ours","error","unavailable")} if isinstance(v, dict) else v) for k,v in (d.get("ref_cuda_baseline") or {}).items() if k != "config2"}, "cpu", d.get("cpu_baseline"))
PY
tail -4 $OUT/bench_default_$TAG.log | grep -E "real|rc="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:integrate_kernel_staged -s 120 -c 2 -f -o $OUT/prof_staged_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-c4 --no-ref-cuda > $OUT/ncu_staged_$TAG.log 2>&1
