#!/usr/bin/env bash
# Last single-GPU call of a round: what the driver runs (whole GPU suite, smoke, default bench) plus memcheck of the smoke scene, an A/B of
# the marching-cubes colour tile, the ncu launch list and one full capture of the default integrate kernel.   usage: tools/gpu_final.sh [tag]
TAG="${1:-r02final}"; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi > $OUT/nvidia_smi_$TAG.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --durations=10 > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_$TAG.log
tail -16 $OUT/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke_$TAG.log; tail -2 $OUT/smoke_$TAG.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/sanitizer_memcheck_$TAG.log 2>&1; echo "memcheck rc=$?" >> $OUT/sanitizer_memcheck_$TAG.log; tail -2 $OUT/sanitizer_memcheck_$TAG.log
VH_MC_COLOR_TILE=1 VH_ALLOC_REV=2 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/sanitizer_memcheck_alt_$TAG.log 2>&1; echo "memcheck rc=$?" >> $OUT/sanitizer_memcheck_alt_$TAG.log; tail -2 $OUT/sanitizer_memcheck_alt_$TAG.log
run() { local label="$1"; shift; local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 900 python bench.py --no-cpu-baseline --no-c4 --no-ref-cuda "$@" > $OUT/bench_${TAG}_$label.log 2>&1
  python - <<PY
import json
for l in open("$OUT/bench_${TAG}_$label.log"):
    if l.startswith("{"):
        d=json.loads(l); print("$label", round(d["value"]), round(d["e2e"]["value"]), round(d["e2e"].get("async_value") or 0), {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["per_frame"].items() if k.startswith("ms_")}, round(d["roofline"]["frac"],3), {k: round(v,3) for k,v in (d["roofline"].get("frac_per_frame") or {}).items()})
PY
}
run c2_mc_gathers VH_MC_COLOR_TILE=0 VH_BENCH_DUMP=$OUT/per_frame_c2_$TAG.csv -- --steps 10 --warmup 3
run c2_mc_colour_tile VH_MC_COLOR_TILE=1 -- --steps 10 --warmup 3
run c4_mc_colour_tile VH_MC_COLOR_TILE=1 -- --config C4 --steps 4 --warmup 1
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $OUT/bench_$TAG.log 2>&1; echo "bench rc=$?" >> $OUT/bench_$TAG.log
python - <<PY
import json
for l in open("$OUT/bench_$TAG.log"):
    if l.startswith("{"):
        d=json.loads(l); print("default", round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"],3), d["roofline"].get("traffic"), d["roofline"].get("traffic_launch_algorithmic_bytes"), "room_scale", {k: (round(v,1) if isinstance(v,float) else v) for k,v in (d.get("room_scale") or {}).items() if k in ("frames_per_sec","voxel_updates_per_sec","error")}, "ref_cuda", {k: ({kk: vv for kk, vv in v.items() if kk in ("reference_frames_per_sec","ours_frames_per_sec","speedup","error","unavailable")} if isinstance(v, dict) else v) for k,v in (d.get("ref_cuda_baseline") or {}).items() if k != "config2"}, "cpu", (d.get("cpu_baseline") or {}).get("value"), "gpu_launches", d.get("gpu_launches"))
PY
grep -E "^real|rc=" $OUT/bench_$TAG.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 420 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-c4 --no-ref-cuda > $OUT/bench_under_ncu_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:integrate_kernel_staged -s 120 -c 2 -f -o $OUT/prof_integrate_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-c4 --no-ref-cuda > $OUT/ncu_integrate_$TAG.log 2>&1
