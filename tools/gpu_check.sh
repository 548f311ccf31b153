#!/usr/bin/env bash
# One GPU-box pass: parity tests, smoke, bench, ncu launch list and full captures of the two hot kernels.
# Usage (from the repo root, under gpurun): tools/gpu_check.sh [tag]
TAG="${1:-r01}"
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi > $OUT/nvidia_smi_$TAG.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke_$TAG.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.log 2>&1; echo "bench rc=$?" >> $OUT/bench_$TAG.log
# memory and race checks of every kernel on the small smoke scene (the reference has none)
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/sanitizer_memcheck_$TAG.log 2>&1; echo "memcheck rc=$?" >> $OUT/sanitizer_memcheck_$TAG.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/sanitizer_racecheck_$TAG.log 2>&1; echo "racecheck rc=$?" >> $OUT/sanitizer_racecheck_$TAG.log
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > $OUT/bench_ref_$TAG.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/bench_under_ncu_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:integrate_kernel -s 120 -c 2 -f -o $OUT/prof_integrate_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_integrate_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mc_mesh_kernel -s 120 -c 2 -f -o $OUT/prof_mc_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_mc_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mc_filter_kernel -s 120 -c 2 -f -o $OUT/prof_mcfilter_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_mcfilter_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:alloc_visible_kernel -s 120 -c 2 -f -o $OUT/prof_alloc_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_alloc_$TAG.log 2>&1
tail -n 3 $OUT/pytest_gpu_$TAG.log; tail -n 2 $OUT/smoke_$TAG.log; tail -c 600 $OUT/bench_$TAG.log
