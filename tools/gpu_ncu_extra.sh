#!/usr/bin/env bash
# one ncu --set full pass over the work-list and marching-cubes kernels of the final code (sequence frames 20/21 of the timed leg)
TAG="${1:-r02n}"; OUT=gpurun_out; mkdir -p $OUT
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"cull_list_kernel|mc_mesh_kernel|mc_filter_kernel" -s 360 -c 6 -f -o $OUT/prof_extra_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-c4 --no-ref-cuda > $OUT/ncu_extra_$TAG.log 2>&1; echo "rc=$?"; tail -2 $OUT/ncu_extra_$TAG.log | cut -c1-200
