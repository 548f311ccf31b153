#!/usr/bin/env bash
# full ncu capture of one kernel inside a short bench run. usage: tools/gpu_ncu_kernel.sh <tag> <kernel-regex> [env assignments...]
TAG="$1"; K="$2"; shift 2
OUT=gpurun_out; mkdir -p $OUT
env "$@" timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 120 -c 2 -f -o $OUT/prof_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_$TAG.log 2>&1
tail -2 $OUT/ncu_$TAG.log | cut -c1-300
