#!/usr/bin/env bash
# First GPU call for the opt-in kernel revisions (integrate_kernel_r1, alloc_visible_kernel_r1: so far checked under CPU emulation only):
# the gated parity tests, the standard suite with VH_STATUS_PUBLISH=1, then the bench once per switch setting, then one
# ncu --set full capture of integrate_kernel_r1.   usage (under gpurun, from the repo root): tools/gpu_rev1.sh [tag]
TAG="${1:-r02rev1}"; OUT=gpurun_out; mkdir -p $OUT
VH_TEST_REV1=1 timeout 2400 python -m pytest tests/test_gpu_integrate_rev1.py tests/test_gpu_alloc_rev1.py tests/test_gpu_mc_rev1.py tests/test_gpu_configs34.py tests/test_gpu_stream.py -q > $OUT/pytest_rev1_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_rev1_$TAG.log
tail -4 $OUT/pytest_rev1_$TAG.log
# the whole standard GPU suite once more with the status block published by a kernel (every test goes through that path)
VH_STATUS_PUBLISH=1 timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_publish_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_publish_$TAG.log
tail -3 $OUT/pytest_publish_$TAG.log
tools/gpu_bench_env.sh $TAG "VH_INTEGRATE_REV=0" "VH_INTEGRATE_REV=1" "VH_INTEGRATE_REV=1 VH_INTEGRATE_CTAS=3" "VH_INTEGRATE_REV=1 VH_INTEGRATE_CTAS=7" "VH_INTEGRATE_REV=0 VH_STATUS_PUBLISH=1" "VH_INTEGRATE_REV=1 VH_STATUS_PUBLISH=1" "VH_ALLOC_REV=1" "VH_MC_REV=1" "VH_MC_FILTER_CTAS=8" "VH_MC_FILTER_CTAS=8 VH_MC_MESH_CTAS=6" "VH_INTEGRATE_REV=1 VH_ALLOC_REV=1 VH_MC_REV=1 VH_STATUS_PUBLISH=1 VH_MC_FILTER_CTAS=8"
VH_INTEGRATE_REV=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:integrate_kernel_r1 -s 120 -c 2 -f -o $OUT/prof_integrate_r1_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_integrate_r1_$TAG.log 2>&1
