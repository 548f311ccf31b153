#!/usr/bin/env bash
# Round-2 first GPU call: the parity tests that had only run under CPU emulation (kernel revisions, BASELINE configs 3/4,
# out-of-core tier), the standard suite with the status block published by a kernel, the bench once per switch, one ncu
# capture of integrate_kernel_r1.   usage (under gpurun, repo root): tools/gpu_r02a.sh [tag]
TAG="${1:-r02a}"; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi > $OUT/nvidia_smi_$TAG.txt 2>&1
VH_TEST_REV1=1 timeout 1500 python -m pytest tests/test_gpu_integrate_rev1.py tests/test_gpu_alloc_rev1.py tests/test_gpu_mc_rev1.py tests/test_gpu_configs34.py tests/test_gpu_stream.py -q --durations=20 > $OUT/pytest_rev1_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_rev1_$TAG.log
tail -6 $OUT/pytest_rev1_$TAG.log
VH_STATUS_PUBLISH=1 timeout 900 python -m pytest tests -m gpu -q --durations=10 > $OUT/pytest_publish_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_publish_$TAG.log
tail -3 $OUT/pytest_publish_$TAG.log
tools/gpu_bench_env.sh $TAG "VH_INTEGRATE_REV=0" "VH_INTEGRATE_REV=1" "VH_INTEGRATE_REV=1 VH_INTEGRATE_CTAS=3" "VH_INTEGRATE_REV=1 VH_INTEGRATE_CTAS=7" "VH_STATUS_PUBLISH=1" "VH_ALLOC_REV=1" "VH_MC_REV=1" "VH_MC_FILTER_CTAS=8" "VH_INTEGRATE_REV=1 VH_ALLOC_REV=1 VH_MC_REV=1 VH_STATUS_PUBLISH=1"
VH_INTEGRATE_REV=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:integrate_kernel_r1 -s 120 -c 2 -f -o $OUT/prof_integrate_r1_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_integrate_r1_$TAG.log 2>&1
