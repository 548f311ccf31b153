#!/usr/bin/env bash
# last check of the shipped binary (gpurun --gpus 2): six sharded modes, the two-kernel allocation's single-GPU tests, smoke
TAG="${1:-r02m}"; OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_worker.py > $OUT/multi_worker_${TAG}_n2.log 2>&1; echo "worker rc=$?" >> $OUT/multi_worker_${TAG}_n2.log
grep -E "ok \[|MULTI_GPU_OK|rc=|Error|error" $OUT/multi_worker_${TAG}_n2.log | head -12
CUDA_VISIBLE_DEVICES=0 timeout 300 python -m pytest tests/test_gpu_alloc_kernels.py tests/test_gpu_configs34.py -m gpu -q -k "keys or config4" > $OUT/pytest_alloc_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_alloc_$TAG.log; tail -2 $OUT/pytest_alloc_$TAG.log
CUDA_VISIBLE_DEVICES=0 timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke_$TAG.log; tail -2 $OUT/smoke_$TAG.log
