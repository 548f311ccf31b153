#!/usr/bin/env bash
# the collective vh_reset with its closing barrier: the two sharded modes that use it (gpurun --gpus 2)
TAG="${1:-r02p}"; OUT=gpurun_out; mkdir -p $OUT
VH_MULTI_ONLY="frames in flight" timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_worker.py > $OUT/multi_worker_${TAG}_n2.log 2>&1; echo "worker rc=$?" >> $OUT/multi_worker_${TAG}_n2.log
grep -E "ok \[|MULTI_GPU_OK|rc=|Error|error" $OUT/multi_worker_${TAG}_n2.log | head -6
