#!/usr/bin/env bash
# Multi-GPU pass (gpurun --gpus N): the sharded-map parity test, then bench.py under torchrun on the headline config and on the
# room-scale config 4 (2 mm voxels, ray step cap scaled to 1100 so that the rays reach the walls), each with the shipped
# kernels and with the opt-in revisions.   usage: tools/gpu_multi_c4.sh <N> [tag]
N="${1:-2}"; TAG="${2:-r02multi}"; OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_multi.py -q > $OUT/pytest_multi_${TAG}_n$N.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_multi_${TAG}_n$N.log
tail -3 $OUT/pytest_multi_${TAG}_n$N.log
run() {   # label, env..., then bench args after --
  local label="$1"; shift; local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 6 --warmup 3 --no-cpu-baseline "$@" > $OUT/bench_${TAG}_n${N}_$label.log 2>&1
  python - <<PY
import json
for l in open("$OUT/bench_${TAG}_n${N}_$label.log"):
    if l.startswith("{"):
        d = json.loads(l); s = d.get("sharded", {})
        print("$label n=$N", "independent maps:", round(d["value"]), "frames/s; sharded:", {k: (round(v.get("frames_per_sec", 0)) if isinstance(v, dict) else v) for k, v in s.items()})
PY
}
REVS="VH_INTEGRATE_REV=1 VH_ALLOC_REV=1 VH_MC_REV=1"
run c2_default VH_INTEGRATE_REV=0 --
run c2_revs $REVS --
run c4_default VH_INTEGRATE_REV=0 -- --config C4 --ray-steps 1100 --pool-blocks 8388608
run c4_revs $REVS -- --config C4 --ray-steps 1100 --pool-blocks 8388608
