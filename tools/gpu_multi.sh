#!/usr/bin/env bash
# Multi-GPU pass (gpurun --gpus N): the sharded-map parity test (all modes), then bench.py under torchrun: the room-scale config 4
# (default at N > 1) and the headline config 2, sharded.   usage: tools/gpu_multi.sh <N> [tag]
N="${1:-2}"; TAG="${2:-r02multi}"; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo_${TAG}_n$N.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -q -k "$N" > $OUT/pytest_multi_${TAG}_n$N.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_multi_${TAG}_n$N.log
tail -12 $OUT/pytest_multi_${TAG}_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_worker.py > $OUT/multi_worker_${TAG}_n$N.log 2>&1; echo "worker rc=$?" >> $OUT/multi_worker_${TAG}_n$N.log
grep -E "ok \[|MULTI_GPU_OK|rc=|Error|error" $OUT/multi_worker_${TAG}_n$N.log | head -20
run() {   # label, env..., then bench args after --
  local label="$1"; shift; local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N "$@" > $OUT/bench_${TAG}_n${N}_$label.log 2>&1
  python - <<PY
import json
ok = False
for l in open("$OUT/bench_${TAG}_n${N}_$label.log"):
    if l.startswith("{"):
        ok = True
        d = json.loads(l)
        sg = (d.get("room_scale") or d.get("curve") or {}).get("single_gpu_same_run") or {}
        print("$label n=$N value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "Mupd/s", round(d["voxel_updates_per_sec"] / 1e6), "| single GPU same run:",
              round(sg.get("frames_per_sec", 0), 1), sg.get("per_frame"), "| last frame rank0", d["sharded"]["last_frame_this_rank"],
              "| c2 sharded", round((d.get("headline_c2_sharded") or {}).get("frames_per_sec", 0)), "| config5", (d.get("config5_independent_maps") or {}))
if not ok:
    print("$label n=$N: no JSON line"); print(open("$OUT/bench_${TAG}_n${N}_$label.log").read()[-1500:])
PY
}
run c4 -- --steps 4 --warmup 1
run c4_nomc -- --steps 4 --warmup 1 --no-mc
run c4_nccl VH_SHARD_BCAST=nccl -- --steps 4 --warmup 1
run c2 -- --config C2 --steps 4 --warmup 1
