#!/usr/bin/env bash
# 8-GPU pass (gpurun --gpus 8): sharded-map parity in every mode, then bench.py --gpus N for N = 8 (and 4, 2 when asked): the room-scale config 4
# sharded over the GPUs (+ same-run single GPU, config 2 sharded, config 5 replicas), and the whole 100-frame room on 8 GPUs.
TAG="${1:-r02g}"; OUT=gpurun_out; mkdir -p $OUT; NS="${2:-8}"
nvidia-smi topo -m > $OUT/topo_$TAG.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_worker.py > $OUT/multi_worker_${TAG}_n8.log 2>&1; echo "worker rc=$?" >> $OUT/multi_worker_${TAG}_n8.log
grep -E "ok \[|MULTI_GPU_OK|rc=|Error|error" $OUT/multi_worker_${TAG}_n8.log | head -20
run() {   # label, N, env..., then bench args after --
  local label="$1"; shift; local N="$1"; shift; local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
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
        print("$label n=$N value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "Gupd/s", round(d["voxel_updates_per_sec"] / 1e9, 1), "blocks", d["sharded"]["allocated_blocks_all_shards"], "largest shard", d["sharded"]["allocated_blocks_largest_shard"],
              "| single GPU same run:", round(sg.get("frames_per_sec", 0), 1) if "error" not in sg else sg, "| last frame rank0", {k: round(v, 3) for k, v in d["sharded"]["last_frame_this_rank"].items()},
              "| c2 sharded", (d.get("headline_c2_sharded") or {}).get("frames_per_sec") or d.get("headline_c2_sharded"), "| config5", (d.get("config5_independent_maps") or {}))
if not ok:
    print("$label n=$N: no JSON line"); print(open("$OUT/bench_${TAG}_n${N}_$label.log").read()[-1500:])
PY
}
for N in $NS; do run c4 $N -- --steps 8 --warmup 2; done
run c4_nomc 8 -- --steps 8 --warmup 2 --no-mc
run c4_full100 8 -- --all-frames --pool-blocks 100663296 --steps 25 --warmup 1
