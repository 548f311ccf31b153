#!/usr/bin/env bash
# one bench.py --gpus N line with the driver's arguments (gpurun --gpus N)
N="${1:-4}"; TAG="${2:-r02l}"; OUT=gpurun_out; mkdir -p $OUT
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 ) > $OUT/bench_${TAG}_n$N.log 2>&1
python - <<PY
import json
ok = False
for l in open("$OUT/bench_${TAG}_n$N.log"):
    if l.startswith("{"):
        ok = True
        d = json.loads(l)
        sg = (d.get("room_scale") or {}).get("single_gpu_same_run") or {}
        print("n=$N", d["metric"], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "Gupd/s", round(d["voxel_updates_per_sec"] / 1e9, 1),
              "| single GPU same run:", round(sg.get("frames_per_sec", 0), 1) if "error" not in sg else sg, "| last frame rank0", {k: round(v, 3) for k, v in d["sharded"]["last_frame_this_rank"].items()},
              "| largest shard", d["sharded"]["allocated_blocks_largest_shard"], "of", d["sharded"]["allocated_blocks_all_shards"],
              "| c2 sharded", (d.get("headline_c2_sharded") or {}).get("frames_per_sec") or d.get("headline_c2_sharded"), "| config5", (d.get("config5_independent_maps") or {}))
if not ok:
    print("no JSON line"); print(open("$OUT/bench_${TAG}_n$N.log").read()[-2500:])
PY
grep -E "^real" $OUT/bench_${TAG}_n$N.log
