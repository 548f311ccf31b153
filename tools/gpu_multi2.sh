#!/usr/bin/env bash
# 2-GPU validation (gpurun --gpus 2): sharded-map parity in every mode, compute-sanitizer memcheck of a 2-rank sharded run, then one
# bench.py --gpus 2 line with the driver's arguments.   usage: tools/gpu_multi2.sh [tag]
TAG="${1:-r02j}"; OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_worker.py > $OUT/multi_worker_${TAG}_n2.log 2>&1; echo "worker rc=$?" >> $OUT/multi_worker_${TAG}_n2.log
grep -E "ok \[|MULTI_GPU_OK|rc=|Error|error" $OUT/multi_worker_${TAG}_n2.log | head -20
# memcheck of both ranks of a small sharded run (2 frames, one mode): every kernel of the sharded path incl. the peer-memory accesses
VH_MULTI_FRAMES=2 VH_MULTI_ONLY="pull+split-rays, frames" timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 --no-python \
    compute-sanitizer --tool memcheck --error-exitcode 9 python tests/multi_gpu_worker.py > $OUT/sanitizer_memcheck_sharded_${TAG}_n2.log 2>&1; echo "sharded memcheck rc=$?" >> $OUT/sanitizer_memcheck_sharded_${TAG}_n2.log
grep -E "ERROR SUMMARY|MULTI_GPU_OK|rc=|not supported|Error" $OUT/sanitizer_memcheck_sharded_${TAG}_n2.log | head -8
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 ) > $OUT/bench_${TAG}_n2.log 2>&1
python - <<PY
import json
ok = False
for l in open("$OUT/bench_${TAG}_n2.log"):
    if l.startswith("{"):
        ok = True
        d = json.loads(l)
        sg = (d.get("room_scale") or {}).get("single_gpu_same_run") or {}
        print("n=2", d["metric"], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "Gupd/s", round(d["voxel_updates_per_sec"] / 1e9, 1),
              "| single GPU same run:", round(sg.get("frames_per_sec", 0), 1) if "error" not in sg else sg, "| last frame rank0", {k: round(v, 3) for k, v in d["sharded"]["last_frame_this_rank"].items()},
              "| c2 sharded", (d.get("headline_c2_sharded") or {}).get("frames_per_sec") or d.get("headline_c2_sharded"), "| config5", (d.get("config5_independent_maps") or {}), "| roofline", (d.get("roofline") or {}).get("frac"))
if not ok:
    print("no JSON line"); print(open("$OUT/bench_${TAG}_n2.log").read()[-2500:])
PY
grep -E "^real" $OUT/bench_${TAG}_n2.log
