#!/usr/bin/env python
"""CPU baseline: the oracle (port of the reference's algorithm) on the box's host cores, per stage, with 1
thread and with all OpenMP threads. usage: python oracle/tools/bench_cpu_oracle.py --config C1 --frames 6"""
import argparse, importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
from oracle import binding as ob  # noqa: E402

ap = argparse.ArgumentParser(); ap.add_argument("--config", default="C1"); ap.add_argument("--frames", type=int, default=6); a = ap.parse_args()
synth = importlib.import_module("voxel-hashing-sdf_b200.synth")
cfg = synth.CONFIGS[a.config]; sc = synth.make_scene(a.config, color=True)
ob.build()
frames = [sc.frame(i) for i in range(a.frames)]
out = {"config": a.config, "frames": a.frames, "host_threads_available": ob.lib().vo_threads()}
for label, nt in (("1_thread", 1), ("all_threads", 0)):
    o = ob.Oracle(ob.params_for_scene(sc, vox_size=cfg["vox_size"], trunc_margin=cfg["trunc"], max_depth=cfg["max_depth"], voxels_per_block=8,
                                      use_color=1, run_mc=1, num_threads=nt))
    t0 = time.perf_counter(); stage = [0.0, 0.0, 0.0]
    for f in frames:
        o.process_frame(*f)
        ts = o.last_times() if hasattr(o, "last_times") else None
        if ts is not None:
            stage = [x + ts[k] for x, k in zip(stage, ("alloc", "integrate", "mc"))]
    dt = time.perf_counter() - t0
    out[label] = {"frames_per_sec": a.frames / dt, "ms_per_frame": 1e3 * dt / a.frames,
                  "ms_alloc_integrate_mc": [1e3 * x / a.frames for x in stage] if any(stage) else None}
print(json.dumps(out))
