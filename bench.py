#!/usr/bin/env python
"""bench.py — depth frames/s of the voxel-hashing TSDF hot path (allocate -> integrate -> marching cubes).

Contract: `python bench.py --gpus N --steps K --warmup W [--impl reference]` prints ONE JSON line on rank 0.

Workload (BASELINE.json configs[1], the headline single-GPU config): the synthetic 640x480 sequence of 500 frames at
5 mm voxels, 8^3 blocks, 2^20-bucket x 4 hash, truncation 2.5 cm, MaxDepth 10, colour images. A *step* is one batch of
FRAMES_PER_STEP consecutive frames through the whole per-frame path (what the reference's processFrame does:
upload, visible-block allocation, TSDF integrate, working-set marching cubes); K=10 steps = the 500-frame sequence,
started from an empty map.

  value   frames/s with the frames already resident in HBM (vh_integrate_device back to back, one sync at the end)
  e2e     frames/s through the reference-facing call (vh_integrate == GpuTsdfGenerator::processFrame): pinned HOST
          buffers, the H2D copies of depth+rgb and the D2H read of the frame's counters inside the timed region,
          synchronous per frame like the reference
          e2e.async_value / e2e.u16_async: the non-blocking calls (vh_integrate_async, vh_integrate_u16_async) from the same
          pinned buffers; u16 = the reference loader's millimetre samples, converted on the GPU
  export  after the last frame: ordered gather + GPU vertex welding + one D2H, and the binary PLY
  roofline  the integrate kernel: algorithmic bytes (16 B x voxel updates + 4 W H + 12 B x visible blocks, + colour
          8 B x updates + 3 W H) / its CUDA-event time, vs the measured HBM copy peak
  cpu_baseline  the CPU oracle (port of the reference's algorithm, oracle/vh_oracle.c) on this box's host cores,
          bounded sample of the same workload

N > 1: one process per GPU (torchrun). The path shards by independent maps (BASELINE config 5: one sequence and one
map per GPU, phase-shifted trajectories), no data-path collective -> "scaling": "weak"; torch.distributed is used
only for the barrier and the max-over-ranks of the device time.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAMES_PER_STEP = 50
METRIC = "depth_frames_per_sec_640x480_5mm"
UNIT = "frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=None, choices=["C1", "C2", "C3", "C4"],
                    help="default: C2 (the headline single-GPU config) at --gpus 1, C4 (room scale, one map sharded over the GPUs) at --gpus > 1")
    ap.add_argument("--no-color", action="store_true")
    ap.add_argument("--no-mc", action="store_true", help="integration only (F_int); default runs working-set MC every frame like the reference")
    ap.add_argument("--cpu-frames", type=int, default=24, help="frames in the bounded cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--all-frames", action="store_true", help="config 4: the whole 100-frame sequence (~90 M blocks: needs 8 GPUs and --pool-blocks 100663296)")
    ap.add_argument("--no-ref-cuda", action="store_true", help="--gpus 1: skip the reference's own CUDA build (oracle/_ref/libref_cuda_*.so) baseline")
    ap.add_argument("--no-c4", action="store_true", help="--gpus 1: skip the room-scale config 4 side measurement (16 frames on one GPU)")
    ap.add_argument("--ray-steps", type=int, default=0, help="max_ray_steps; 0 = the reference's 100 (configs 3/4: ~1100 reach the walls, SURVEY.md section 8d)")
    ap.add_argument("--frames-per-step", type=int, default=0, help="frames in one step; 0 = 50 (config 4 with long rays allocates ~0.9 M blocks per frame: use 4)")
    ap.add_argument("--pool-blocks", type=int, default=0, help="voxel-block pool size; 0 = 3 Mi blocks (a long-ray config 4 run needs ~8 Mi)")
    return ap.parse_args()


def ray_kw(args):
    return dict(max_ray_steps=args.ray_steps) if args.ray_steps else {}


def workload(args, rank):
    synth = importlib.import_module("voxel-hashing-sdf_b200.synth")
    cfg = synth.CONFIGS[args.config]
    # config 5 style for N > 1: every rank gets its own phase-shifted trajectory over the same room
    sc = synth.make_scene(args.config, color=not args.no_color, phase=0.37 * rank)
    return synth, cfg, sc


def config_dict(args, cfg, sc, world):
    return {
        "workload": f"BASELINE config {args.config[1]}: synthetic {sc.width}x{sc.height} depth+rgb sequence, {frame_cap(args, sc)} frames"
                    + (" (the first 16 of 100: ~0.9 M new blocks per frame, 13.4 M blocks = what one GPU holds; cycled when more steps are asked for), " if frame_cap(args, sc) < sc.n_frames else ", ")
                    + 
                    f"{cfg['vox_size'] * 1000:g} mm voxels, 8^3 blocks, {cfg['num_buckets']}-bucket x4 hash, trunc {cfg['trunc'] * 100:g} cm, MaxDepth {cfg['max_depth']:g}",
        "max_ray_steps": args.ray_steps or 100,
        "frames_per_step": FRAMES_PER_STEP,
        "sequence_replay": (f"{args.steps} steps x {FRAMES_PER_STEP} frames = {args.steps * FRAMES_PER_STEP} frames over a {frame_cap(args, sc)}-frame sequence: "
                            + ("the sequence is replayed; later passes revisit the map of the first (no new blocks, weights keep growing)" if args.steps * FRAMES_PER_STEP > frame_cap(args, sc) else "one pass from an empty map")),
        "per_frame_path": "upload + allocate + integrate" + ("" if args.no_mc else " + working-set marching cubes"),
        "color": not args.no_color,
        "multi_gpu": "one independent sequence+map per GPU (config 5 style), no collective on the data path" if world > 1 else "single map",
        "l2": "inputs larger than L2: every frame is a distinct image and touches a voxel working set of "
              + {"C1": "~0.45 GB", "C2": "~0.28-0.45 GB", "C3": "~0.3 GB", "C4": "~7 GB (1.2 M visible blocks)"}.get(args.config, "hundreds of MB") + " (L2 = 126 MB)",
        # run-time switches in effect; empty = the shipped defaults
        "switches": {k: v for k, v in sorted(os.environ.items()) if k.startswith("VH_") and k != "VH_TEST_REV1"},
    }


# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_run(args, frames_wanted, per_step_frames, steps, warmup):
    """The reference's algorithm on host cores: oracle/vh_oracle.c with all OpenMP threads (set explicitly: torchrun exports
    OMP_NUM_THREADS=1 to its children)."""
    from oracle import binding as ob
    synth, cfg, sc = workload(args, 0)
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    p = ob.params_for_scene(sc, vox_size=cfg["vox_size"], trunc_margin=cfg["trunc"], max_depth=cfg["max_depth"], voxels_per_block=8,
                            use_color=0 if args.no_color else 1, run_mc=0 if args.no_mc else 1, num_threads=ncpu, **ray_kw(args))
    fresh_map_per_step = args.config == "C4"      # ~0.9 M new blocks (5 GB of host memory) per frame: bounded by starting every step from an empty map
    o = ob.Oracle(p)
    threads = ncpu
    frames = [sc.frame(i) for i in range(frames_wanted)]
    i = 0
    for _ in range(warmup * per_step_frames):
        o.process_frame(*frames[i % len(frames)]); i += 1
    n, dt = 0, 0.0
    for _ in range(steps):
        if fresh_map_per_step:
            del o
            o = ob.Oracle(p)
        t0 = time.perf_counter()
        for _ in range(per_step_frames):
            o.process_frame(*frames[i % len(frames)]); i += 1; n += 1
        dt += time.perf_counter() - t0
    return n / dt, dt, threads, n


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    synth, cfg, sc = workload(args, 0)
    per_step = 1   # bounded sample: one frame of the sequence per step
    warm = min(args.warmup, 1) if args.config == "C4" else args.warmup      # a room-scale frame takes the host cores several seconds
    fps, dt, threads, n = cpu_reference_run(args, min(args.steps + warm, frame_cap(args, sc)), per_step, args.steps, warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * dt / max(args.steps, 1), "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, cfg, sc, 1),
        "scaling_note": "the reference has no multi-GPU path: one process on the host cores at every --gpus N",
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{n} frames of the workload (1 frame per step" + (", every step from an empty map, 1 warm-up frame" if args.config == "C4" else f", frames {args.warmup}..{args.warmup + n - 1}") + "), "
                                   f"oracle/vh_oracle.c (CPU restatement of the reference's processFrame, pinned against its emulated tsdf.cu) with OpenMP on {threads} threads"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(gpu_index)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def run_ours(args):
    import numpy as np
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    vh = importlib.import_module("voxel-hashing-sdf_b200")
    if not os.path.exists(vh.LIB_PATH):
        vh.build()
    synth, cfg, sc = workload(args, rank)
    color = not args.no_color
    W, H = sc.width, sc.height
    n_timed = args.steps * FRAMES_PER_STEP
    n_warm = args.warmup * FRAMES_PER_STEP
    n_frames = min(max(n_timed, n_warm), frame_cap(args, sc))

    # ---- synthetic frames: pinned host copies (e2e) and HBM-resident copies (value) ----
    h_depth = torch.empty((n_frames, H, W), dtype=torch.float32).pin_memory()
    h_rgb = torch.empty((n_frames, H, W, 3), dtype=torch.uint8).pin_memory()
    poses = np.zeros((n_frames, 16), np.float32)
    for i in range(n_frames):
        d, rgb, c2w = sc.frame(i)
        h_depth[i] = torch.from_numpy(d); h_rgb[i] = torch.from_numpy(rgb); poses[i] = c2w
    d_depth = h_depth.cuda(non_blocking=False)
    d_rgb = h_rgb.cuda(non_blocking=False)
    torch.cuda.synchronize()

    p = vh.params_for_scene(sc, vox_size=cfg["vox_size"], trunc_margin=cfg["trunc"], max_depth=cfg["max_depth"],
                            num_buckets=cfg["num_buckets"], entries_per_bucket=4, pool_blocks=args.pool_blocks or (3 << 20),
                            use_color=1 if color else 0, mc_per_frame=0 if args.no_mc else 1, device=local, tri_arena_bytes=4 << 30, **ray_kw(args))
    eng = vh.TsdfEngine(p)
    stream = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", local))
    dptr = lambda t, i: t[i].data_ptr()
    rgb_dev = (lambda i: dptr(d_rgb, i)) if color else (lambda i: None)
    rgb_host = (lambda i: dptr(h_rgb, i)) if color else (lambda i: None)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def frames_of(n):
        return [i % n_frames for i in range(n)]

    # ---- warm-up (untimed), then start the timed sequences from an empty map ----
    for i in frames_of(n_warm):
        eng.integrate_device(dptr(d_depth, i), rgb_dev(i), poses[i])
    eng.sync()
    for i in frames_of(min(n_warm, 2 * FRAMES_PER_STEP)):
        eng.processFrame(h_depth[i].numpy(), h_rgb[i].numpy() if color else None, poses[i])

    sampler = ClockSampler(local) if rank == 0 else None

    # ---- value: inputs resident in HBM ----
    eng.reset()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in frames_of(n_timed):
        eng.integrate_device(dptr(d_depth, i), rgb_dev(i), poses[i])
    e1.record(stream)
    eng.sync()
    barrier()
    ms_value = max_over_ranks(e0.elapsed_time(e1))
    st_end = eng.stats()

    # ---- e2e: the reference-facing synchronous call with pinned host buffers ----
    eng.reset()
    barrier()
    L = eng.L
    hp = eng.h
    t0 = time.perf_counter()
    e0.record(stream)
    for i in frames_of(n_timed):
        rc = L.vh_integrate(hp, dptr(h_depth, i), rgb_host(i), poses[i].ctypes.data)
        if rc != 0:
            raise vh.VhError(rc, L.vh_last_error().decode())
    e1.record(stream)
    torch.cuda.synchronize()
    wall_e2e = (time.perf_counter() - t0) * 1000.0
    barrier()
    ms_e2e = max_over_ranks(max(e0.elapsed_time(e1), wall_e2e))

    # ---- the same host buffers through the non-blocking call (vh_integrate_async, one vh_sync at the end): uploads of
    #      frame k+1 overlap the kernels of frame k ----
    eng.reset()
    barrier()
    t0 = time.perf_counter()
    e0.record(stream)
    for i in frames_of(n_timed):
        rc = L.vh_integrate_async(hp, dptr(h_depth, i), rgb_host(i), poses[i].ctypes.data)
        if rc != 0:
            raise vh.VhError(rc, L.vh_last_error().decode())
    eng.sync()
    e1.record(stream)
    torch.cuda.synchronize()
    wall_async = (time.perf_counter() - t0) * 1000.0
    barrier()
    ms_e2e_async = max_over_ranks(max(e0.elapsed_time(e1), wall_async))

    # ---- frame ingestion as the reference's loader delivers it (u16 millimetres, SaveFrame.cpp:174-180): 2 bytes per depth
    #      sample over PCIe, converted on the GPU (vh_integrate_u16_async); measured for SURVEY.md 8(f) row 2 ----
    u16 = None
    try:
        if world > 1:
            raise RuntimeError("measured at --gpus 1 only")
        h_depth16 = (h_depth * 1000.0).round().to(torch.int16).pin_memory()      # bit pattern of the u16 sample (< 32768 mm)
        eng.reset()
        barrier()
        t0 = time.perf_counter()
        e0.record(stream)
        for i in frames_of(n_timed):
            rc = L.vh_integrate_u16_async(hp, dptr(h_depth16, i), 0.001, rgb_host(i), poses[i].ctypes.data)
            if rc != 0:
                raise vh.VhError(rc, L.vh_last_error().decode())
        eng.sync()
        e1.record(stream)
        torch.cuda.synchronize()
        wall_u16 = (time.perf_counter() - t0) * 1000.0
        barrier()
        ms_u16 = max_over_ranks(max(e0.elapsed_time(e1), wall_u16))
        u16 = {"value": sum_over_ranks(float(n_timed)) / (ms_u16 / 1000.0),
               "h2d_bytes_per_step": (W * H * 2 + (W * H * 3 if color else 0) + 64) * FRAMES_PER_STEP,
               "call": "vh_integrate_u16_async per frame (pinned u16-millimetre depth + rgb, converted on the GPU) + one vh_sync"}
        del h_depth16
    except Exception as ex:      # never lose the headline to an auxiliary measurement
        u16 = {"error": f"{type(ex).__name__}: {ex}"}

    # ---- per-kernel profile pass (untimed for the headline): CUDA-event time of every stage of every frame ----
    eng.reset()
    acc = dict(alloc=0.0, integrate=0.0, mc=0.0, cull=0.0, upload=0.0, updates=0, visible=0, tris=0, culled=0)
    per_frame_rows = []          # (frame, ms_integrate, voxel updates, visible blocks, blocks discarded whole)
    for i in frames_of(n_timed):
        eng.integrate_device(dptr(d_depth, i), rgb_dev(i), poses[i])
        s = eng.stats()
        acc["alloc"] += s.ms_alloc; acc["integrate"] += s.ms_integrate; acc["mc"] += s.ms_mc; acc["cull"] += s.ms_cull
        acc["updates"] += s.voxel_updates; acc["visible"] += s.visible_blocks; acc["tris"] += s.triangles; acc["culled"] += s.culled_blocks
        per_frame_rows.append((i, s.ms_integrate, s.voxel_updates, s.visible_blocks, s.culled_blocks, s.ms_alloc, s.ms_mc, s.triangles, s.ms_cull))
    clocks = sampler.stop() if sampler else None
    st_last = eng.stats()
    allocated = st_last.allocated_blocks

    # ---- mesh export of the finished map: persistent per-block meshes gathered in tsdf2mesh order,
    #      welded on the GPU, copied out once; then the binary PLY of the same mesh. Rank 0 only, outside every timed region. ----
    export = None
    if rank == 0 and not args.no_mc:
        try:
            import ctypes as C
            nv, nf = C.c_uint64(), C.c_uint64()
            t0 = time.perf_counter()
            rc = L.vh_weld_mesh(hp, vh.VH_MESH_REF_PERSISTENT, None, 0, C.byref(nv), None, 0, C.byref(nf))
            t_weld = time.perf_counter() - t0
            if rc != 0:
                raise vh.VhError(rc, L.vh_last_error().decode())
            ply = os.path.join(tempfile.gettempdir(), f"vh_bench_{os.getpid()}.ply")
            t0 = time.perf_counter()
            eng.save_ply_binary(ply)
            t_ply = time.perf_counter() - t0
            ply_bytes = os.path.getsize(ply)
            os.remove(ply)
            export = {"faces": int(nf.value), "vertices": int(nv.value), "ms_gather_weld_copy": 1000.0 * t_weld,
                      "faces_per_sec": nf.value / t_weld if t_weld > 0 else None,
                      "ms_save_ply_binary": 1000.0 * t_ply, "ply_bytes": ply_bytes,
                      "call": "vh_weld_mesh (counts only: ordered gather + GPU weld + one D2H) and vh_save_ply_binary after the last frame"}
        except Exception as ex:
            export = {"error": f"{type(ex).__name__}: {ex}"}

    total_frames = sum_over_ranks(float(n_timed))
    value = total_frames / (ms_value / 1000.0)
    e2e = total_frames / (ms_e2e / 1000.0)
    upd_per_frame = acc["updates"] / n_timed
    vis_per_frame = acc["visible"] / n_timed
    tris_per_frame = acc["tris"] / n_timed
    bytes_int = 16.0 * upd_per_frame + 4.0 * W * H + 12.0 * vis_per_frame + ((8.0 * upd_per_frame + 3.0 * W * H) if color else 0.0)
    bytes_mc = (4.0 + (3.0 if color else 0.0)) * 512 * vis_per_frame + 48.0 * tris_per_frame + 12.0 * vis_per_frame
    ms_int = acc["integrate"] / n_timed
    ms_mc = acc["mc"] / n_timed
    ms_alloc = acc["alloc"] / n_timed
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    ach = bytes_int / (ms_int * 1e-3) / 1e9 if ms_int > 0 else 0.0
    # the same fraction frame by frame (the mix matters: frames in which few voxels pass the gate carry few algorithmic bytes)
    def frame_bytes(u, v):
        return 16.0 * u + 4.0 * W * H + 12.0 * v + ((8.0 * u + 3.0 * W * H) if color else 0.0)
    fr = sorted(frame_bytes(r[2], r[3]) / (r[1] * 1e-3) / 1e9 / peak for r in per_frame_rows if r[1] > 0)
    frac_frames = {"min": fr[0], "p10": fr[len(fr) // 10], "median": fr[len(fr) // 2], "p90": fr[(9 * len(fr)) // 10], "max": fr[-1]} if fr else None
    dump = os.environ.get("VH_BENCH_DUMP")
    if dump and rank == 0:
        with open(dump, "w") as f:
            f.write("frame,ms_integrate,voxel_updates,visible_blocks,blocks_discarded,ms_alloc,ms_mc,triangles,ms_cull\n")
            for r in per_frame_rows:
                f.write(",".join(str(x) for x in r) + "\n")
    # DRAM traffic of ONE named launch (ncu --set full capture committed under profiles/) next to the algorithmic bytes of the SAME launch
    traffic, traffic_src, traffic_alg = None, None, None
    tpath = os.path.join(ROOT, "profiles", "integrate_traffic.json")
    if os.path.exists(tpath) and args.config == "C2" and color:
        tj = json.load(open(tpath)); traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
        fr_i = int(tj.get("sequence_frame", -1))
        row = next((r for r in per_frame_rows if r[0] == fr_i), None)
        if row is not None:
            traffic_alg = 16.0 * row[2] + 4.0 * W * H + 12.0 * row[3] + 8.0 * row[2] + 3.0 * W * H

    h2d = W * H * 4 + (W * H * 3 if color else 0) + 64
    d2h = 104
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_value / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, cfg, sc, world),
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d * FRAMES_PER_STEP, "d2h_bytes_per_step": d2h * FRAMES_PER_STEP,
                "call": "vh_integrate (GpuTsdfGenerator::processFrame drop-in), pinned host depth+rgb, synchronous per frame",
                "async_value": total_frames / (ms_e2e_async / 1000.0),
                "async_call": "vh_integrate_async per frame from the same pinned host buffers + one vh_sync: uploads overlap kernels",
                "u16_async": u16},
        "gpu_launches": (6 if not args.no_mc else 4) * n_timed + (0 if os.environ.get("VH_STATUS_PUBLISH") == "0" else n_timed)
                        + (n_timed if os.environ.get("VH_ALLOC_REV") == "2" else 0),      # pack, allocate (1 or 2 kernels), work list, integrate (+ mc_filter, mc_mesh), status per frame
        "voxel_updates_per_sec": sum_over_ranks(float(acc["updates"])) / (ms_value / 1000.0),
        "per_frame": {"voxel_updates": upd_per_frame, "visible_blocks": vis_per_frame, "blocks_discarded_whole": acc["culled"] / n_timed,
                      "triangles": tris_per_frame,
                      "ms_alloc": ms_alloc, "ms_cull_list": acc["cull"] / n_timed, "ms_integrate": ms_int, "ms_mc": ms_mc, "allocated_blocks_end": allocated,
                      "arena_compactions_in_500_frames": int(st_last.arena_compactions)},
        "roofline": {"kernel": "vh::integrate_kernel_staged" if os.environ.get("VH_INTEGRATE_REV") == "2" else "vh::integrate_kernel_direct", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "frac_of_nominal_8TBs": ach / 8000.0, "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_src,
                     "traffic_launch_algorithmic_bytes": traffic_alg, "algorithmic_bytes_per_launch": bytes_int, "avg_launch_ms": ms_int, "frac_per_frame": frac_frames},
        "marching_cubes": {"kernels": "vh::mc_filter_kernel + vh::mc_mesh_kernel", "avg_ms_per_frame": ms_mc, "triangles_per_frame": tris_per_frame,
                           "dram_bytes_per_frame_ncu": 31.3e6, "dram_source": "profiles/r02_ncu/ncu_full_mcfilter.txt + ncu_full_mcmesh.txt (23.6 MB + 7.7 MB read; the triangle writes stay in L2)",
                           "note": "latency-bound chains of look-ups; the sign filter keeps the kernels from reading most of the working set, so an algorithmic-bytes/time figure would overstate bandwidth use: the statement is time"},
        "clocks": clocks,
        "export": export,
    }
    # ---- the room-scale config 4 on this one GPU (the N = 1 point of the multi-GPU curve `bench.py --gpus N` reports) ----
    if world == 1 and args.config == "C2" and not args.no_c4:
        eng.close(); eng = None
        del d_depth, d_rgb, h_depth, h_rgb
        torch.cuda.empty_cache()
        try:
            c4 = synth.CONFIGS["C4"]
            s4 = synth.make_scene("C4", color=color)
            fps4 = C4_DEFAULTS["frames_per_step"]
            line["room_scale"] = dict(single_gpu_run(vh, torch, np, "C4", c4, s4, color, not args.no_mc, local, fps4, 4 * fps4, C4_DEFAULTS["ray_steps"],
                                                     C4_DEFAULTS["pool_blocks"], max_frames=C4_DEFAULTS["frames"]), n_gpus=1,
                                      workload="BASELINE config 4: 640x480, 2 mm voxels, 10 m room centred on the origin, 2^24-bucket x4 hash, trunc 1 cm, "
                                               "max_ray_steps 1100, first 16 frames (~0.9 M new blocks per frame: the 100-frame map does not fit one GPU)")
        except Exception as ex:
            line["room_scale"] = {"error": f"{type(ex).__name__}: {ex}"}
    # ---- north_star's first baseline: the reference's OWN CUDA build on this box (oracle/build_ref_cuda.sh compiles its src/tsdf.cu for
    #      sm_100a; oracle/tools/bench_ref_cuda.py drives its processFrame and this engine's vh_integrate over the same frames) ----
    if world == 1 and args.config == "C2" and not args.no_ref_cuda:
        line["ref_cuda_baseline"] = ref_cuda_baseline()
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            fps, dt, threads, n = cpu_reference_run(args, args.cpu_frames, 1, args.cpu_frames, 0)
            line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"first {n} frames of the same workload ({dt:.1f} s), oracle/vh_oracle.c with OpenMP on {threads} threads"}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if eng is not None:
        eng.close()
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------
def ref_cuda_baseline():
    """The reference's own CUDA build (baseline only, never the optimisation target). One subprocess per config: the reference keeps
    its tables in file-scope globals (tsdf.cu:19-20). Config 2 is outside what the reference can run (see `config2`)."""
    out = {"config2": "not runnable by the reference build: at 5 mm its stream-in stage would upload ~3.6 M blocks (22 GB) per frame into a "
                      "400,000-block heap and keep ~0.5 TB of host triangle slots (SURVEY.md 8(a13), BASELINE.md 3a)"}
    tool = os.path.join(ROOT, "oracle", "tools", "bench_ref_cuda.py")
    for cfg_name, frames, warm, variant in (("R8", 4, 1, "8"), ("C1", 2, 1, "8c1")):
        lib = os.path.join(ROOT, "oracle", "_ref", f"libref_cuda_vpb{variant}.so")
        if not os.path.exists(lib):
            out[cfg_name] = {"unavailable": f"{os.path.relpath(lib, ROOT)} not built (oracle/build_ref_cuda.sh needs /root/reference)"}
            continue
        try:
            r = subprocess.run([sys.executable, tool, "--config", cfg_name, "--frames", str(frames), "--warmup", str(warm)], capture_output=True, text=True, timeout=300, cwd=ROOT)
            js = [l for l in r.stdout.splitlines() if l.startswith("{")]
            if not js:
                out[cfg_name] = {"error": (r.stderr or r.stdout)[-400:]}
                continue
            d = json.loads(js[-1])
            out[cfg_name] = {k: d[k] for k in ("baseline", "config", "frames_timed", "call", "reference_frames_per_sec", "ours_frames_per_sec", "speedup",
                                               "faces_reference", "faces_ours", "visible_blocks_reference", "visible_blocks_ours", "note") if k in d}
        except Exception as ex:
            out[cfg_name] = {"error": f"{type(ex).__name__}: {ex}"}
    return out


def peak_hbm():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def make_frames(torch, np, sc, n_frames, color, pinned=True, device=True):
    H, W = sc.height, sc.width
    h_depth = torch.empty((n_frames, H, W), dtype=torch.float32)
    h_rgb = torch.empty((n_frames, H, W, 3), dtype=torch.uint8)
    if pinned:
        h_depth, h_rgb = h_depth.pin_memory(), h_rgb.pin_memory()
    poses = np.zeros((n_frames, 16), np.float32)
    for i in range(n_frames):
        d, rgb, c2w = sc.frame(i)
        h_depth[i] = torch.from_numpy(d); h_rgb[i] = torch.from_numpy(rgb); poses[i] = c2w
    d_depth = h_depth.cuda() if device else None
    d_rgb = h_rgb.cuda() if device else None
    return h_depth, h_rgb, d_depth, d_rgb, poses


def single_gpu_run(vh, torch, np, cfg_name, cfg, sc, color, mc, local, n_warm, n_timed, ray_steps, pool_blocks, frames=None, max_frames=None):
    """One map on ONE GPU, frames resident in its HBM: frames/s, voxel updates/s and the per-stage times (used for config 4 at
    --gpus 1 and, on rank 0, as the same-run single-GPU point of a multi-GPU run)."""
    n_frames = min(max(n_timed, n_warm), max_frames or sc.n_frames)
    if frames is None:
        frames = make_frames(torch, np, sc, n_frames, color, pinned=False)
    _, _, d_depth, d_rgb, poses = frames
    n_frames = len(poses)
    p = vh.params_for_scene(sc, vox_size=cfg["vox_size"], trunc_margin=cfg["trunc"], max_depth=cfg["max_depth"], num_buckets=cfg["num_buckets"],
                            entries_per_bucket=4, pool_blocks=pool_blocks, use_color=1 if color else 0, mc_per_frame=1 if mc else 0, device=local,
                            tri_arena_bytes=4 << 30, **(dict(max_ray_steps=ray_steps) if ray_steps else {}))
    eng = vh.TsdfEngine(p)
    stream = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", local))
    rgbp = (lambda i: d_rgb[i].data_ptr()) if color else (lambda i: None)
    for k in range(n_warm):
        i = k % n_frames
        eng.integrate_device(d_depth[i].data_ptr(), rgbp(i), poses[i])
    eng.sync()
    eng.reset()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(n_timed):
        i = k % n_frames
        eng.integrate_device(d_depth[i].data_ptr(), rgbp(i), poses[i])
    e1.record(stream)
    eng.sync()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    updates_total = eng.stats().voxel_updates_total
    allocated = eng.stats().allocated_blocks
    # per-stage pass (untimed for the figures above)
    eng.reset()
    acc = dict(alloc=0.0, cull=0.0, integrate=0.0, mc=0.0, updates=0, visible=0, culled=0, tris=0)
    for k in range(n_timed):
        i = k % n_frames
        eng.integrate_device(d_depth[i].data_ptr(), rgbp(i), poses[i])
        st = eng.stats()
        acc["alloc"] += st.ms_alloc; acc["cull"] += st.ms_cull; acc["integrate"] += st.ms_integrate; acc["mc"] += st.ms_mc
        acc["updates"] += st.voxel_updates; acc["visible"] += st.visible_blocks; acc["culled"] += st.culled_blocks; acc["tris"] += st.triangles
    eng.close()
    W, H = sc.width, sc.height
    u, v = acc["updates"] / n_timed, acc["visible"] / n_timed
    bytes_int = 16.0 * u + 4.0 * W * H + 12.0 * v + ((8.0 * u + 3.0 * W * H) if color else 0.0)
    ms_int = acc["integrate"] / n_timed
    peak, _ = peak_hbm()
    return {"frames_per_sec": n_timed / (ms / 1000.0), "voxel_updates_per_sec": updates_total / (ms / 1000.0), "ms_per_frame": ms / n_timed, "frames": n_timed,
            "allocated_blocks_end": int(allocated),
            "per_frame": {"voxel_updates": u, "visible_blocks": v, "blocks_discarded_whole": acc["culled"] / n_timed, "triangles": acc["tris"] / n_timed,
                          "ms_alloc": acc["alloc"] / n_timed, "ms_cull_list": acc["cull"] / n_timed, "ms_integrate": ms_int, "ms_mc": acc["mc"] / n_timed},
            "integrate_roofline_frac": (bytes_int / (ms_int * 1e-3) / 1e9 / peak) if ms_int > 0 else None}


def sharded_run(vh, torch, np, dist, cfg, sc, color, mc, rank, world, local, n_warm, n_timed, ray_steps, pool_blocks, frames, with_host_inputs=True):
    """ONE map sharded over the GPUs by block hash on rank 0's sequence (north_star's multi-GPU path): per frame rank 0 puts
    the frame into its ring (device copy, or H2D from pinned host buffers for the e2e figure), every other GPU's copy engine pulls it
    over NVLink one frame ahead, every GPU marches its share of the rays, sends block keys to their owners, integrates and meshes its
    own blocks. Strong scaling: the work is fixed. Time = max over ranks of the CUDA-event time on the engine's stream (and of the
    wall clock of the enqueue loop, whichever is larger). Returns (device-resident result, host-input result) from ONE engine."""
    h_depth, h_rgb, d_depth, d_rgb, poses = frames
    n_frames = len(poses)
    p = vh.params_for_scene(sc, vox_size=cfg["vox_size"], trunc_margin=cfg["trunc"], max_depth=cfg["max_depth"], num_buckets=cfg["num_buckets"],
                            entries_per_bucket=4, pool_blocks=pool_blocks // world + (1 << 19), use_color=1 if color else 0, mc_per_frame=1 if mc else 0,
                            device=local, shard_rank=rank, shard_count=world, tri_arena_bytes=(4 << 30) // world + (1 << 30),
                            **(dict(max_ray_steps=ray_steps) if ray_steps else {}))
    eng = vh.TsdfEngine(p)
    ids = [vh.TsdfEngine.shard_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    eng.shard_connect(ids[0])
    stream = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", local))

    def one_pass(n, host_inputs):
        for k in range(n):
            i = k % n_frames
            if rank != 0:
                eng.integrate_sharded_device(None, None, poses[i])
            elif host_inputs:
                eng.integrate_sharded(h_depth[i].data_ptr(), h_rgb[i].data_ptr() if color else None, poses[i])
            else:
                eng.integrate_sharded_device(d_depth[i].data_ptr(), d_rgb[i].data_ptr() if color else None, poses[i])
        eng.sync()

    def measure(host_inputs, warm):
        one_pass(warm, host_inputs)
        eng.reset()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        one_pass(n_timed, host_inputs)
        e1.record(stream)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1000.0
        dist.barrier()
        t = torch.tensor([max(e0.elapsed_time(e1), wall)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        mine = eng.stats()
        u = torch.tensor([float(mine.voxel_updates_total), float(mine.allocated_blocks), float(mine.allocated_blocks)], dtype=torch.float64, device="cuda")
        dist.all_reduce(u[:2], op=dist.ReduceOp.SUM)
        dist.all_reduce(u[2:], op=dist.ReduceOp.MAX)
        return {"frames_per_sec": n_timed / (ms / 1000.0), "voxel_updates_per_sec": float(u[0].item()) / (ms / 1000.0), "ms_per_frame": ms / n_timed, "frames": n_timed,
                "allocated_blocks_all_shards": int(u[1].item()), "allocated_blocks_largest_shard": int(u[2].item()),
                "last_frame_this_rank": {"ms_alloc": mine.ms_alloc, "ms_cull_list": mine.ms_cull, "ms_integrate": mine.ms_integrate, "ms_mc": mine.ms_mc}}
    dev = measure(False, n_warm)
    host = measure(True, min(n_warm, FRAMES_PER_STEP)) if with_host_inputs else None
    dist.barrier()
    eng.close()
    return dev, host


def run_multi(args):
    """--gpus N > 1 (torchrun, one process per GPU): `value` = frames/s of ONE map sharded over the N GPUs (default: the room-scale
    config 4, "scaling": "strong"); side keys: the same workload on one GPU in the same run (rank 0), the headline config 2 sharded,
    and config 5 (independent maps, one per GPU)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    vh = importlib.import_module("voxel-hashing-sdf_b200")
    synth = importlib.import_module("voxel-hashing-sdf_b200.synth")
    color, mc = not args.no_color, not args.no_mc
    cfg = synth.CONFIGS[args.config]
    sc = synth.make_scene(args.config, color=color)               # every rank integrates rank 0's trajectory
    n_timed, n_warm = args.steps * FRAMES_PER_STEP, args.warmup * FRAMES_PER_STEP
    n_frames = min(max(n_timed, n_warm), frame_cap(args, sc))
    pool = args.pool_blocks or (3 << 20)
    frames = make_frames(torch, np, sc, n_frames, color, pinned=(rank == 0), device=(rank == 0))
    sampler = ClockSampler(local) if rank == 0 else None
    main, e2e = sharded_run(vh, torch, np, dist, cfg, sc, color, mc, rank, world, local, n_warm, n_timed, args.ray_steps, pool, frames)
    clocks = sampler.stop() if sampler else None
    side = {}

    def guarded(name, fn):        # an auxiliary measurement never costs the headline
        try:
            side[name] = fn()
        except Exception as ex:
            side[name] = {"error": f"{type(ex).__name__}: {ex}"}
        dist.barrier()

    # the same workload on ONE GPU in the same run (rank 0; the other ranks wait): the N = 1 point of this curve
    def single():
        if rank != 0:
            return None
        need_gb = pool * (6144 if color else 4096) / 1e9
        if need_gb > 150:
            return {"not_run": f"a pool of {pool} blocks ({need_gb:.0f} GB) does not fit one GPU: this workload exists on the sharded map only"}
        return single_gpu_run(vh, torch, np, args.config, cfg, sc, color, mc, local, n_warm, n_timed, args.ray_steps, pool, frames=frames)
    guarded("single_gpu_same_run", single)
    if args.config != "C2":       # the headline config sharded the same way: a 0.2 ms frame, latency-bound
        def c2():
            c = synth.CONFIGS["C2"]
            s2 = synth.make_scene("C2", color=color)
            f2 = make_frames(torch, np, s2, 100, color, pinned=(rank == 0), device=(rank == 0))
            return sharded_run(vh, torch, np, dist, c, s2, color, mc, rank, world, local, 50, 100, 0, 3 << 20, f2, with_host_inputs=False)[0]
        guarded("headline_c2_sharded", c2)
    # BASELINE config 5: one independent 640x480 / 5 mm sequence and map per GPU (phase-shifted trajectories), no data-path collective
    def replicas():
        c = synth.CONFIGS["C2"]
        s5 = synth.make_scene("C2", color=color, phase=0.37 * rank)
        r = single_gpu_run(vh, torch, np, "C2", c, s5, color, mc, local, 50, 150, 0, 3 << 20)
        t = torch.tensor([r["ms_per_frame"] * r["frames"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return {"frames_per_sec_all_gpus": world * r["frames"] / (float(t.item()) / 1000.0), "frames_per_gpu": r["frames"], "scaling": "weak"}
    guarded("config5_independent_maps", replicas)

    if rank == 0:
        W, H = sc.width, sc.height
        h2d = W * H * 4 + (W * H * 3 if color else 0) + 64
        nkern = 7 + (3 if mc else 0)      # frame wait/ready, pack, ray keys, barrier, insert, work list, integrate (+ barrier, mc filter, mc mesh), status
        line = {
            "metric": METRIC, "value": main["frames_per_sec"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": main["ms_per_frame"] * FRAMES_PER_STEP, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": dict(config_dict(args, cfg, sc, world), multi_gpu="ONE map sharded by block-coordinate hash (owner = hash of the block's 8^3-block cube mod n_gpus): "
                           "rank 0's frame read by every GPU's pack kernel over NVLink, rays split across the GPUs, keys sent to their owners' inboxes, "
                           "every GPU integrates and meshes its own blocks (marching-cubes halos read from the owner GPU), two flag barriers per frame",
                           note="at --gpus 1 the line's workload is BASELINE config 2 (metric depth_frames_per_sec_640x480_5mm); this line's N = 1 point is "
                                "room_scale.single_gpu_same_run (the same frames on one GPU in the same run)"),
            "voxel_updates_per_sec": main["voxel_updates_per_sec"],
            "e2e": {"value": e2e["frames_per_sec"], "unit": UNIT, "h2d_bytes_per_step": h2d * FRAMES_PER_STEP, "d2h_bytes_per_step": 128 * FRAMES_PER_STEP * world,
                    "call": "vh_integrate_sharded per frame: pinned host depth+rgb on rank 0 -> H2D into rank 0's frame ring (upload stream) -> every GPU; one vh_sync at the end"},
            "gpu_launches": nkern * n_timed * world,
            "sharded": main, "sharded_e2e": e2e,
            "room_scale" if args.config == "C4" else "curve": {"n_gpus": world, "frames_per_sec": main["frames_per_sec"], "voxel_updates_per_sec": main["voxel_updates_per_sec"],
                                                              "single_gpu_same_run": side.get("single_gpu_same_run")},
            "headline_c2_sharded": side.get("headline_c2_sharded"),
            "config5_independent_maps": side.get("config5_independent_maps"),
            "roofline": None, "cpu_baseline": None, "clocks": clocks,
        }
        sg = side.get("single_gpu_same_run") or {}
        if isinstance(sg, dict) and sg.get("per_frame"):
            peak, peak_src = peak_hbm()
            pf = sg["per_frame"]
            b = 16.0 * pf["voxel_updates"] + 4.0 * W * H + 12.0 * pf["visible_blocks"] + ((8.0 * pf["voxel_updates"] + 3.0 * W * H) if color else 0.0)
            ach = b / (pf["ms_integrate"] * 1e-3) / 1e9
            line["roofline"] = {"kernel": "vh::integrate_kernel_direct" if os.environ.get("VH_INTEGRATE_REV") == "1" else "vh::integrate_kernel_staged",
                                "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": peak_src, "traffic": None,
                                "algorithmic_bytes_per_launch": b, "avg_launch_ms": pf["ms_integrate"], "measured_on": "one GPU, whole map (single_gpu_same_run)"}
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()


C4_DEFAULTS = dict(ray_steps=1100, frames_per_step=4, pool_blocks=16 << 20, frames=16)    # room scale: the step cap scaled to the block size; ~0.9 M new blocks per
# frame, so the sequence is the first 16 frames (13.4 M blocks, 80 GB: what ONE GPU holds), cycled when more steps are asked for


def frame_cap(args, sc):
    return min(sc.n_frames, C4_DEFAULTS["frames"]) if args.config == "C4" and not args.all_frames else sc.n_frames


def apply_defaults(a):
    """--gpus 1: BASELINE config 2 (headline). --gpus N > 1: BASELINE config 4 (room scale, ONE map sharded over the GPUs)."""
    global FRAMES_PER_STEP, METRIC
    if a.config is None:
        a.config = "C2" if a.gpus <= 1 else "C4"
    if a.config == "C4":
        a.ray_steps = a.ray_steps or C4_DEFAULTS["ray_steps"]
        a.frames_per_step = a.frames_per_step or C4_DEFAULTS["frames_per_step"]
        a.pool_blocks = a.pool_blocks or C4_DEFAULTS["pool_blocks"]
        METRIC = "depth_frames_per_sec_640x480_2mm_room_scale"
    elif a.config != "C2":
        METRIC = f"depth_frames_per_sec_{a.config}"
    if a.frames_per_step > 0:
        FRAMES_PER_STEP = a.frames_per_step


if __name__ == "__main__":
    a = parse()
    apply_defaults(a)
    if a.impl == "reference":
        run_reference_arm(a)
    elif int(os.environ.get("WORLD_SIZE", "1")) > 1:
        run_multi(a)
    else:
        run_ours(a)
