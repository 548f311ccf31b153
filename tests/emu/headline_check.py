#!/usr/bin/env python3
"""BASELINE-shaped frames WITHOUT a GPU: the engine's kernel sources under CPU emulation (tests/emu) against the oracle, bit for bit —
visible set, voxel updates and triangle count per frame, then every voxel, the ordered triangle soup and the full-map mesh.
About one minute per 640x480 frame. Examples:
  tests/emu/headline_check.py --config C2 --frames 3 --revs 2 2 1       (revs: integrate 1 direct / 2 staged, allocation 0 one kernel / 2 keys + insert, MC 1 = colour tile)
  tests/emu/headline_check.py --config C3 --frames 1                   (1280x720, 4 mm, 2^24 buckets)
  tests/emu/headline_check.py --config C4 --frames 1 --ranks 4         (room scale, 2 mm, one map sharded over 4 emulated ranks)"""
import argparse
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C2", choices=["C1", "C2", "C3", "C4"])
ap.add_argument("--frames", type=int, default=3)
ap.add_argument("--revs", type=int, nargs=3, default=[0, 0, 0], metavar=("INTEGRATE", "ALLOC", "MC"))
ap.add_argument("--ranks", type=int, default=1)
ap.add_argument("--ray-steps", type=int, default=0, help="max_ray_steps (default: the reference's 100)")
ap.add_argument("--buckets-log2", type=int, default=0, help="override the config's bucket count (emulation memory)")
args = ap.parse_args()

vh = importlib.import_module("voxel-hashing-sdf_b200")
synth = importlib.import_module("voxel-hashing-sdf_b200.synth")
from oracle import binding as ob  # noqa: E402
from emu.binding import EmuEngine, EmuGroup, mesh_order  # noqa: E402
from util import engine_params, oracle_params, key_set  # noqa: E402

cfg = synth.CONFIGS[args.config]
sc = synth.make_scene(args.config, color=True, holes=0.01)
case = dict(scene=dict(color=True), vpb=8, vox_size=cfg["vox_size"], trunc=cfg["trunc"], max_depth=cfg["max_depth"])
buckets = (1 << args.buckets_log2) if args.buckets_log2 else cfg["num_buckets"]
over = dict(max_ray_steps=args.ray_steps) if args.ray_steps else {}
o = ob.Oracle(oracle_params(ob, sc, case, **over))
t0 = time.time()
kw = dict(integrate_rev=args.revs[0], alloc_rev=args.revs[1], mc_rev=args.revs[2])


def make(rank, n):
    return engine_params(vh, sc, case, num_buckets=buckets, pool_blocks=1 << (20 if args.ray_steps > 400 else 17), tri_arena_bytes=(args.frames + 2) * 800_000 * 48,     # no arena compaction under emulation: room for every frame's mesh
                          shard_rank=rank, shard_count=n, **over)


group = EmuGroup(make, args.ranks, **kw)
R = group.ranks
for i in range(args.frames):
    d, rgb, c2w = sc.frame(i)
    o.process_frame(d, rgb, c2w)
    group.process_frame(d, rgb, c2w)
    vis = [key_set(e.visible_keys()) for e in R]
    assert sum(len(v) for v in vis) == o.num_visible and set().union(*vis) == key_set(o.visible_keys()), f"visible set differs in frame {i}"
    assert sum(e.last_updates for e in R) == o.last_updates, (i, o.last_updates)
    assert sum(e.last_triangles for e in R) == o.last_triangles, (i, o.last_triangles)
    print("frame", i, "visible", o.num_visible, "updates", o.last_updates, "tris", o.last_triangles, f"{time.time() - t0:.0f}s", flush=True)
keys = o.all_keys()
owner = np.array([vh.owner_of_block(int(k[0]), int(k[1]), int(k[2]), args.ranks, 8) if args.ranks > 1 else 0 for k in keys])
so, wo, co, _ = o.get_blocks(keys)
for r, e in enumerate(R):
    m = owner == r
    se, we, ce, found, neg = e.get_blocks(keys[m])
    assert found.all() and np.array_equal(se.view(np.uint32), so[m].view(np.uint32)) and np.array_equal(we, wo[m]) and np.array_equal(ce, co[m]), f"voxels differ on rank {r}"
    assert np.array_equal(neg, (se < 0).sum(1))
ordered = mesh_order(keys)
own_o = [vh.owner_of_block(int(k[0]), int(k[1]), int(k[2]), args.ranks, 8) if args.ranks > 1 else 0 for k in ordered]


def gather(full):
    parts = [R[r].block_triangles(k[None, :], full_map=full) for k, r in zip(ordered, own_o)] if args.ranks > 1 else [R[0].block_triangles(ordered, full_map=full)]
    return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])


xyz_o, rgb_o = o.triangles()
xyz_e, rgb_e = gather(False)
assert xyz_e.shape == xyz_o.shape and np.array_equal(xyz_e.view(np.uint32), xyz_o.view(np.uint32)) and np.array_equal(rgb_e, rgb_o), "per-frame mesh differs"
n_full = sum(e.full_map_mc() for e in R)
assert o.full_map_mc() == n_full
fx_o, fr_o = o.triangles()
fx_e, fr_e = gather(True)
assert fx_e.shape == fx_o.shape and np.array_equal(fx_e.view(np.uint32), fx_o.view(np.uint32)) and np.array_equal(fr_e, fr_o), "full-map mesh differs"
print(f"config {args.config} ({sc.width}x{sc.height}, {cfg['vox_size'] * 1000:g} mm), {args.frames} frames, {args.ranks} rank(s), kernel revisions {args.revs}: "
      f"{len(keys)} blocks, {len(xyz_o)} triangles, full map {n_full} triangles — ALL BIT-EXACT ({time.time() - t0:.0f} s)")
