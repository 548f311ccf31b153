#!/usr/bin/env python3
"""Headline-shape check of the opt-in kernel revisions WITHOUT a GPU: BASELINE config 2 frames (640x480, 5 mm, colour, 1 % holes)
through the emulated kernel sources (tests/emu) with VH_INTEGRATE_REV / VH_ALLOC_REV / VH_MC_REV = 1, against the oracle, bit for bit.
~1 minute per frame. usage: tools/emu_headline_check.py [frames=3] [integrate_rev alloc_rev mc_rev = 1 1 1]"""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np
vh = importlib.import_module("voxel-hashing-sdf_b200")
synth = importlib.import_module("voxel-hashing-sdf_b200.synth")
from oracle import binding as ob
from emu.binding import EmuEngine, mesh_order
from util import engine_params, oracle_params, key_set
sc = synth.make_scene("C2", color=True, holes=0.01)
case = dict(scene=dict(color=True), vpb=8, vox_size=0.005, trunc=0.025, max_depth=10.0)
o = ob.Oracle(oracle_params(ob, sc, case))
NFRAMES = int(sys.argv[1]) if len(sys.argv) > 1 else 3
REVS = [int(x) for x in sys.argv[2:5]] if len(sys.argv) > 4 else [1, 1, 1]
t0 = time.time()
with EmuEngine(engine_params(vh, sc, case, num_buckets=1 << 18, pool_blocks=1 << 17, tri_arena_bytes=256 << 20), integrate_rev=REVS[0], alloc_rev=REVS[1], mc_rev=REVS[2]) as e:
    for i in range(NFRAMES):
        d, rgb, c2w = sc.frame(i)
        o.process_frame(d, rgb, c2w)
        e.process_frame(d, rgb, c2w)
        assert key_set(e.visible_keys()) == key_set(o.visible_keys()), f"visible {i}"
        assert e.last_updates == o.last_updates, (i, e.last_updates, o.last_updates)
        assert e.last_triangles == o.last_triangles, (i, e.last_triangles, o.last_triangles)
        print("frame", i, "visible", e.num_visible, "updates", e.last_updates, "tris", e.last_triangles, f"{time.time()-t0:.0f}s", flush=True)
    keys = o.all_keys()
    so, wo, co, _ = o.get_blocks(keys)
    se, we, ce, found, neg = e.get_blocks(keys)
    assert found.all() and np.array_equal(se.view(np.uint32), so.view(np.uint32)) and np.array_equal(we, wo) and np.array_equal(ce, co)
    assert np.array_equal(neg, (se < 0).sum(1))
    xyz_o, rgb_o = o.triangles()
    xyz_e, rgb_e = e.block_triangles(mesh_order(keys))
    assert xyz_e.shape == xyz_o.shape and np.array_equal(xyz_e.view(np.uint32), xyz_o.view(np.uint32)) and np.array_equal(rgb_e, rgb_o)
    print(f"C2-shape 640x480 5 mm, {NFRAMES} frames, kernel revisions {REVS} under emulation: blocks", len(keys), "triangles", len(xyz_o), "ALL BIT-EXACT", f"{time.time()-t0:.0f}s")
