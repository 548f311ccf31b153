"""GPU (B200): BASELINE configs 3 and 4 as parity cases (BASELINE.json: "the other configs are parity-test cases").
Config 3: 1280x720, 4 mm voxels, 2^24-bucket hash, truncation 3 cm, full-map mesh extraction.
Config 4: 10 m room centred on the origin (negative block coordinates), 2 mm voxels, truncation 1 cm, at the reference's 100 ray
steps and with the step cap scaled to the block size (1,100: the two-kernel allocation form). First green on B200 in profiles/r02a."""
import os

import numpy as np
import pytest

from test_gpu_parity import assert_triangles_match, assert_voxels_match
from util import engine_params, key_set, oracle_params

pytestmark = pytest.mark.gpu


def run_config(vh, ob, synth, name, frames, over, full_map=False, **eng):
    cfg = synth.CONFIGS[name]
    sc = synth.make_scene(name, color=True)
    case = dict(scene=dict(color=True), vpb=8, vox_size=cfg["vox_size"], trunc=cfg["trunc"], max_depth=cfg["max_depth"])
    o = ob.Oracle(oracle_params(ob, sc, case, **over))
    with vh.TsdfEngine(engine_params(vh, sc, case, num_buckets=cfg["num_buckets"], **over, **eng)) as e:
        for i in range(frames):
            d, rgb, c2w = sc.frame(i)
            o.process_frame(d, rgb, c2w)
            e.processFrame(d, rgb, c2w)
            assert key_set(e.visible_keys()) == key_set(o.visible_keys()), f"{name}: visible set differs in frame {i}"
            st = e.stats()
            assert st.voxel_updates == o.last_updates and st.triangles == o.last_triangles, f"{name}: frame {i}"
        keys = o.all_keys()
        assert key_set(e.allocated_keys()) == key_set(keys)
        sdf, w, rgb_, _ = o.get_blocks(keys)
        assert_voxels_match(e, keys, sdf, w, rgb_, True)
        assert_triangles_match(*e.triangles(), *o.triangles(), True)
        if full_map:      # every allocated block re-meshed against the whole map, exact against the oracle's full-map pass (destructive there: last)
            n_ref = len(e.triangles()[0])
            full = e.triangles(vh.VH_MESH_FULL_MAP)
            assert o.full_map_mc() == len(full[0]) >= n_ref
            assert_triangles_match(*full, *o.triangles(), True)
        return e.stats()


def test_config3_1280x720_4mm_full_map_mesh(vh, ob, synth):
    run_config(vh, ob, synth, "C3", 3, {}, full_map=True, pool_blocks=1 << 19, tri_arena_bytes=2 << 30)


def test_config4_room_scale_2mm_negative_coordinates(vh, ob, synth):
    # 100 ray steps (the reference's value) and the step cap scaled to the block size
    run_config(vh, ob, synth, "C4", 2, {}, pool_blocks=1 << 19, tri_arena_bytes=1 << 30)
    run_config(vh, ob, synth, "C4", 1, dict(max_ray_steps=1100), pool_blocks=1 << 21, tri_arena_bytes=2 << 30)
