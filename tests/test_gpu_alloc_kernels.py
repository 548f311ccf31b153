"""GPU (B200): both allocation forms — the one-kernel alloc_visible_kernel (VH_ALLOC_REV=0, default at the reference's 100-step ray
cap on one GPU) and ray_keys_kernel + insert_keys_kernel (VH_ALLOC_REV=2: the DDA as a three-way merge, keys through an inbox;
default for sharded maps and long ray caps) — through the C ABI against the goldens and the oracle."""
import os

import pytest

from test_gpu_parity import assert_triangles_match, assert_voxels_match, run_pair
from util import CASES, engine_params, key_set, load_golden, oracle_params

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["0", "2"], ids=["one-kernel", "keys+insert"])
def alloc1(request, monkeypatch):
    monkeypatch.setenv("VH_ALLOC_REV", request.param)


@pytest.mark.parametrize("name", ["g8_color_holes", "g8_negative_coords"])
def test_alloc_matches_reference_golden(name, vh, synth, alloc1):
    case, g = CASES[name], load_golden(name)
    sc = synth.Scene(**case["scene"])
    color = bool(case["scene"].get("color"))
    with vh.TsdfEngine(engine_params(vh, sc, case)) as eng:
        for i in range(case["frames"]):
            eng.processFrame(*sc.frame(i))
            assert key_set(eng.visible_keys()) == key_set(g[f"visible_{i}"]), f"visible set differs in frame {i}"
        assert key_set(eng.allocated_keys()) == key_set(g["keys"])
        assert_voxels_match(eng, g["keys"], g["sdf"], g["weight"], g["rgb"], color)
        assert_triangles_match(*eng.triangles(), g["tri_xyz"], g["tri_rgb"], color)


def test_alloc_headline_sequence(vh, ob, synth, alloc1):
    sc = synth.make_scene("C2", color=True)
    case = dict(scene=dict(color=True), vpb=8, vox_size=0.005, trunc=0.025, max_depth=10.0)
    run_pair(vh, ob, sc, case, frames=24, num_buckets=1 << 20, pool_blocks=1 << 19, tri_arena_bytes=2 << 30)


def test_alloc_1cm_and_long_rays(vh, ob, synth, alloc1):
    sc = synth.make_scene("C1")
    case = dict(scene={}, vpb=8, vox_size=0.01, trunc=0.05, max_depth=10.0)
    run_pair(vh, ob, sc, case, frames=4, num_buckets=1 << 20, pool_blocks=1 << 19, tri_arena_bytes=512 << 20)
    over = dict(max_ray_steps=400, dda_stride=7)           # 400 steps: 32 KB of crossing times + 38 KB of keys per CTA (opt-in shared-memory size)
    o = ob.Oracle(oracle_params(ob, sc, case, **over))
    with vh.TsdfEngine(engine_params(vh, sc, case, num_buckets=1 << 20, pool_blocks=1 << 20, tri_arena_bytes=1 << 30, **over)) as eng:
        for i in range(2):
            d, rgb, c2w = sc.frame(i)
            o.process_frame(d, None, c2w); eng.processFrame(d, None, c2w)
            assert key_set(eng.visible_keys()) == key_set(o.visible_keys()), f"visible set differs in frame {i}"
            assert eng.stats().voxel_updates == o.last_updates
