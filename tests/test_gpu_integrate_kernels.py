"""GPU (B200): both integrate kernels (integrate_kernel_staged: planes through shared memory by bulk async copies, the default;
integrate_kernel_direct: per-lane plane loads) at their launch shapes, through the C ABI against the goldens and the oracle."""
import os

import numpy as np
import pytest

from test_gpu_parity import assert_triangles_match, assert_voxels_match, run_pair
from util import CASES, engine_params, load_golden

pytestmark = pytest.mark.gpu


# (VH_INTEGRATE_REV, VH_INTEGRATE_PARTS, VH_INTEGRATE_CTAS): 2 = staged (cp.async.bulk + mbarrier; work items = whole blocks or x-halves,
# chosen per frame by the host when not forced), 1 = direct (per-lane plane loads after the gate)
@pytest.fixture(params=[("2", "2", "5"), ("2", "1", "4"), ("2", "2", "4"), ("2", "", ""), ("1", "", "3"), ("1", "", "4")],
                ids=["staged-halves-5ctas", "staged-whole-4ctas", "staged-halves-4ctas", "staged-auto", "direct-3ctas", "direct-4ctas"])
def kernel(request, monkeypatch):
    monkeypatch.setenv("VH_INTEGRATE_REV", request.param[0])
    for name, v in (("VH_INTEGRATE_PARTS", request.param[1]), ("VH_INTEGRATE_CTAS", request.param[2])):
        if v:
            monkeypatch.setenv(name, v)
        else:
            monkeypatch.delenv(name, raising=False)


@pytest.mark.parametrize("name", ["g8_color_holes", "g8_negative_coords"])
def test_integrate_matches_reference_golden(name, vh, synth, kernel):
    case, g = CASES[name], load_golden(name)
    sc = synth.Scene(**case["scene"])
    color = bool(case["scene"].get("color"))
    with vh.TsdfEngine(engine_params(vh, sc, case)) as eng:
        for i in range(case["frames"]):
            eng.processFrame(*sc.frame(i))
        assert_voxels_match(eng, g["keys"], g["sdf"], g["weight"], g["rgb"], color)
        assert_triangles_match(*eng.triangles(), g["tri_xyz"], g["tri_rgb"], color)
        cs = eng.checksum()
        assert cs["sum_w"] == g["checksum"][1] and cs["n_observed"] == g["checksum"][2] and cs["n_negative"] == g["checksum"][3]


def test_integrate_general_colour_path(vh, synth, kernel, monkeypatch):
    monkeypatch.setenv("VH_INTEGRATE_EXACT_COLOR", "1")     # weights "above 65536": the general division sequence
    name = "g8_color_holes"
    case, g = CASES[name], load_golden(name)
    sc = synth.Scene(**case["scene"])
    with vh.TsdfEngine(engine_params(vh, sc, case)) as eng:
        for i in range(case["frames"]):
            eng.processFrame(*sc.frame(i))
        assert_voxels_match(eng, g["keys"], g["sdf"], g["weight"], g["rgb"], True)


def test_integrate_headline_sequence(vh, ob, synth, kernel):
    """40 frames of BASELINE config 2 (the bench workload): per-frame counters, every voxel, the ordered mesh."""
    sc = synth.make_scene("C2", color=True)
    case = dict(scene=dict(color=True), vpb=8, vox_size=0.005, trunc=0.025, max_depth=10.0)
    run_pair(vh, ob, sc, case, frames=40, num_buckets=1 << 20, pool_blocks=1 << 19, tri_arena_bytes=2 << 30)


def test_integrate_revisits_and_hostile_depth(vh, ob, synth, kernel):
    """weights > 1 and NaN / inf / negative / denormal depth samples (the out-of-line IEEE redo of a step)"""
    sc = synth.Scene(width=320, height=240, room=(5.0, 4.0, 2.6), n_frames=40, spheres=((3.9, 2.0, 1.0, 0.5),), color=True)
    case = dict(scene=dict(color=True), vpb=8, vox_size=0.02, trunc=0.1, max_depth=3.5)
    o = ob.Oracle(__import__("util").oracle_params(ob, sc, case))
    with vh.TsdfEngine(engine_params(vh, sc, case, num_buckets=1 << 18, pool_blocks=1 << 17)) as eng:
        for i in range(12):
            d, rgb, c2w = sc.frame(i)
            rng = np.random.RandomState(11 + i)
            bad = np.array([np.nan, np.inf, -np.inf, -1.0, 1e-40, 50.0, 0.0], np.float32)
            idx = rng.randint(0, d.size, 2000)
            d.reshape(-1)[idx] = bad[rng.randint(0, len(bad), 2000)]
            o.process_frame(d, rgb, c2w)
            eng.processFrame(d, rgb, c2w)
            assert eng.stats().voxel_updates == o.last_updates, f"voxel updates differ in frame {i}"
        keys = o.all_keys()
        sdf, w, rgb_, _ = o.get_blocks(keys)
        s2, w2, c2, found = eng.download_blocks(keys)
        assert found.all()
        assert np.array_equal(s2.view(np.uint32), sdf.view(np.uint32)) and np.array_equal(w2.view(np.uint32), w.view(np.uint32))
        assert np.array_equal(c2, rgb_)


def test_integrate_verify_mode_counts_no_disagreement(vh, synth, kernel, monkeypatch):
    """VH_INTEGRATE_VERIFY=1: fast and IEEE formulations side by side inside the kernel, zero disagreements"""
    monkeypatch.setenv("VH_INTEGRATE_VERIFY", "1")
    sc = synth.make_scene("C2", color=True)
    case = dict(scene=dict(color=True), vpb=8, vox_size=0.005, trunc=0.025, max_depth=10.0)
    with vh.TsdfEngine(engine_params(vh, sc, case, num_buckets=1 << 20, pool_blocks=1 << 19, tri_arena_bytes=1 << 30)) as eng:
        total = 0
        for i in range(6):
            eng.processFrame(*sc.frame(i))
            st = eng.stats()
            total += st.voxel_updates
            assert st.debug_mismatches == 0
        assert total > 5e7
