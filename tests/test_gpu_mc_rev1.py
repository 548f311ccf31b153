"""GPU (B200): the mesh kernel's emit-pass revision (VH_MC_REV=1: a triangle's six colour gathers are issued before the
first interpolation) against the goldens and the oracle — bit-exact ordered triangle soup. Opt-in and, when written,
checked under CPU emulation only; runs with VH_TEST_REV1=1 (tools/gpu_rev1.sh), see tests/test_gpu_integrate_rev1.py."""
import os

import pytest

from test_gpu_parity import assert_triangles_match, run_pair
from util import CASES, engine_params, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture
def mc1(monkeypatch):
    monkeypatch.setenv("VH_MC_REV", "1")


@pytest.mark.parametrize("name", ["g8_color_holes", "g8_negative_coords"])
def test_mc_rev1_matches_reference_golden(name, vh, synth, mc1):
    case, g = CASES[name], load_golden(name)
    sc = synth.Scene(**case["scene"])
    with vh.TsdfEngine(engine_params(vh, sc, case)) as eng:
        for i in range(case["frames"]):
            eng.processFrame(*sc.frame(i))
        assert_triangles_match(*eng.triangles(), g["tri_xyz"], g["tri_rgb"], bool(case["scene"].get("color")))


def test_mc_rev1_headline_sequence(vh, ob, synth, mc1):
    sc = synth.make_scene("C2", color=True)
    case = dict(scene=dict(color=True), vpb=8, vox_size=0.005, trunc=0.025, max_depth=10.0)
    run_pair(vh, ob, sc, case, frames=24, num_buckets=1 << 20, pool_blocks=1 << 19, tri_arena_bytes=2 << 30)
