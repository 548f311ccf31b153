"""GPU (B200): the marching-cubes mesh kernel with and without the colour tile staged in shared memory (VH_MC_COLOR_TILE), through
the C ABI against the goldens and the oracle (ordered triangle soup incl. vertex colours, bit for bit)."""
import pytest

from test_gpu_parity import assert_triangles_match, assert_voxels_match, run_pair
from util import CASES, engine_params, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["0", "1"], ids=["colour-gathers", "colour-tile"])
def mesh_kernel(request, monkeypatch):
    monkeypatch.setenv("VH_MC_COLOR_TILE", request.param)


@pytest.mark.parametrize("name", ["g8_color_holes", "g8_negative_coords"])
def test_mc_matches_reference_golden(name, vh, synth, mesh_kernel):
    case, g = CASES[name], load_golden(name)
    sc = synth.Scene(**case["scene"])
    color = bool(case["scene"].get("color"))
    with vh.TsdfEngine(engine_params(vh, sc, case)) as eng:
        for i in range(case["frames"]):
            eng.processFrame(*sc.frame(i))
        assert_voxels_match(eng, g["keys"], g["sdf"], g["weight"], g["rgb"], color)
        assert_triangles_match(*eng.triangles(), g["tri_xyz"], g["tri_rgb"], color)
        full = eng.triangles(vh.VH_MESH_FULL_MAP)
        assert len(full[0]) >= len(g["tri_xyz"])


def test_mc_headline_sequence(vh, ob, synth, mesh_kernel):
    """24 frames of BASELINE config 2 with colour: per-frame triangle counts, every voxel, the ordered coloured mesh"""
    sc = synth.make_scene("C2", color=True)
    case = dict(scene=dict(color=True), vpb=8, vox_size=0.005, trunc=0.025, max_depth=10.0)
    run_pair(vh, ob, sc, case, frames=24, num_buckets=1 << 20, pool_blocks=1 << 19, tri_arena_bytes=2 << 30)
