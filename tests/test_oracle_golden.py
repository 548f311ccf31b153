"""CPU: the oracle (oracle/vh_oracle.c) against fixtures produced by the reference's own tsdf.cu under CPU
emulation (tests/golden/make_golden.py). Bit-exact: visible lists in order, every stored voxel, ordered triangles."""
import numpy as np
import pytest

from util import CASES, load_golden, oracle_params


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_golden(name, ob, synth):
    case, g = CASES[name], load_golden(name)
    sc = synth.Scene(**case["scene"])
    o = ob.Oracle(oracle_params(ob, sc, case))
    for i in range(case["frames"]):
        d, rgb, c2w = sc.frame(i)
        o.process_frame(d, rgb, c2w)
        assert np.array_equal(o.visible_keys(), g[f"visible_{i}"]), f"visible list differs in frame {i}"
        assert o.streamed_blocks == int(g["streamed"][i])
    sdf, w, rgb, found = o.get_blocks(g["keys"])
    assert found.all()
    assert np.array_equal(sdf, g["sdf"])            # bit-exact float32
    assert np.array_equal(w, g["weight"])
    assert np.array_equal(rgb, g["rgb"])
    xyz, trgb = o.triangles()
    assert xyz.shape == g["tri_xyz"].shape           # triangle count exact
    assert np.array_equal(xyz, g["tri_xyz"])         # same order, same bits
    assert np.array_equal(trgb, g["tri_rgb"])
    cs = o.checksum()
    assert cs["sum_w"] == g["checksum"][1] and cs["n_observed"] == g["checksum"][2] and cs["n_negative"] == g["checksum"][3]
    # every block the oracle ever allocated beyond the golden key list was never visible -> must not exist
    assert len(o.all_keys()) == len(g["keys"])


def test_oracle_stage_api_equals_process_frame(ob, synth):
    case = CASES["g8_color_holes"]
    sc = synth.Scene(**case["scene"])
    a = ob.Oracle(oracle_params(ob, sc, case))
    b = ob.Oracle(oracle_params(ob, sc, case))
    for i in range(2):
        d, rgb, c2w = sc.frame(i)
        a.process_frame(d, rgb, c2w)
        b.begin_frame(c2w); b.stage_allocate(d); b.stage_integrate(d, rgb); b.stage_mc()
        assert a.last_updates == b.last_updates and a.last_triangles == b.last_triangles
    assert a.checksum() == b.checksum()
    assert np.array_equal(a.triangles()[0], b.triangles()[0])


def test_oracle_threads_do_not_change_results(ob, synth):
    case = CASES["g8_negative_coords"]
    sc = synth.Scene(**case["scene"])
    a = ob.Oracle(oracle_params(ob, sc, case, num_threads=1))
    b = ob.Oracle(oracle_params(ob, sc, case, num_threads=0))
    for i in range(2):
        d, rgb, c2w = sc.frame(i)
        a.process_frame(d, rgb, c2w); b.process_frame(d, rgb, c2w)
    assert a.checksum() == b.checksum()
    assert np.array_equal(a.triangles()[0], b.triangles()[0])


def test_empty_and_invalid_depth(ob, synth):
    """all-zero depth: no ray passes the gate, nothing becomes visible (the reference would abort)."""
    case = CASES["g8_color_holes"]
    sc = synth.Scene(**case["scene"])
    o = ob.Oracle(oracle_params(ob, sc, case))
    d, rgb, c2w = sc.frame(0)
    o.process_frame(np.zeros_like(d), rgb, c2w)
    assert o.num_visible == 0 and o.last_updates == 0 and len(o.triangles()[0]) == 0
    o.process_frame(np.full_like(d, 50.0), rgb, c2w)     # beyond MaxDepth: gated at tsdf.cu:2119
    assert o.num_visible == 0
