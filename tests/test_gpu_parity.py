"""GPU (B200): the CUDA engine through its C ABI against the CPU oracle and the reference-generated goldens.

Bars (BASELINE.json north_star): allocated block set and triangle count exact; per-voxel TSDF and weight within 1e-5
absolute; mesh vertices within 1e-4 voxel. The engine computes in IEEE binary32 in the reference's expression order, so
these tests demand MORE: bit-exact voxels and bit-exact ordered triangles; the tolerances are asserted as well so a
future relaxation of the exactness still has the stated bar written down.
"""
import numpy as np
import pytest

from util import CASES, engine_params, key_set, load_golden, oracle_params, sort_keys

pytestmark = pytest.mark.gpu

TOL_VOXEL = 1e-5          # absolute, sdf and weight
TOL_VERTEX = 1e-4         # in voxel units (triangle soup is in voxel-index units)


def assert_voxels_match(eng, keys, sdf, w, rgb, color):
    s2, w2, c2, found = eng.download_blocks(keys)
    assert found.all()
    assert np.max(np.abs(s2 - sdf), initial=0) <= TOL_VOXEL and np.max(np.abs(w2 - w), initial=0) <= TOL_VOXEL
    assert np.array_equal(s2, sdf), "sdf not bit-exact"
    assert np.array_equal(w2, w), "weight not bit-exact"
    if color:
        assert np.array_equal(c2, rgb), "rgb not exact"


def assert_triangles_match(xyz, rgb, xyz_ref, rgb_ref, color):
    assert xyz.shape == xyz_ref.shape, f"triangle count {len(xyz)} != {len(xyz_ref)}"
    if len(xyz):
        assert np.max(np.abs(xyz - xyz_ref)) <= TOL_VERTEX
    assert np.array_equal(xyz, xyz_ref), "triangle soup not bit-exact / not in tsdf2mesh order"
    if color:
        assert np.array_equal(rgb, rgb_ref)


@pytest.mark.parametrize("name", ["g8_color_holes", "g8_negative_coords"])
def test_engine_matches_reference_golden(name, vh, synth):
    """Fixtures come from the reference's own tsdf.cu run under CPU emulation (tests/golden/make_golden.py)."""
    case, g = CASES[name], load_golden(name)
    sc = synth.Scene(**case["scene"])
    color = bool(case["scene"].get("color"))
    with vh.TsdfEngine(engine_params(vh, sc, case)) as eng:
        for i in range(case["frames"]):
            d, rgb, c2w = sc.frame(i)
            eng.processFrame(d, rgb, c2w)
            assert key_set(eng.visible_keys()) == key_set(g[f"visible_{i}"]), f"visible set differs in frame {i}"
            assert eng.stats().visible_blocks == len(g[f"visible_{i}"])
        assert key_set(eng.allocated_keys()) == key_set(g["keys"])          # allocated block set exact
        assert_voxels_match(eng, g["keys"], g["sdf"], g["weight"], g["rgb"], color)
        xyz, trgb = eng.triangles()
        assert_triangles_match(xyz, trgb, g["tri_xyz"], g["tri_rgb"], color)
        cs = eng.checksum()
        assert cs["sum_w"] == g["checksum"][1] and cs["n_observed"] == g["checksum"][2] and cs["n_negative"] == g["checksum"][3]


def test_general_colour_path_matches_golden(vh, synth, monkeypatch):
    """weights above 4096 switch the integrate kernel from the short exact colour average to the general division
    sequence; VH_INTEGRATE_EXACT_COLOR=1 forces that variant from the first frame."""
    monkeypatch.setenv("VH_INTEGRATE_EXACT_COLOR", "1")
    name = "g8_color_holes"
    case, g = CASES[name], load_golden(name)
    sc = synth.Scene(**case["scene"])
    with vh.TsdfEngine(engine_params(vh, sc, case)) as eng:
        for i in range(case["frames"]):
            eng.processFrame(*sc.frame(i))
        assert_voxels_match(eng, g["keys"], g["sdf"], g["weight"], g["rgb"], True)
        assert_triangles_match(*eng.triangles(), g["tri_xyz"], g["tri_rgb"], True)


def run_pair(vh, ob, sc, case, frames, check_every=1, **eng_over):
    o = ob.Oracle(oracle_params(ob, sc, case))
    color = bool(case["scene"].get("color"))
    with vh.TsdfEngine(engine_params(vh, sc, case, **eng_over)) as eng:
        for i in range(frames):
            d, rgb, c2w = sc.frame(i)
            o.process_frame(d, rgb, c2w)
            eng.processFrame(d, rgb, c2w)
            if i % check_every == 0 or i == frames - 1:
                assert key_set(eng.visible_keys()) == key_set(o.visible_keys()), f"visible set differs in frame {i}"
                st = eng.stats()
                assert st.visible_blocks == o.num_visible
                assert st.voxel_updates == o.last_updates, f"voxel updates differ in frame {i}"
                assert st.triangles == o.last_triangles, f"working-set triangle count differs in frame {i}"
        keys = o.all_keys()
        assert key_set(eng.allocated_keys()) == key_set(keys)
        sdf, w, rgb_, found = o.get_blocks(keys)
        assert_voxels_match(eng, keys, sdf, w, rgb_, color)
        xyz_o, rgb_o = o.triangles()
        xyz, trgb = eng.triangles()
        assert_triangles_match(xyz, trgb, xyz_o, rgb_o, color)
        return eng.stats()


def test_engine_matches_oracle_640x480_1cm(vh, ob, synth):
    """BASELINE config 1 shape (640x480, 1 cm, 8^3, 2^20 buckets x 4), first frames of the 100-frame circle."""
    sc = synth.make_scene("C1")
    case = dict(scene={}, vpb=8, vox_size=0.01, trunc=0.05, max_depth=10.0)
    run_pair(vh, ob, sc, case, frames=4, num_buckets=1 << 20, pool_blocks=1 << 19, tri_arena_bytes=512 << 20)


def test_engine_matches_oracle_640x480_5mm_color(vh, ob, synth):
    """BASELINE config 2 shape (5 mm voxels, headline), colour images, holes in the depth."""
    sc = synth.make_scene("C2", color=True, holes=0.02, spheres=((6.5, 3.0, 1.2, 0.6),))
    case = dict(scene=dict(color=True), vpb=8, vox_size=0.005, trunc=0.025, max_depth=10.0)
    run_pair(vh, ob, sc, case, frames=3, num_buckets=1 << 20, pool_blocks=1 << 19, tri_arena_bytes=512 << 20)


def test_engine_matches_oracle_5mm_sequence(vh, ob, synth):
    """40 consecutive frames of the headline config (the bench workload): per-frame visible set, voxel updates and
    working-set triangle count, then every voxel and the final ordered mesh."""
    sc = synth.make_scene("C2", color=True)
    case = dict(scene=dict(color=True), vpb=8, vox_size=0.005, trunc=0.025, max_depth=10.0)
    run_pair(vh, ob, sc, case, frames=40, num_buckets=1 << 20, pool_blocks=1 << 19, tri_arena_bytes=2 << 30)


def test_engine_matches_oracle_revisit_many_frames(vh, ob, synth):
    """20 frames over a short loop: blocks are re-integrated (weights > 1, harmonic sdf growth) and
    re-meshed ('last frame that saw the block wins', tsdf.cu:534-540)."""
    sc = synth.Scene(width=320, height=240, room=(5.0, 4.0, 2.6), n_frames=40, spheres=((3.9, 2.0, 1.0, 0.5), (1.0, 3.0, 1.6, 0.4)), color=True)
    case = dict(scene=dict(color=True), vpb=8, vox_size=0.02, trunc=0.1, max_depth=3.5)
    run_pair(vh, ob, sc, case, frames=20, check_every=5, num_buckets=1 << 18, pool_blocks=1 << 17)


def test_unbounded_world_and_ray_cap(vh, ob, synth):
    """runtime versions of the reference's macros: no +-64-chunk limit, longer rays, other DDA stride."""
    sc = synth.Scene(width=200, height=150, room=(6.0, 5.0, 2.6), room_min=(-3.0, -2.5, -1.3), n_frames=9)
    case = dict(scene={}, vpb=8, vox_size=0.02, trunc=0.06, max_depth=6.0)
    over = dict(max_chunk_num=0, max_ray_steps=160, dda_stride=7)
    o = ob.Oracle(oracle_params(ob, sc, case, **over))
    with vh.TsdfEngine(engine_params(vh, sc, case, num_buckets=1 << 18, pool_blocks=1 << 17, **over)) as eng:
        for i in range(3):
            d, rgb, c2w = sc.frame(i)
            o.process_frame(d, None, c2w); eng.processFrame(d, None, c2w)
            assert key_set(eng.visible_keys()) == key_set(o.visible_keys())
        keys = o.all_keys()
        sdf, w, rgb_, _ = o.get_blocks(keys)
        assert_voxels_match(eng, keys, sdf, w, rgb_, False)
        assert_triangles_match(*eng.triangles(), *o.triangles(), False)


def test_special_depth_values_and_partial_views(vh, ob, synth):
    """NaN, +-inf, negative, zero, beyond-MaxDepth and sub-millimetre depth pixels, a camera close to a wall (blocks that
    straddle the image border and the near plane) and a narrow truncation: the gates, the whole-block discard and the
    allocation pass must treat every one of them as the reference's comparisons do (NaN passes `dv <= 0 || dv > max`)."""
    sc = synth.Scene(width=320, height=240, room=(3.0, 2.5, 2.4), n_frames=24, radius_frac=0.42, spheres=((2.2, 1.2, 1.1, 0.35),), color=True)
    case = dict(scene=dict(color=True), vpb=8, vox_size=0.01, trunc=0.03, max_depth=2.5)
    o = ob.Oracle(oracle_params(ob, sc, case))
    rng = np.random.RandomState(5)
    with vh.TsdfEngine(engine_params(vh, sc, case, num_buckets=1 << 18, pool_blocks=1 << 18, tri_arena_bytes=256 << 20)) as eng:
        for i in range(6):
            d, rgb, c2w = sc.frame(i)
            d = d.copy()
            r = rng.random_sample(d.shape)
            d[r < 0.01] = np.nan
            d[(r >= 0.01) & (r < 0.02)] = np.inf
            d[(r >= 0.02) & (r < 0.03)] = -np.inf
            d[(r >= 0.03) & (r < 0.04)] = -1.5
            d[(r >= 0.04) & (r < 0.05)] = 7.0            # beyond MaxDepth
            d[(r >= 0.05) & (r < 0.06)] = 1e-4
            d[:16, :48] = np.nan                          # whole tiles of NaN / zero
            d[100:132, 200:232] = 0.0
            o.process_frame(d, rgb, c2w)
            eng.processFrame(d, rgb, c2w)
            st = eng.stats()
            assert key_set(eng.visible_keys()) == key_set(o.visible_keys()), f"visible set differs in frame {i}"
            assert (st.voxel_updates, st.triangles) == (o.last_updates, o.last_triangles), f"frame {i}"
        keys = o.all_keys()
        sdf, w, rgb_, _ = o.get_blocks(keys)
        assert_voxels_match(eng, keys, sdf, w, rgb_, True)
        assert_triangles_match(*eng.triangles(), *o.triangles(), True)


def test_stage_entry_points_with_oracle_visible_list(vh, ob, synth):
    """integrate and marching cubes driven by a visible list produced by the ORACLE."""
    case = CASES["g8_color_holes"]
    sc = synth.Scene(**case["scene"])
    o = ob.Oracle(oracle_params(ob, sc, case))
    with vh.TsdfEngine(engine_params(vh, sc, case)) as eng:
        for i in range(3):
            d, rgb, c2w = sc.frame(i)
            o.process_frame(d, rgb, c2w)
            eng.upload_frame(d, rgb)
            eng.set_visible(o.visible_keys(), c2w)
            assert eng.stats().visible_blocks == o.num_visible
            eng.stage_integrate()
            assert eng.stats().voxel_updates == o.last_updates
            eng.stage_marching_cubes()
            assert eng.stats().triangles == o.last_triangles
        keys = o.all_keys()
        sdf, w, rgb_, _ = o.get_blocks(keys)
        assert_voxels_match(eng, keys, sdf, w, rgb_, True)
        assert_triangles_match(*eng.triangles(), *o.triangles(), True)


def test_async_pipeline_equals_sync(vh, synth):
    """vh_integrate_async back to back (uploads overlapping kernels) gives the same map as the synchronous call."""
    sc = synth.Scene(width=320, height=240, room=(5.0, 4.0, 2.6), n_frames=30, color=True)
    case = dict(scene=dict(color=True), vpb=8, vox_size=0.02, trunc=0.1, max_depth=3.5)
    frames = [sc.frame(i) for i in range(12)]
    with vh.TsdfEngine(engine_params(vh, sc, case)) as a, vh.TsdfEngine(engine_params(vh, sc, case)) as b:
        for d, rgb, c2w in frames:
            a.processFrame(d, rgb, c2w)
        for d, rgb, c2w in frames:
            b.integrate_async(d, rgb, c2w)
        b.sync()
        ka = sort_keys(a.allocated_keys())
        assert np.array_equal(ka, sort_keys(b.allocated_keys()))
        for x, y in zip(a.download_blocks(ka), b.download_blocks(ka)):
            assert np.array_equal(x, y)
        assert np.array_equal(a.triangles()[0], b.triangles()[0])


def test_pinned_host_frames_equal_pageable(vh, synth):
    """pinned caller buffers take a different route (the ray pass reads the depth samples straight from mapped host memory
    while the images upload): same map as with pageable buffers, through both the synchronous and the async call."""
    import ctypes as C
    sc = synth.Scene(width=320, height=240, room=(5.0, 4.0, 2.6), n_frames=30, color=True, holes=0.02)
    case = dict(scene=dict(color=True), vpb=8, vox_size=0.02, trunc=0.1, max_depth=3.5)
    frames = [sc.frame(i) for i in range(6)]
    L = vh.load_library()
    nd, nc = frames[0][0].nbytes, frames[0][1].nbytes
    pd, pc = C.c_void_p(), C.c_void_p()
    assert L.vh_host_alloc(C.byref(pd), nd * len(frames)) == 0 and L.vh_host_alloc(C.byref(pc), nc * len(frames)) == 0
    try:
        for i, (d, rgb, _) in enumerate(frames):
            C.memmove(pd.value + i * nd, d.ctypes.data, nd); C.memmove(pc.value + i * nc, rgb.ctypes.data, nc)
        with vh.TsdfEngine(engine_params(vh, sc, case)) as a, vh.TsdfEngine(engine_params(vh, sc, case)) as b, \
                vh.TsdfEngine(engine_params(vh, sc, case)) as c:
            for i, (d, rgb, c2w) in enumerate(frames):
                a.processFrame(d, rgb, c2w)                                                       # pageable
                assert L.vh_integrate(b.h, pd.value + i * nd, pc.value + i * nc, c2w.ctypes.data) == 0   # pinned, synchronous
                c.integrate_async(pd.value + i * nd, pc.value + i * nc, c2w)                      # pinned, asynchronous
            c.sync()
            keys = sort_keys(a.allocated_keys())
            for other in (b, c):
                assert np.array_equal(keys, sort_keys(other.allocated_keys()))
                for x, y in zip(a.download_blocks(keys), other.download_blocks(keys)):
                    assert np.array_equal(x, y)
                assert np.array_equal(a.triangles()[0], other.triangles()[0])
    finally:
        L.vh_host_free(pd); L.vh_host_free(pc)


def test_u16_depth_ingestion_equals_host_conversion(vh, synth):
    """u16-millimetre depth converted on the GPU gives the map of the host conversion frameLoad performs
    (convertTo(CV_32FC1) then *= 0.001, /root/reference/src/SaveFrame.cpp:174-180)."""
    sc = synth.Scene(width=320, height=240, room=(5.0, 4.0, 2.6), n_frames=30, color=True, holes=0.02)
    case = dict(scene=dict(color=True), vpb=8, vox_size=0.02, trunc=0.1, max_depth=3.5)
    with vh.TsdfEngine(engine_params(vh, sc, case)) as a, vh.TsdfEngine(engine_params(vh, sc, case)) as b:
        keep = []
        for i in range(5):
            d, rgb, c2w = sc.frame(i)
            mm = np.ascontiguousarray(np.round(d.astype(np.float64) * 1000.0).astype(np.uint16))
            host = (mm.astype(np.float32).astype(np.float64) * 0.001).astype(np.float32)
            a.processFrame(host, rgb, c2w)
            keep.append((mm, rgb, c2w))
            b.integrate_u16_async(mm, 0.001, rgb, c2w)
        b.sync()
        keys = sort_keys(a.allocated_keys())
        assert len(keys) > 0 and np.array_equal(keys, sort_keys(b.allocated_keys()))
        for x, y in zip(a.download_blocks(keys), b.download_blocks(keys)):
            assert np.array_equal(x, y)
        assert np.array_equal(a.triangles()[0], b.triangles()[0])


def test_empty_frames_and_errors(vh, synth):
    case = CASES["g8_color_holes"]
    sc = synth.Scene(**case["scene"])
    with vh.TsdfEngine(engine_params(vh, sc, case)) as eng:
        d, rgb, c2w = sc.frame(0)
        eng.processFrame(np.zeros_like(d), rgb, c2w)          # reference would abort on an empty working set (Q8)
        st = eng.stats()
        assert st.visible_blocks == 0 and st.voxel_updates == 0 and st.triangles == 0
        assert len(eng.triangles()[0]) == 0 and len(eng.allocated_keys()) == 0
        eng.processFrame(d, rgb, c2w)
        assert eng.stats().visible_blocks > 0
        eng.reset()
        assert len(eng.allocated_keys()) == 0 and eng.stats().frames == 0
    with pytest.raises(vh.VhError):
        vh.TsdfEngine(engine_params(vh, sc, case, voxels_per_block=5))
    # pool exhaustion is reported, not a hang (reference: operator[] spins forever / "out of block memory")
    with vh.TsdfEngine(engine_params(vh, sc, case, pool_blocks=64)) as eng:
        with pytest.raises(vh.VhError) as ei:
            eng.processFrame(d, rgb, c2w)
        assert ei.value.code == 5


def test_triangle_arena_growth_and_compaction(vh, ob, synth):
    """a deliberately tiny arena: overflow -> compaction/growth -> same mesh as the oracle."""
    sc = synth.Scene(width=320, height=240, room=(5.0, 4.0, 2.6), n_frames=40, spheres=((3.9, 2.0, 1.0, 0.5),))
    case = dict(scene={}, vpb=8, vox_size=0.02, trunc=0.1, max_depth=3.5)
    run_pair(vh, ob, sc, case, frames=8, check_every=4, tri_arena_bytes=48 * 4096, num_buckets=1 << 18, pool_blocks=1 << 17)


def test_full_map_mesh_superset_and_ply(vh, ob, synth, tmp_path):
    case = CASES["g8_color_holes"]
    sc = synth.Scene(**case["scene"])
    with vh.TsdfEngine(engine_params(vh, sc, case)) as eng:
        for i in range(3):
            eng.processFrame(*sc.frame(i))
        ref_xyz, _ = eng.triangles(vh.VH_MESH_REF_PERSISTENT)
        full_xyz, _ = eng.triangles(vh.VH_MESH_FULL_MAP)
        assert len(full_xyz) >= len(ref_xyz) > 0
        verts, faces = eng.weld()
        assert len(faces) == len(ref_xyz) and faces.max() == len(verts) - 1
        # welded vertices reproduce the soup (scaled by vox_size) and are unique
        soup = (ref_xyz.reshape(-1, 3) * np.float32(case["vox_size"]))
        assert np.array_equal(verts["xyz"][faces.reshape(-1)], soup)
        assert len(np.unique(ref_xyz.reshape(-1, 3), axis=0)) == len(verts)
        # numbering = tsdf2mesh's: ids in order of first appearance, the first occurrence's colour (tsdf.cu:1810-1821)
        ref_rgb = eng.triangles(vh.VH_MESH_REF_PERSISTENT)[1].reshape(-1, 3)
        pts = ref_xyz.reshape(-1, 3) + np.float32(0.0)
        uniq, first, inv = np.unique(pts, axis=0, return_index=True, return_inverse=True)
        rank = np.empty(len(uniq), np.int64); rank[np.argsort(first, kind="stable")] = np.arange(len(uniq))
        assert np.array_equal(faces.reshape(-1), rank[inv.reshape(-1)])
        order = np.sort(first)
        assert np.array_equal(verts["xyz"], pts[order] * np.float32(case["vox_size"]))
        assert np.array_equal(verts["rgb"][:, :3], ref_rgb[order])
        ply = tmp_path / "model.ply"
        eng.SavePLY(str(ply))
        head = ply.read_text().split("end_header")[0]
        assert f"element vertex {len(verts)}" in head and f"element face {len(faces)}" in head and "comment stanford bunny" in head
        # binary PLY: the same elements, exact
        bply = tmp_path / "model_bin.ply"
        eng.save_ply_binary(str(bply))
        raw = bply.read_bytes()
        hdr, body = raw.split(b"end_header\n", 1)
        assert b"format binary_little_endian 1.0" in hdr and f"element vertex {len(verts)}".encode() in hdr
        vrec = np.dtype([("xyz", "<f4", 3), ("rgb", "u1", 3)])
        frec = np.dtype([("n", "u1"), ("idx", "<i4", 3)])
        bv = np.frombuffer(body, vrec, len(verts))
        bf = np.frombuffer(body, frec, len(faces), len(verts) * vrec.itemsize)
        assert len(body) == len(verts) * 15 + len(faces) * 13
        assert np.array_equal(bv["xyz"], verts["xyz"]) and np.array_equal(bv["rgb"], verts["rgb"][:, :3])
        assert np.all(bf["n"] == 3) and np.array_equal(bf["idx"], faces)


def test_full_size_properties_5mm(vh, synth):
    """BASELINE config 2 at full size: properties that need no oracle."""
    sc = synth.make_scene("C2")
    p = vh.params_for_scene(sc, vox_size=0.005, trunc_margin=0.025, max_depth=10.0, use_color=0, num_buckets=1 << 20, pool_blocks=1 << 20)
    with vh.TsdfEngine(p) as a, vh.TsdfEngine(p) as b:
        total_updates, seen = 0, set()
        for i in range(6):
            d, rgb, c2w = sc.frame(i)
            a.processFrame(d, None, c2w); b.processFrame(d, None, c2w)
            st = a.stats()
            total_updates += st.voxel_updates
            vis = a.visible_keys()
            assert len(key_set(vis)) == len(vis) == st.visible_blocks            # compacted list has no duplicates
            seen |= key_set(vis)
        assert key_set(a.allocated_keys()) == seen                               # allocated == union of visible sets
        ca, cb = a.checksum(), b.checksum()
        assert all(ca[k] == cb[k] for k in ("sum_w", "n_observed", "n_negative")) and abs(ca["sum_sdf"] - cb["sum_sdf"]) <= 1e-9 * abs(ca["sum_sdf"])   # deterministic map (double sums are order-dependent)
        assert ca["sum_w"] == total_updates                                      # every update adds exactly 1 to one weight
        assert a.stats().voxel_updates_total == total_updates
        keys = sort_keys(a.allocated_keys())[:2000]
        sdf, w, _, _ = a.download_blocks(keys, want_rgb=False)
        assert np.array_equal(w, np.round(w)) and w.min() >= 0
        assert np.all(sdf[w == 0] == 0)                                          # untouched voxels stay zero
        assert np.all(np.abs(sdf[w > 0]) <= np.log(w[w > 0]) + 1.0 + 1e-5)       # harmonic bound of quirk Q2
        assert np.array_equal(a.triangles()[0], b.triangles()[0])
        # re-integrating the same frame again adds exactly the same number of updates (same gate, weights +1)
        d, rgb, c2w = sc.frame(5)
        before = a.stats().voxel_updates
        a.processFrame(d, None, c2w)
        assert a.stats().voxel_updates == before
        assert a.checksum()["sum_w"] == total_updates + before


def test_fast_projection_and_divisions_agree_with_ieee_path(vh, synth, monkeypatch):
    """VH_INTEGRATE_VERIFY=1 makes the integrate kernel evaluate, for every voxel, both its fast path (approximate
    projection + tie guard, shared-reciprocal divisions) and the plain IEEE expressions of the reference, counting any
    disagreement in pixel choice, dist, sdf or colour. Over full-size frames the count must be exactly zero."""
    monkeypatch.setenv("VH_INTEGRATE_VERIFY", "1")
    for name, kw in (("C2", dict(vox_size=0.005, trunc_margin=0.025)), ("C1", dict(vox_size=0.01, trunc_margin=0.05))):
        sc = synth.make_scene(name, color=True, holes=0.01)
        p = vh.params_for_scene(sc, max_depth=10.0, use_color=1, num_buckets=1 << 20, pool_blocks=1 << 20, mc_per_frame=0, **kw)
        with vh.TsdfEngine(p) as eng:
            total = 0
            for i in list(range(0, 12)) + [0, 1, 2, 3]:          # revisits: weights > 1 exercise the shared-reciprocal divisions
                eng.processFrame(*sc.frame(i))
                st = eng.stats()
                assert st.debug_mismatches == 0, f"{name} frame {i}: {st.debug_mismatches} fast/IEEE disagreements"
                total += st.voxel_updates
            assert total > 5e7
