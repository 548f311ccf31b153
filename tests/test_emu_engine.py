"""CPU: the engine's whole per-frame path — frame pack, visible-block allocation (ray DDA, lock-free hash insertion,
warp-aggregated slot pops), integration and both marching-cubes kernels — compiled FROM THE PRODUCT'S KERNEL SOURCES for
the host (tests/emu) and checked against the oracle with the bars of the GPU parity tests: visible set per frame, voxel
updates, triangle counts, every voxel bit for bit, and the ordered triangle soup bit for bit.
The GPU parity tests (-m gpu) remain the authority for the compiled SASS; this file guards the kernel logic on machines
without a GPU.
"""
import numpy as np
import pytest

from emu.binding import EmuEngine, mesh_order
from util import CASES, engine_params, key_set, load_golden, oracle_params


def run_pair(vh, ob, synth, scene_kw, case, frames, rev=0, alloc_rev=0, mc_rev=0, mutate=None, **eng_over):
    sc = synth.Scene(**scene_kw)
    color = bool(case["scene"].get("color"))
    o = ob.Oracle(oracle_params(ob, sc, case))
    with EmuEngine(engine_params(vh, sc, case, **eng_over), integrate_rev=rev, alloc_rev=alloc_rev, mc_rev=mc_rev) as e:
        for i in range(frames):
            d, rgb, c2w = sc.frame(i)
            if mutate is not None:
                d = mutate(i, d)
            o.process_frame(d, rgb, c2w)
            e.process_frame(d, rgb, c2w)
            assert key_set(e.visible_keys()) == key_set(o.visible_keys()), f"visible set differs in frame {i}"
            assert e.num_visible == o.num_visible
            assert e.last_updates == o.last_updates, f"voxel updates differ in frame {i}"
            assert e.last_triangles == o.last_triangles, f"working-set triangle count differs in frame {i}"
        keys = o.all_keys()
        assert key_set(e.all_keys()) == key_set(keys)
        so, wo, co, _ = o.get_blocks(keys)
        se, we, ce, found, neg = e.get_blocks(keys)
        assert found.all()
        assert np.array_equal(se.view(np.uint32), so.view(np.uint32)), "sdf not bit-exact"
        assert np.array_equal(we.view(np.uint32), wo.view(np.uint32)), "weight not bit-exact"
        if color:
            assert np.array_equal(ce, co), "rgb not exact"
        assert np.array_equal(neg, (se < 0).sum(1))
        xyz_o, rgb_o = o.triangles()
        xyz_e, rgb_e = e.block_triangles(mesh_order(keys))
        assert xyz_e.shape == xyz_o.shape, f"triangle count {len(xyz_e)} != {len(xyz_o)}"
        assert np.array_equal(xyz_e.view(np.uint32), xyz_o.view(np.uint32)), "triangle soup not bit-exact / not in tsdf2mesh order"
        if color:
            assert np.array_equal(rgb_e, rgb_o)
        return len(keys), len(xyz_o)


SMALL = dict(width=160, height=120, room=(4.0, 3.0, 2.5), n_frames=60, spheres=((2.8, 1.5, 1.0, 0.4),), color=True)
CASE = dict(scene=dict(color=True), vpb=8, vox_size=0.04, trunc=0.2, max_depth=3.0)


# (integrate kernel: 1 direct, 2 staged; allocation: 0 one kernel, 2 keys + insert; marching-cubes mesh kernel: 1 = colour tile through shared memory)
@pytest.mark.parametrize("rev,alloc_rev,mc_rev", [(1, 0, 0), (2, 0, 1), (1, 2, 1), (2, 2, 0)])
def test_emulated_engine_matches_oracle(vh, ob, synth, rev, alloc_rev, mc_rev):
    nblocks, ntris = run_pair(vh, ob, synth, SMALL, CASE, frames=3, rev=rev, alloc_rev=alloc_rev, mc_rev=mc_rev, num_buckets=1 << 12,
                              pool_blocks=1 << 12, tri_arena_bytes=8 << 20)
    assert nblocks > 200 and ntris > 1000


@pytest.mark.parametrize("alloc_rev", [0, 2])
def test_emulated_engine_negative_coordinates_no_colour(vh, ob, synth, alloc_rev):
    sc = dict(width=160, height=120, room=(4.0, 3.0, 2.5), room_min=(-2.0, -1.5, -1.25), n_frames=60, holes=0.02)
    case = dict(scene={}, vpb=8, vox_size=0.04, trunc=0.2, max_depth=3.0)
    run_pair(vh, ob, synth, sc, case, frames=3, alloc_rev=alloc_rev, num_buckets=1 << 12, pool_blocks=1 << 12, tri_arena_bytes=8 << 20)


def test_emulated_engine_tiny_table_probes_and_wraps(vh, ob, synth):
    """a table barely larger than the block count: long linear-probe runs that wrap around the end of the table"""
    run_pair(vh, ob, synth, SMALL, CASE, frames=2, num_buckets=128, entries_per_bucket=4, pool_blocks=600, tri_arena_bytes=8 << 20)


@pytest.mark.parametrize("name", ["g8_color_holes"])
def test_emulated_engine_matches_reference_golden(name, vh, synth):
    """the fixture generated from the reference's own tsdf.cu (tests/golden/make_golden.py)"""
    case, g = CASES[name], load_golden(name)
    sc = synth.Scene(**case["scene"])
    with EmuEngine(engine_params(vh, sc, case)) as e:
        for i in range(case["frames"]):
            e.process_frame(*sc.frame(i))
            assert key_set(e.visible_keys()) == key_set(g[f"visible_{i}"]), f"visible set differs in frame {i}"
        assert key_set(e.all_keys()) == key_set(g["keys"])
        s, w, c, found, _ = e.get_blocks(g["keys"])
        assert found.all() and np.array_equal(s, g["sdf"]) and np.array_equal(w, g["weight"]) and np.array_equal(c, g["rgb"])
        xyz, trgb = e.block_triangles(mesh_order(g["keys"]))
        assert np.array_equal(xyz, g["tri_xyz"]) and np.array_equal(trgb, g["tri_rgb"])


@pytest.mark.parametrize("nranks,group,alloc_rev", [(2, 8, 0), (3, 1, 0), (2, 8, 2), (3, 1, 2), (4, 2, 2)])
def test_emulated_sharded_map_equals_single_map(vh, ob, synth, nranks, group, alloc_rev):
    """one map sharded by block-coordinate hash over N emulated ranks (ownership filter in the allocation kernel, remote
    table probes and halo reads in both marching-cubes kernels): the union is the oracle's single map, bit for bit"""
    from emu.binding import EmuGroup
    sc = synth.Scene(**SMALL)
    o = ob.Oracle(oracle_params(ob, sc, CASE))

    def make(rank, n):
        return engine_params(vh, sc, CASE, num_buckets=1 << 12, pool_blocks=1 << 12, tri_arena_bytes=8 << 20, shard_rank=rank, shard_count=n,
                             shard_group=group)

    # allocation revision 2: the rays are split across the ranks and every key travels to its owner's inbox before the frame barrier
    with EmuGroup(make, nranks, mc_rev=nranks % 2, alloc_rev=alloc_rev, integrate_rev=1 + nranks % 2) as g:
        for i in range(3):
            d, rgb, c2w = sc.frame(i)
            o.process_frame(d, rgb, c2w)
            g.process_frame(d, rgb, c2w)
            vis = [key_set(e.visible_keys()) for e in g.ranks]
            assert sum(len(v) for v in vis) == o.num_visible and set().union(*vis) == key_set(o.visible_keys()), f"frame {i}: visible sets"
            assert sum(e.last_updates for e in g.ranks) == o.last_updates
            assert sum(e.last_triangles for e in g.ranks) == o.last_triangles, f"frame {i}: triangle counts"
        assert all(e.num_blocks > 0 for e in g.ranks), "a rank owns nothing: the case does not exercise sharding"
        keys = o.all_keys()
        owner = np.array([vh.owner_of_block(int(k[0]), int(k[1]), int(k[2]), nranks, group) for k in keys])
        so, wo, co, _ = o.get_blocks(keys)
        for r, e in enumerate(g.ranks):
            mine = keys[owner == r]
            assert key_set(e.all_keys()) == key_set(mine), f"rank {r} holds blocks it does not own (or misses some)"
            s, w, c, found, neg = e.get_blocks(mine)
            assert found.all()
            assert np.array_equal(s.view(np.uint32), so[owner == r].view(np.uint32)) and np.array_equal(w, wo[owner == r]) and np.array_equal(c, co[owner == r])
        # gathered mesh: blocks in tsdf2mesh order, each fetched from its owner
        xyz_o, rgb_o = o.triangles()
        parts = []
        for k in mesh_order(keys):
            r = vh.owner_of_block(int(k[0]), int(k[1]), int(k[2]), nranks, group)
            parts.append(g.ranks[r].block_triangles(k[None, :]))
        xyz = np.concatenate([p[0] for p in parts]) if parts else np.zeros((0, 3, 3), np.float32)
        trgb = np.concatenate([p[1] for p in parts]) if parts else np.zeros((0, 3, 3), np.uint8)
        assert xyz.shape == xyz_o.shape and np.array_equal(xyz.view(np.uint32), xyz_o.view(np.uint32)) and np.array_equal(trgb, rgb_o)


@pytest.mark.parametrize("alloc_rev", [0, 2])
def test_emulated_engine_other_launch_shapes(vh, ob, synth, alloc_rev):
    """non-default run-time values of the reference's macros: DDA stride 7, 160 ray steps (a larger dynamic shared-memory
    carve-out in the allocation kernel), unbounded chunk world, 16:9 image"""
    sc = dict(width=192, height=108, room=(5.0, 4.0, 2.6), room_min=(-2.5, -2.0, -1.3), n_frames=60, color=True)
    case = dict(scene=dict(color=True), vpb=8, vox_size=0.03, trunc=0.09, max_depth=5.0)
    over = dict(max_chunk_num=0, max_ray_steps=160, dda_stride=7)
    s = synth.Scene(**sc)
    o = ob.Oracle(oracle_params(ob, s, case, **over))
    with EmuEngine(engine_params(vh, s, case, num_buckets=1 << 12, pool_blocks=1 << 13, tri_arena_bytes=16 << 20, **over), alloc_rev=alloc_rev) as e:
        for i in range(2):
            d, rgb, c2w = s.frame(i)
            o.process_frame(d, rgb, c2w)
            e.process_frame(d, rgb, c2w)
            assert key_set(e.visible_keys()) == key_set(o.visible_keys())
            assert e.last_updates == o.last_updates and e.last_triangles == o.last_triangles
        keys = o.all_keys()
        so, wo, co, _ = o.get_blocks(keys)
        se, we, ce, found, _ = e.get_blocks(keys)
        assert found.all() and np.array_equal(se, so) and np.array_equal(we, wo) and np.array_equal(ce, co)
        xyz_o, _ = o.triangles()
        xyz_e, _ = e.block_triangles(mesh_order(keys))
        assert np.array_equal(xyz_e, xyz_o)


def test_emulated_engine_pool_and_table_exhaustion_raise_flags(vh, synth):
    """the reference throws "out of block memory" (blockalloc.h:51) / asserts on a full table (vhashing.h:104-112); the
    kernels raise sticky error bits instead (the C ABI turns them into VH_ERR_*), keep running and never write out of bounds"""
    sc = synth.Scene(**SMALL)
    d, rgb, c2w = sc.frame(0)
    with EmuEngine(engine_params(vh, sc, CASE, num_buckets=1 << 12, pool_blocks=40, tri_arena_bytes=8 << 20)) as e:      # ~200 blocks wanted
        rc = e.process_frame(d, rgb, c2w, check=False)
        assert rc & 2, "MAP_POOL_FULL not raised"
        assert e.num_blocks >= 40
    with EmuEngine(engine_params(vh, sc, CASE, num_buckets=16, entries_per_bucket=4, pool_blocks=1 << 12, tri_arena_bytes=8 << 20)) as e:   # capacity clamps to 1024
        rc = 0
        for i in range(0, 40, 4):
            d, rgb, c2w = sc.frame(i)
            rc |= e.process_frame(d, rgb, c2w, check=False)
        assert rc & 1, "MAP_TABLE_FULL not raised"


@pytest.mark.parametrize("alloc_rev", [0, 2])
def test_emulated_allocation_dda_ties(vh, ob, synth, alloc_rev):
    """axis-aligned poses with the camera on block boundaries: for the pixels on the image diagonals two axes of the DDA
    cross block faces at exactly the same ray parameter, so the reference's tie rule decides which block is visited
    (tsdf.cu:2217-2233: x only if strictly smallest, then z only if strictly before y). The step-by-step kernel and the
    merge formulation of revision 1 must both reproduce the oracle's visible sets."""
    sc = synth.Scene(width=160, height=120, room=(4.0, 3.0, 2.5), n_frames=4, radius_frac=0.0, color=False)
    case = dict(scene={}, vpb=8, vox_size=0.03125, trunc=0.1, max_depth=3.0)      # 0.25 m blocks: the camera (2, 1.5, 1.25) sits on block faces
    o = ob.Oracle(oracle_params(ob, sc, case))
    ties = 0
    with EmuEngine(engine_params(vh, sc, case, num_buckets=1 << 13, pool_blocks=1 << 13, tri_arena_bytes=16 << 20, mc_per_frame=0), alloc_rev=alloc_rev) as e:
        for i in range(4):
            d, rgb, c2w = sc.frame(i)
            o.begin_frame(c2w); o.stage_allocate(d)
            e.process_frame(d, None, c2w)
            assert key_set(e.visible_keys()) == key_set(o.visible_keys()), f"visible set differs in frame {i}"
            o.stage_integrate(d, None)
            # count rays with an exact tie between two axes' first crossing times, to be sure the case bites
            m = c2w.reshape(4, 4)
            for v in range(0, 120, 10):
                for u in range(0, 160, 10):
                    dirc = m[:3, :3] @ np.array([(u - sc.cx) / sc.fx, (v - sc.cy) / sc.fy, 1.0])
                    a = np.abs(dirc)
                    ties += int(np.sum(np.isclose(a[:, None], a[None, :], rtol=0, atol=1e-9)) > 3)
    assert ties > 20


def test_emulated_dda_merge_equals_step_by_step_march(vh, synth):
    """allocation revision 1 replaces the sequential 3-D DDA by a three-way merge of per-axis crossing times; here the two
    are compared key by key, step by step, for every sampled ray — generic poses and the axis-aligned tie cases above"""
    from emu.binding import compare_march
    total = 0
    for scene_kw, vox in ((SMALL, 0.04), (dict(width=160, height=120, room=(4.0, 3.0, 2.5), n_frames=4, radius_frac=0.0), 0.03125),
                          (dict(width=192, height=108, room=(5.0, 4.0, 2.6), room_min=(-2.5, -2.0, -1.3), n_frames=7), 0.01)):
        sc = synth.Scene(**scene_kw)
        case = dict(scene={}, vpb=8, vox_size=vox, trunc=3 * vox, max_depth=3.0)
        for steps in (100, 37):
            p = engine_params(vh, sc, case, max_ray_steps=steps)
            for i in range(min(sc.n_frames, 5)):
                d, _, c2w = sc.frame(i)
                bad, n, nonempty = compare_march(p, d, c2w)
                assert bad == 0, f"{bad} of {n} DDA steps differ (vox {vox}, frame {i}, {steps} steps)"
                total += nonempty
    assert total > 50000


def _random_pose_scene(synth, seed, **kw):
    """a Scene whose poses are random rigid motions (any pitch / roll / yaw, anywhere in the room) instead of the horizontal
    circle: projections with all nine rotation entries non-trivial, blocks that straddle the camera plane, views of the
    floor and the ceiling"""
    rng = np.random.RandomState(seed)

    class RandomPoses(synth.Scene):
        def pose(self, i):
            r = np.random.RandomState(seed * 1000 + i)
            q = r.normal(size=4); q /= np.linalg.norm(q)
            w, x, y, z = q
            rot = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                            [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                            [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
            lo = np.array(self.room_min); ext = np.array(self.room)
            pos = lo + ext * (0.25 + 0.5 * r.random_sample(3))
            m = np.eye(4); m[:3, :3] = rot; m[:3, 3] = pos
            return m.astype(np.float32).reshape(16)

    del rng
    return RandomPoses(**kw)


@pytest.mark.parametrize("seed,revs", [(1, (1, 0, 0)), (2, (1, 2, 1)), (3, (2, 0, 1)), (4, (1, 0, 0)), (5, (2, 2, 0)), (6, (1, 2, 1))])
def test_emulated_engine_random_rigid_poses(vh, ob, synth, seed, revs):
    """odd image size (not a multiple of the 16-pixel tiles or the 10-pixel ray stride), off-centre principal point,
    random rigid poses; three frames that overlap only by chance"""
    sc = _random_pose_scene(synth, seed, width=150, height=113, room=(3.0, 2.6, 2.2), room_min=(-1.2, -0.9, -0.4), n_frames=8, color=True, holes=0.03)
    sc.cx, sc.cy = 71.3, 59.8
    case = dict(scene=dict(color=True), vpb=8, vox_size=0.025, trunc=0.08, max_depth=2.5)
    o = ob.Oracle(oracle_params(ob, sc, case))
    with EmuEngine(engine_params(vh, sc, case, num_buckets=1 << 13, pool_blocks=1 << 13, tri_arena_bytes=16 << 20), integrate_rev=revs[0], alloc_rev=revs[1],
                   mc_rev=revs[2]) as e:
        for i in range(3):
            d, rgb, c2w = sc.frame(i)
            o.process_frame(d, rgb, c2w)
            e.process_frame(d, rgb, c2w)
            assert key_set(e.visible_keys()) == key_set(o.visible_keys()), f"visible set differs in frame {i}"
            assert e.last_updates == o.last_updates and e.last_triangles == o.last_triangles, f"frame {i}"
        keys = o.all_keys()
        so, wo, co, _ = o.get_blocks(keys)
        se, we, ce, found, neg = e.get_blocks(keys)
        assert found.all() and np.array_equal(se.view(np.uint32), so.view(np.uint32)) and np.array_equal(we, wo) and np.array_equal(ce, co)
        xyz_o, rgb_o = o.triangles()
        xyz_e, rgb_e = e.block_triangles(mesh_order(keys))
        assert xyz_e.shape == xyz_o.shape and np.array_equal(xyz_e.view(np.uint32), xyz_o.view(np.uint32)) and np.array_equal(rgb_e, rgb_o)
        assert len(keys) > 100


@pytest.mark.parametrize("mc_rev", [0, 1])
def test_emulated_full_map_extraction(vh, ob, synth, mc_rev):
    """VH_MESH_FULL_MAP: every allocated block re-meshed against the whole map (a corner counts if its block is allocated at
    all) — list_all_blocks_kernel + both marching-cubes kernels with full_map = 1 — against the oracle's full-map pass;
    the per-frame meshes are untouched by it and it yields at least as many triangles"""
    sc = synth.Scene(**SMALL)
    o = ob.Oracle(oracle_params(ob, sc, CASE))
    with EmuEngine(engine_params(vh, sc, CASE, num_buckets=1 << 12, pool_blocks=1 << 12, tri_arena_bytes=16 << 20), mc_rev=mc_rev) as e:
        for i in (0, 6, 12):                      # views that overlap only partly: blocks whose neighbours were not in their frame's working set
            d, rgb, c2w = sc.frame(i)
            o.process_frame(d, rgb, c2w)
            e.process_frame(d, rgb, c2w)
        keys = mesh_order(o.all_keys())
        ref_xyz, ref_rgb = o.triangles()
        n_full = e.full_map_mc()
        per_frame_xyz, per_frame_rgb = e.block_triangles(keys)
        assert np.array_equal(per_frame_xyz, ref_xyz) and np.array_equal(per_frame_rgb, ref_rgb), "full-map extraction disturbed the per-frame meshes"
        assert o.full_map_mc() == n_full
        full_xyz_o, full_rgb_o = o.triangles()
        full_xyz, full_rgb = e.block_triangles(keys, full_map=True)
        assert full_xyz.shape == full_xyz_o.shape and np.array_equal(full_xyz.view(np.uint32), full_xyz_o.view(np.uint32)) and np.array_equal(full_rgb, full_rgb_o)
        assert n_full > len(ref_xyz) > 0


def test_emulated_engine_under_another_thread_order():
    """the emulator runs a CTA's threads round-robin in index order, which could hide a missing barrier (the writer happens to
    run first); VH_EMU_ORDER re-runs with every round in reverse / random order — results must not depend on it"""
    import os
    import subprocess
    import sys
    from emu.binding import lib
    lib()                                     # built once, here; the child processes load it as it is
    here = os.path.dirname(os.path.abspath(__file__))
    for order in ("reverse", "random:3"):
        env = dict(os.environ, VH_EMU_ORDER=order, VH_EMU_NO_REBUILD="1")
        r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_emu_engine.py"), "-x", "-q", "-p", "no:cacheprovider",
                            "-k", "matches_oracle and 2-0-1 or sharded and 3-1-2 or merge_equals"], env=env, capture_output=True, text=True, cwd=os.path.dirname(here))
        assert r.returncode == 0, f"VH_EMU_ORDER={order}:\n{r.stdout[-2000:]}"
        assert "3 passed" in r.stdout, r.stdout[-500:]
