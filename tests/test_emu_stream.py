"""CPU: the out-of-core tier's kernels (csrc/vh_stream.cu: far-block selection, eviction, upload, table rebuild) executed
under the SIMT stand-in of tests/emu together with the per-frame path, against the oracle."""
import numpy as np
import pytest

from emu.binding import EmuEngine, mesh_order
from util import engine_params, key_set, oracle_params

SMALL = dict(width=160, height=120, room=(4.0, 3.0, 2.5), n_frames=60, spheres=((2.8, 1.5, 1.0, 0.4),), color=True)
CASE = dict(scene=dict(color=True), vpb=8, vox_size=0.04, trunc=0.2, max_depth=3.0)
ENG = dict(num_buckets=1 << 10, pool_blocks=1 << 11, tri_arena_bytes=16 << 20)


def check_voxels(e, o, keys):
    so, wo, co, _ = o.get_blocks(keys)
    se, we, ce, found, neg = e.get_blocks(keys)
    assert found.all()
    assert np.array_equal(se.view(np.uint32), so.view(np.uint32)) and np.array_equal(we, wo) and np.array_equal(ce, co)
    assert np.array_equal(neg, (se < 0).sum(1))


def test_evict_and_upload_round_trip_is_invisible_to_later_frames(vh, ob, synth):
    """half of the map is evicted and uploaded again between frames (other pool slots, other table entries): every later
    frame and the final full-map mesh equal the oracle's uninterrupted run, bit for bit"""
    sc = synth.Scene(**SMALL)
    o = ob.Oracle(oracle_params(ob, sc, CASE))
    with EmuEngine(engine_params(vh, sc, CASE, **ENG)) as e:
        for i in range(5):
            d, rgb, c2w = sc.frame(i)
            o.process_frame(d, rgb, c2w)
            e.process_frame(d, rgb, c2w)
            assert key_set(e.visible_keys()) == key_set(o.visible_keys()) and e.last_updates == o.last_updates and e.last_triangles == o.last_triangles
            if i in (1, 3):
                keys = e.all_keys()
                n0, free0 = e.num_blocks, e.free_slots
                out = keys[(i % 2)::2]                                         # every other block, in insertion order
                listed = np.concatenate([out, out[:7], np.array([[900, 900, 900]], np.int32)])      # repeats and a key that does not exist
                sdf, w, c, found, released = e.evict_blocks(listed)
                assert released == len(out) and found[:len(out)].all() and not found[-1]
                assert e.num_blocks == n0 - len(out) and e.free_slots == free0 + len(out) and e.num_visible == 0
                assert np.array_equal(e.all_keys(), keys[np.arange(len(keys)) % 2 != i % 2]), "key_heap lost its insertion order"
                assert not e.get_blocks(out)[3].any(), "evicted blocks are still found"
                check_voxels(e, o, e.all_keys())                              # the rest is untouched
                order = np.random.RandomState(i).permutation(len(out))       # come back in another order
                e.upload_blocks(out[order], sdf[:len(out)][order], w[:len(out)][order], c[:len(out)][order])
                assert e.num_blocks == n0 and e.free_slots == free0
                check_voxels(e, o, o.all_keys())
        keys = o.all_keys()
        assert key_set(e.all_keys()) == key_set(keys)
        check_voxels(e, o, keys)
        assert e.full_map_mc() == o.full_map_mc()
        xyz_o, rgb_o = o.triangles()
        xyz_e, rgb_e = e.block_triangles(mesh_order(keys), full_map=True)
        assert np.array_equal(xyz_e.view(np.uint32), xyz_o.view(np.uint32)) and np.array_equal(rgb_e, rgb_o)


def test_far_blocks_follow_the_reference_residency_rule(vh, ob, synth):
    """vh_far_blocks = allocated blocks whose chunk is outside the chunk cube / sphere around the frustum centre
    (tsdf.cu:166-187,300-312), i.e. the blocks the reference would not stream in for that pose"""
    sc = synth.Scene(**dict(SMALL, room=(9.0, 7.0, 2.5), spheres=()))
    case = dict(CASE, max_depth=4.0)
    o = ob.Oracle(oracle_params(ob, sc, case))
    with EmuEngine(engine_params(vh, sc, case, num_buckets=1 << 12, pool_blocks=1 << 12, tri_arena_bytes=16 << 20, mc_per_frame=0)) as e:
        for i in (0, 10, 20, 30, 40, 50):                                       # a full turn: the map extends all around the camera circle
            d, rgb, c2w = sc.frame(i)
            o.begin_frame(c2w); o.stage_allocate(d); o.stage_integrate(d, rgb)
            e.process_frame(d, rgb, c2w)
        keys = e.all_keys()
        _, _, c2w = sc.frame(30)
        far = e.far_blocks(c2w)
        c2w32 = np.ascontiguousarray(c2w, np.float32)
        want = [tuple(k) for k in keys.tolist() if not o.L.vo_block_is_candidate(o.h, c2w32.ctypes.data, int(k[0]), int(k[1]), int(k[2]))]
        assert key_set(far) == set(want) and 0 < len(far) < len(keys)
        # evicting exactly those leaves a map that still integrates the frame of that pose like the oracle
        e.evict_blocks(far)
        d, rgb, c2w = sc.frame(30)
        o.begin_frame(c2w); o.stage_allocate(d); upd = o.stage_integrate(d, rgb)
        e.process_frame(d, rgb, c2w)
        assert key_set(e.visible_keys()) == key_set(o.visible_keys()) and e.last_updates == upd
        check_voxels(e, o, o.visible_keys())


def test_tombstones_and_table_rebuild(vh, ob, synth):
    """released entries stay as tombstones (probe sequences run through them); the rebuild re-inserts the live entries into
    a cleared table and the map goes on working"""
    sc = synth.Scene(**SMALL)
    o = ob.Oracle(oracle_params(ob, sc, CASE))
    with EmuEngine(engine_params(vh, sc, CASE, num_buckets=128, entries_per_bucket=4, pool_blocks=600, tri_arena_bytes=16 << 20)) as e:   # table of 1024
        for i in range(2):
            d, rgb, c2w = sc.frame(i)
            o.process_frame(d, rgb, c2w); e.process_frame(d, rgb, c2w)
        keys = e.all_keys()
        sdf, w, c, found, released = e.evict_blocks(keys[::3])
        assert released == len(keys[::3])
        check_voxels(e, o, e.all_keys())                                      # look-ups through tombstones
        live = e.rebuild_table()
        assert live == e.num_blocks
        check_voxels(e, o, e.all_keys())
        e.upload_blocks(keys[::3], sdf, w, c)
        for i in range(2, 4):
            d, rgb, c2w = sc.frame(i)
            o.process_frame(d, rgb, c2w); e.process_frame(d, rgb, c2w)
            assert key_set(e.visible_keys()) == key_set(o.visible_keys()) and e.last_updates == o.last_updates and e.last_triangles == o.last_triangles
        check_voxels(e, o, o.all_keys())


def test_pool_exhaustion_keeps_the_free_list_and_key_heap_consistent(vh, synth):
    """inserting past the pool: MAP_POOL_FULL is raised, the lanes without a slot publish the 'pool full' sentinel (not the
    'not yet published' one a reader would spin on), free_top ends at 0 (never negative), key_heap holds exactly the blocks that
    got storage; evicting then pushes inside the stack and the freed slots are handed out again (vh_map.cuh:111,123)"""
    sc = synth.Scene(**SMALL)
    pool = 40
    with EmuEngine(engine_params(vh, sc, CASE, num_buckets=1 << 8, pool_blocks=pool, tri_arena_bytes=1 << 20)) as e:
        keys = np.stack([np.arange(100), np.arange(100) % 7, -np.arange(100)], 1).astype(np.int32)
        z = np.zeros((len(keys), 512), np.float32)
        rc = e.upload_blocks(keys, z - 1.0, z + 3.0, np.zeros((len(keys), 512, 3), np.uint8), check=False)
        assert rc & 2, "MAP_POOL_FULL not raised"
        assert e.free_slots == 0 and e.num_blocks == pool
        have = e.get_blocks(keys)[3]
        assert have.sum() == pool
        assert key_set(e.all_keys()) == key_set(keys[have]), "key_heap does not list exactly the blocks that got a slot"
        stored = keys[have]
        _, _, _, found, released = e.evict_blocks(stored[:10])
        assert released == 10 and found.all() and e.free_slots == 10 and e.num_blocks == pool - 10
        fresh = np.array([[500 + i, 1, 2] for i in range(5)], np.int32)
        e.upload_blocks(fresh, z[:5] + 0.5, z[:5] + 1.0, np.zeros((5, 512, 3), np.uint8), check=False)
        assert e.free_slots == 5 and e.num_blocks == pool - 5
        s, w, _, found, neg = e.get_blocks(fresh)
        assert found.all() and (s == 0.5).all() and (w == 1.0).all() and (neg == 0).all()
        s, w, _, found, neg = e.get_blocks(stored[10:])
        assert found.all() and (s == -1.0).all() and (w == 3.0).all() and (neg == 512).all(), "an older block lost its slot"


def test_host_residency_rule_matches_oracle(vh, ob, synth):
    """vh_blocks_resident (host arithmetic through the C ABI, no GPU) against the oracle's chunk-candidate test, for blocks all
    over the place and several poses, including chunk-sphere boundary cases at three voxel sizes"""
    rng = np.random.RandomState(3)
    for vox, room in ((0.04, (9.0, 7.0, 2.5)), (0.005, (8.0, 6.0, 3.0)), (0.0225, (8.0, 6.0, 3.0))):
        sc = synth.Scene(width=160, height=120, room=room, n_frames=12)
        case = dict(scene={}, vpb=8, vox_size=vox, trunc=5 * vox, max_depth=4.0)
        o = ob.Oracle(oracle_params(ob, sc, case))
        p = engine_params(vh, sc, case)
        span = int(6.0 / (8 * vox))
        keys = rng.randint(-span // 4, span, size=(4000, 3)).astype(np.int32)
        for i in (0, 3, 7):
            c2w = np.ascontiguousarray(sc.pose(i), np.float32)
            want = np.array([bool(o.L.vo_block_is_candidate(o.h, c2w.ctypes.data, int(k[0]), int(k[1]), int(k[2]))) for k in keys])
            got = vh.blocks_resident(p, c2w, keys)
            assert np.array_equal(got, want)
            assert 0 < want.sum() < len(keys)
