"""GPU (B200): the out-of-core tier through the C ABI (vh_far_blocks / vh_evict_blocks / vh_upload_blocks, csrc/vh_stream.cu):
a run that streams blocks out and back in around every frame reproduces the oracle. First green on B200 in profiles/r02a."""
import os

import numpy as np
import pytest

from test_gpu_parity import assert_triangles_match, assert_voxels_match
from util import engine_params, key_set, oracle_params

pytestmark = pytest.mark.gpu

SMALL = dict(width=320, height=240, room=(9.0, 7.0, 2.6), n_frames=60, spheres=((6.9, 3.5, 1.0, 0.5),), color=True)
CASE = dict(scene=dict(color=True), vpb=8, vox_size=0.02, trunc=0.1, max_depth=4.0)


def test_evict_upload_round_trip_and_residency(vh, ob, synth):
    sc = synth.Scene(**SMALL)
    o = ob.Oracle(oracle_params(ob, sc, CASE))
    p = engine_params(vh, sc, CASE, num_buckets=1 << 16, pool_blocks=1 << 16, tri_arena_bytes=512 << 20)
    store = {}
    with vh.TsdfEngine(p) as e:
        for step, i in enumerate((0, 10, 20, 30, 40, 50, 0, 10)):             # a full turn and on: blocks leave and come back
            d, rgb, c2w = sc.frame(i)
            # stream in what the reference would upload for this pose, stream out what it would not keep
            if store:
                keys = np.array(list(store.keys()), np.int32)
                back = keys[vh.blocks_resident(p, c2w, keys)]
                if len(back):
                    e.upload_blocks(back, np.stack([store[tuple(k)][0] for k in back.tolist()]), np.stack([store[tuple(k)][1] for k in back.tolist()]),
                                    np.stack([store[tuple(k)][2] for k in back.tolist()]))
                    for k in back.tolist():
                        del store[tuple(k)]
            far = e.far_blocks(c2w)
            if len(far):
                n0 = e.stats().allocated_blocks
                s, w, c, found = e.evict_blocks(far)
                assert found.all() and e.stats().allocated_blocks == n0 - len(far)
                for k, a, b, cc in zip(far.tolist(), s, w, c):
                    store[tuple(k)] = (a.copy(), b.copy(), cc.copy())
            o.process_frame(d, rgb, c2w)
            e.processFrame(d, rgb, c2w)
            assert key_set(e.visible_keys()) == key_set(o.visible_keys()), f"visible set differs at step {step}"
            st = e.stats()
            assert st.voxel_updates == o.last_updates and st.triangles == o.last_triangles, f"step {step}"
        assert store, "nothing was ever streamed out: the case does not exercise the tier"
        # device map + store = the oracle's map, bit for bit
        keys = o.all_keys()
        on_dev = key_set(e.allocated_keys())
        assert on_dev | set(store.keys()) == key_set(keys) and not (on_dev & set(store.keys()))
        dev_keys = np.array([k for k in keys.tolist() if tuple(k) in on_dev], np.int32)
        sdf, w, rgb_, _ = o.get_blocks(dev_keys)
        assert_voxels_match(e, dev_keys, sdf, w, rgb_, True)
        st_keys = np.array([k for k in keys.tolist() if tuple(k) in store], np.int32)
        sdf, w, rgb_, _ = o.get_blocks(st_keys)
        assert np.array_equal(np.stack([store[tuple(k)][0] for k in st_keys.tolist()]), sdf)
        assert np.array_equal(np.stack([store[tuple(k)][1] for k in st_keys.tolist()]), w)
        assert np.array_equal(np.stack([store[tuple(k)][2] for k in st_keys.tolist()]), rgb_)
        # everything back on the device: the full-map mesh is the oracle's
        e.upload_blocks(st_keys, np.stack([store[tuple(k)][0] for k in st_keys.tolist()]), np.stack([store[tuple(k)][1] for k in st_keys.tolist()]),
                        np.stack([store[tuple(k)][2] for k in st_keys.tolist()]))
        full = e.triangles(vh.VH_MESH_FULL_MAP)
        assert o.full_map_mc() == len(full[0])
        assert_triangles_match(*full, *o.triangles(), True)
