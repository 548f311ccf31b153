"""One rank of the sharded-map GPU test (one process per GPU). Launched by tests/test_gpu_multi.py through
torch.distributed.run; torch.distributed (gloo) only carries the NCCL id and the comparison data between the ranks —
the frames travel by the engine's own ncclBroadcast."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import key_set, sort_keys  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    vh = importlib.import_module("voxel-hashing-sdf_b200")
    synth = importlib.import_module("voxel-hashing-sdf_b200.synth")
    torch.cuda.set_device(rank)
    sc = synth.Scene(width=320, height=240, room=(5.0, 4.0, 2.6), n_frames=40, spheres=((3.9, 2.0, 1.0, 0.5), (1.0, 3.0, 1.6, 0.4)), color=True, holes=0.02)
    kw = dict(vox_size=0.02, trunc_margin=0.1, max_depth=3.5, use_color=1, num_buckets=1 << 16, pool_blocks=1 << 16, tri_arena_bytes=64 << 20)
    n_frames = int(os.environ.get("VH_MULTI_FRAMES", "8"))
    eng = vh.TsdfEngine(vh.params_for_scene(sc, device=rank, shard_rank=rank, shard_count=world, **kw))
    ids = [vh.TsdfEngine.shard_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    eng.shard_connect(ids[0])

    single = vh.TsdfEngine(vh.params_for_scene(sc, device=0, **kw)) if rank == 0 else None
    for i in range(n_frames):
        d, rgb, c2w = sc.frame(i)
        if rank == 0:
            eng.integrate_sharded(d, rgb, c2w)
            single.processFrame(d, rgb, c2w)
        else:
            eng.integrate_sharded(None, None, c2w if i % 2 == 0 else None)      # both pose paths: given, and taken from the broadcast
        eng.sync()
        g = eng.shard_stats()
        # every rank's visible blocks are its own
        vis = eng.visible_keys()
        assert all(vh.owner_of_block(*k, world) == rank for k in vis)
        if rank == 0:
            s1 = single.stats()
            assert (g.visible_blocks, g.voxel_updates, g.triangles) == (s1.visible_blocks, s1.voxel_updates, s1.triangles), \
                f"frame {i}: group {(g.visible_blocks, g.voxel_updates, g.triangles)} vs single {(s1.visible_blocks, s1.voxel_updates, s1.triangles)}"
    # voxels: every shard's blocks equal the single-GPU map's
    mine = sort_keys(eng.allocated_keys())
    sdf, w, rgbv, found = eng.download_blocks(mine)
    assert found.all()
    payload = [None] * world
    dist.all_gather_object(payload, (mine, sdf, w, rgbv))
    xyz, trgb = eng.shard_triangles()
    xyz_full, _ = eng.shard_triangles(vh.VH_MESH_FULL_MAP)
    if rank == 0:
        all_keys = np.concatenate([p[0] for p in payload])
        assert key_set(all_keys) == key_set(single.allocated_keys()) and len(all_keys) == len(key_set(all_keys))
        for keys_r, sdf_r, w_r, rgb_r in payload:
            s1, w1, c1, f1 = single.download_blocks(keys_r)
            assert f1.all() and np.array_equal(s1, sdf_r) and np.array_equal(w1, w_r) and np.array_equal(c1, rgb_r)
        x1, c1 = single.triangles()
        assert xyz.shape == x1.shape and np.array_equal(xyz, x1) and np.array_equal(trgb, c1), "gathered mesh != single-GPU mesh"
        x1f, _ = single.triangles(vh.VH_MESH_FULL_MAP)
        assert xyz_full.shape == x1f.shape and np.array_equal(xyz_full, x1f), "gathered full-map mesh != single-GPU full-map mesh"
        print(f"MULTI_GPU_OK world={world} frames={n_frames} blocks={len(all_keys)} per_rank={[len(p[0]) for p in payload]} triangles={len(xyz)} full_map={len(xyz_full)}",
              flush=True)
    else:
        assert len(xyz) == 0
    dist.barrier()
    eng.close()
    if single:
        single.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
