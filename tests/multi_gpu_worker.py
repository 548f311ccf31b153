"""One rank of the sharded-map GPU test (one process per GPU). Launched by tests/test_gpu_multi.py through
torch.distributed.run; torch.distributed (gloo) only carries the NCCL id and the comparison data between the ranks —
the frames travel by the engine's own means (peer-memory reads of rank 0's frame ring, or ncclBroadcast)."""
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


MODES = [   # label, environment, frames in flight (no sync between frames, device-resident inputs on rank 0)
    ("pull+split-rays", {"VH_ALLOC_REV": "2"}, False),
    ("pull+split-rays, frames in flight", {"VH_ALLOC_REV": "2"}, True),
    ("pull+replicated-rays", {"VH_ALLOC_REV": "0"}, False),
    ("nccl-broadcast+split-rays", {"VH_ALLOC_REV": "2", "VH_SHARD_BCAST": "nccl"}, True),
    ("pull+split-rays+direct-integrate", {"VH_ALLOC_REV": "2", "VH_INTEGRATE_REV": "1"}, True),
    ("pull+split-rays+mc-colour-tile", {"VH_ALLOC_REV": "2", "VH_MC_COLOR_TILE": "1"}, True),
]


def run_mode(vh, sc, kw, rank, world, n_frames, label, env, in_flight, ref):
    saved = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        eng = vh.TsdfEngine(vh.params_for_scene(sc, device=rank, shard_rank=rank, shard_count=world, **kw))
        ids = [vh.TsdfEngine.shard_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        eng.shard_connect(ids[0])
    finally:
        for k, v in saved.items():
            os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
    frames = [sc.frame(i) for i in range(n_frames)]
    if in_flight:
        dev = [(torch.from_numpy(d).cuda(), torch.from_numpy(c).cuda()) for d, c, _ in frames] if rank == 0 else None
        torch.cuda.synchronize()
        for rep in range(2):          # twice: vh_reset of a sharded map in between (collective)
            for i, (d, rgb, c2w) in enumerate(frames):
                if rank == 0:
                    eng.integrate_sharded_device(dev[i][0].data_ptr(), dev[i][1].data_ptr(), c2w)
                else:
                    eng.integrate_sharded_device(None, None, c2w)
            eng.sync()
            if rep == 0:
                eng.reset()
    else:
        for i, (d, rgb, c2w) in enumerate(frames):
            if rank == 0:
                eng.integrate_sharded(d, rgb, c2w)
            else:
                eng.integrate_sharded(None, None, c2w if i % 2 == 0 else None)      # both pose paths: given, and taken from the frame
            eng.sync()
            g = eng.shard_stats()
            vis = eng.visible_keys()
            assert all(vh.owner_of_block(*k, world) == rank for k in vis), f"{label}: a rank lists blocks it does not own"
            if rank == 0:
                s1 = ref["stats"][i]
                assert (g.visible_blocks, g.voxel_updates, g.triangles) == s1, f"{label}: frame {i}: group {(g.visible_blocks, g.voxel_updates, g.triangles)} vs single {s1}"
    mine = sort_keys(eng.allocated_keys())
    sdf, w, rgbv, found = eng.download_blocks(mine)
    assert found.all()
    payload = [None] * world
    dist.all_gather_object(payload, (mine, sdf, w, rgbv))
    xyz, trgb = eng.shard_triangles()
    xyz_full, _ = eng.shard_triangles(vh.VH_MESH_FULL_MAP)
    msg = None
    if rank == 0:
        single = ref["engine"]
        all_keys = np.concatenate([p[0] for p in payload])
        assert key_set(all_keys) == key_set(single.allocated_keys()) and len(all_keys) == len(key_set(all_keys)), f"{label}: block sets differ"
        for keys_r, sdf_r, w_r, rgb_r in payload:
            s1, w1, c1, f1 = single.download_blocks(keys_r)
            assert f1.all() and np.array_equal(s1, sdf_r) and np.array_equal(w1, w_r) and np.array_equal(c1, rgb_r), f"{label}: voxels differ"
        x1, c1 = ref["mesh"]
        assert xyz.shape == x1.shape and np.array_equal(xyz, x1) and np.array_equal(trgb, c1), f"{label}: gathered mesh != single-GPU mesh"
        x1f = ref["full"]
        assert xyz_full.shape == x1f.shape and np.array_equal(xyz_full, x1f), f"{label}: gathered full-map mesh != single-GPU full-map mesh"
        msg = f"  ok [{label}] blocks={len(all_keys)} per_rank={[len(p[0]) for p in payload]} triangles={len(xyz)} full_map={len(xyz_full)}"
    else:
        assert len(xyz) == 0
    dist.barrier()
    eng.close()
    return msg


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    vh = importlib.import_module("voxel-hashing-sdf_b200")
    synth = importlib.import_module("voxel-hashing-sdf_b200.synth")
    torch.cuda.set_device(rank)
    sc = synth.Scene(width=320, height=240, room=(5.0, 4.0, 2.6), n_frames=40, spheres=((3.9, 2.0, 1.0, 0.5), (1.0, 3.0, 1.6, 0.4)), color=True, holes=0.02)
    kw = dict(vox_size=0.02, trunc_margin=0.1, max_depth=3.5, use_color=1, num_buckets=1 << 16, pool_blocks=1 << 16, tri_arena_bytes=64 << 20)
    n_frames = int(os.environ.get("VH_MULTI_FRAMES", "8"))
    ref = None
    if rank == 0:      # the single-GPU engine on the same frames: per-frame counters, final map, meshes
        single = vh.TsdfEngine(vh.params_for_scene(sc, device=0, **kw))
        stats = []
        for i in range(n_frames):
            single.processFrame(*sc.frame(i))
            s1 = single.stats()
            stats.append((s1.visible_blocks, s1.voxel_updates, s1.triangles))
        ref = dict(engine=single, stats=stats, mesh=single.triangles(), full=single.triangles(vh.VH_MESH_FULL_MAP)[0])
    lines = []
    only = os.environ.get("VH_MULTI_ONLY")
    for label, env, in_flight in MODES:
        if only and only not in label:
            continue
        lines.append(run_mode(vh, sc, kw, rank, world, n_frames, label, env, in_flight, ref))
    if rank == 0:
        print("\n".join(lines), flush=True)
        print(f"MULTI_GPU_OK world={world} frames={n_frames} modes={len(lines)}", flush=True)
        ref["engine"].close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
