"""Async sharded pipeline timing (run under torch.distributed.run): N frames back to back, one sync at the end."""
import importlib, os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
dist.init_process_group("gloo"); rank, world = dist.get_rank(), dist.get_world_size()
vh = importlib.import_module("voxel-hashing-sdf_b200"); synth = importlib.import_module("voxel-hashing-sdf_b200.synth")
torch.cuda.set_device(rank)
sc = synth.make_scene("C2", color=True); N = int(os.environ.get("N", "300")); ARENA = int(os.environ.get("ARENA_MB", "1280"))
hd = torch.empty((N, 480, 640), dtype=torch.float32).pin_memory(); hc = torch.empty((N, 480, 640, 3), dtype=torch.uint8).pin_memory(); poses = []
for i in range(N):
    d, rgb, c2w = sc.frame(i); hd[i] = torch.from_numpy(d); hc[i] = torch.from_numpy(rgb); poses.append(c2w)
p = vh.params_for_scene(sc, vox_size=0.005, trunc_margin=0.025, max_depth=10.0, num_buckets=1 << 20, pool_blocks=1 << 20, use_color=1,
                        mc_per_frame=1, device=rank, shard_rank=rank, shard_count=world, tri_arena_bytes=ARENA << 20)
eng = vh.TsdfEngine(p)
ids = [vh.TsdfEngine.shard_unique_id() if rank == 0 else None]; dist.broadcast_object_list(ids, src=0); eng.shard_connect(ids[0])
for rep in range(2):
    eng.reset(); dist.barrier(); t0 = time.perf_counter(); tq = 0.0
    for i in range(N):
        t1 = time.perf_counter()
        if rank == 0: eng.integrate_sharded(hd[i].data_ptr(), hc[i].data_ptr(), poses[i])
        else: eng.integrate_sharded(None, None, poses[i])
        tq += time.perf_counter() - t1
    eng.sync(); dt = time.perf_counter() - t0; s = eng.stats()
    print(f"rank {rank} rep {rep}: {1e3*dt/N:.3f} ms/frame, enqueue {1e3*tq/N:.3f} ms/frame, compactions {s.arena_compactions}, forced syncs {s.forced_syncs}", flush=True)
    dist.barrier()
eng.close(); dist.destroy_process_group()
