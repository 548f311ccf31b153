"""Stage times of the sharded pipeline (run under torch.distributed.run, one rank per GPU): prints rank 0's averages."""
import importlib, os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
dist.init_process_group("gloo"); rank, world = dist.get_rank(), dist.get_world_size()
vh = importlib.import_module("voxel-hashing-sdf_b200"); synth = importlib.import_module("voxel-hashing-sdf_b200.synth")
torch.cuda.set_device(rank)
sc = synth.make_scene("C2", color=True); N = int(os.environ.get("N", "120"))
frames = [sc.frame(i) for i in range(N)]
for mc in (0, 1):
    p = vh.params_for_scene(sc, vox_size=0.005, trunc_margin=0.025, max_depth=10.0, num_buckets=1 << 20, pool_blocks=1 << 20, use_color=1,
                            mc_per_frame=mc, device=rank, shard_rank=rank, shard_count=world, tri_arena_bytes=2 << 30)
    eng = vh.TsdfEngine(p)
    ids = [vh.TsdfEngine.shard_unique_id() if rank == 0 else None]; dist.broadcast_object_list(ids, src=0); eng.shard_connect(ids[0])
    acc = np.zeros(4); t0 = time.perf_counter()
    for i, (d, rgb, c2w) in enumerate(frames):
        eng.integrate_sharded(d if rank == 0 else None, rgb if rank == 0 else None, c2w)
        s = eng.stats()
        if i >= 20: acc += [s.ms_upload, s.ms_alloc, s.ms_integrate, s.ms_mc]
    dt = time.perf_counter() - t0
    if rank == 0:
        print(f"world={world} mc={mc} avg ms: bcast+h2d {acc[0]/(N-20):.4f} pack+alloc {acc[1]/(N-20):.4f} integrate {acc[2]/(N-20):.4f} mc+barriers {acc[3]/(N-20):.4f}  (sync loop {1e3*dt/N:.3f} ms/frame)", flush=True)
    dist.barrier(); eng.close()
dist.destroy_process_group()
