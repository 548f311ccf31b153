import importlib, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
vh = importlib.import_module("voxel-hashing-sdf_b200"); synth = importlib.import_module("voxel-hashing-sdf_b200.synth")
sc = synth.make_scene("C2", color=True)
p = vh.params_for_scene(sc, vox_size=0.005, trunc_margin=0.025, max_depth=10.0, num_buckets=1<<20, pool_blocks=3<<20, use_color=1, tri_arena_bytes=int(sys.argv[1]) << 20)
a = vh.TsdfEngine(p); b = vh.TsdfEngine(p)
N = int(sys.argv[2])
frames = [sc.frame(i) for i in range(N)]
dd = [torch.from_numpy(f[0]).cuda() for f in frames]; rr = [torch.from_numpy(f[1]).cuda() for f in frames]
ta, tb = [], []
for i in range(N):
    a.processFrame(*frames[i]); ta.append(a.stats().triangles)
for i in range(N):
    b.integrate_device(dd[i].data_ptr(), rr[i].data_ptr(), frames[i][2]); s = b.stats(); tb.append(s.triangles)
print("sync ", ta[:5], ta[-5:], sum(ta))
print("async", tb[:5], tb[-5:], sum(tb), "compactions", a.stats().arena_compactions, b.stats().arena_compactions)
print("mismatch frames", [i for i in range(N) if ta[i] != tb[i]][:20])
b.sync()
xa, _ = a.triangles(); xb, _ = b.triangles()
print("final mesh equal", xa.shape, xb.shape, np.array_equal(xa, xb))
