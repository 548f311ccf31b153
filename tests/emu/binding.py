"""ctypes binding of tests/emu/libvh_emu.so: the engine's kernel sources executed on the CPU (test infrastructure).

Only tests import this. The product (voxel-hashing-sdf_b200/, include/) never does: the engine has no CPU path.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = os.environ.get("VH_EMU_LIB", "libvh_emu.so")      # libvh_emu_asan.so: the AddressSanitizer build (tests/emu/Makefile)
LIB = os.path.join(HERE, LIB_NAME)


class IntegrateIO(C.Structure):
    _fields_ = [("W", C.c_int), ("H", C.c_int),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("max_depth", C.c_float),
                ("vox_size", C.c_float), ("trunc", C.c_float),
                ("use_color", C.c_int), ("cull", C.c_int), ("two_steps", C.c_int), ("verify", C.c_int), ("exact_color", C.c_int),
                ("ctas", C.c_int), ("variant", C.c_int),
                ("weight_bound", C.c_uint), ("frame", C.c_uint), ("rcp_seed", C.c_uint),
                ("c2w", C.c_void_p), ("depth", C.c_void_p), ("rgb", C.c_void_p),
                ("n_visible", C.c_int), ("keys_xyz", C.c_void_p), ("slots", C.c_void_p),
                ("sdf", C.c_void_p), ("wgt", C.c_void_p), ("rgb4", C.c_void_p), ("neg_count", C.c_void_p),
                ("voxel_updates", C.c_ulonglong), ("culled", C.c_ulonglong), ("mismatch", C.c_ulonglong), ("collectives", C.c_ulonglong), ("slow_steps", C.c_ulonglong),
                ("engine_error", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if os.environ.get("VH_EMU_NO_REBUILD") != "1" or not os.path.exists(LIB):      # (a parent test process has just built it)
            subprocess.check_call(["make", "-s", "-B", "-C", HERE, LIB_NAME])   # always rebuilt: it mirrors the kernel sources of the moment
        _lib = C.CDLL(LIB)
        _lib.emu_integrate.argtypes = [C.POINTER(IntegrateIO)]
        _lib.emu_integrate.restype = C.c_int
    return _lib


class EmuMap:
    """Voxel planes + key -> slot dictionary driven by the emulated integrate kernel (the visible list comes from outside)."""

    def __init__(self, scene, vox_size, trunc, max_depth, pool_blocks, color):
        self.sc, self.vox, self.trunc, self.maxd, self.color = scene, vox_size, trunc, max_depth, color
        self.sdf = np.zeros(pool_blocks * 512, np.float32)
        self.wgt = np.zeros(pool_blocks * 512, np.float32)
        self.rgb4 = np.zeros(pool_blocks * 512 * 4, np.uint8)
        self.neg = np.zeros(pool_blocks, np.int32)
        self.slot_of = {}
        self.frames = 0

    def slots_for(self, keys):
        out = np.empty(len(keys), np.int32)
        for i, k in enumerate(map(tuple, np.asarray(keys).tolist())):
            if k not in self.slot_of:
                self.slot_of[k] = len(self.slot_of)
            out[i] = self.slot_of[k]
        assert len(self.slot_of) <= len(self.neg), "emulated pool too small"
        return out

    def integrate(self, keys, depth, rgb, c2w, *, cull=1, two_steps=0, verify=0, exact_color=0, ctas=2, variant=0, rcp_seed=1):
        self.frames += 1
        keys = np.ascontiguousarray(keys, np.int32).reshape(-1, 3)
        slots = self.slots_for(keys)
        depth = np.ascontiguousarray(depth, np.float32)
        rgb = np.ascontiguousarray(rgb, np.uint8)
        c2w = np.ascontiguousarray(c2w, np.float32)
        sc = self.sc
        io = IntegrateIO(W=sc.width, H=sc.height, fx=sc.fx, fy=sc.fy, cx=sc.cx, cy=sc.cy, max_depth=self.maxd, vox_size=self.vox,
                         trunc=self.trunc, use_color=int(self.color), cull=cull, two_steps=two_steps, verify=verify,
                         exact_color=exact_color, ctas=ctas, variant=variant, weight_bound=self.frames, frame=self.frames, rcp_seed=rcp_seed,
                         c2w=c2w.ctypes.data, depth=depth.ctypes.data, rgb=rgb.ctypes.data, n_visible=len(keys),
                         keys_xyz=keys.ctypes.data, slots=slots.ctypes.data, sdf=self.sdf.ctypes.data, wgt=self.wgt.ctypes.data,
                         rgb4=self.rgb4.ctypes.data, neg_count=self.neg.ctypes.data)
        rc = lib().emu_integrate(C.byref(io))
        assert rc == 0, f"emu_integrate failed: {rc}"
        return io

    def blocks(self, keys):
        slots = self.slots_for(np.asarray(keys).reshape(-1, 3))
        s = self.sdf.reshape(-1, 512)[slots]
        w = self.wgt.reshape(-1, 512)[slots]
        c = self.rgb4.reshape(-1, 512, 4)[slots]
        return s, w, c, slots


# ---- the whole per-frame path under emulation (emu_engine.cpp) --------------------------------------------------------
class EmuEngine:
    """pack -> allocate -> integrate -> marching cubes with the engine's kernel sources on the CPU; mirrors the subset of
    voxel-hashing-sdf_b200.TsdfEngine the parity tests use."""

    def __init__(self, params, integrate_rev=0, cull=1, exact_color=0, alloc_rev=0, mc_rev=0):
        L = lib()
        L.emu_create.restype = C.c_void_p
        L.emu_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.emu_process_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.emu_phase_integrate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.emu_phase_mc.argtypes = [C.c_void_p]
        L.emu_phase_keys.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.emu_connect.argtypes = [C.c_void_p, C.c_int]
        L.emu_last_updates.restype = C.c_ulonglong
        L.emu_last_triangles.restype = C.c_ulonglong
        L.emu_block_triangles.restype = C.c_longlong
        for f in ("emu_destroy", "emu_num_visible", "emu_last_updates", "emu_last_triangles", "emu_num_blocks", "emu_free_slots"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.emu_visible_keys.argtypes = [C.c_void_p, C.c_void_p]
        L.emu_all_keys.argtypes = [C.c_void_p, C.c_void_p]
        L.emu_get_blocks.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 5
        L.emu_block_triangles.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.emu_full_map_mc.restype = C.c_longlong
        L.emu_full_map_mc.argtypes = [C.c_void_p]
        self.L, self.params = L, params
        self.h = L.emu_create(C.addressof(params), integrate_rev, cull, exact_color, alloc_rev, mc_rev)
        assert self.h, "emu_create rejected the parameters"

    def close(self):
        if self.h:
            self.L.emu_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def process_frame(self, depth, rgb, c2w, check=True):
        """returns the sticky error flags: MapError bits (1 table full, 2 pool full, 4 key range) | engine_error << 8"""
        depth = np.ascontiguousarray(depth, np.float32)
        c2w = np.ascontiguousarray(c2w, np.float32)
        rgb = None if rgb is None else np.ascontiguousarray(rgb, np.uint8)
        rc = self.L.emu_process_frame(self.h, depth.ctypes.data, None if rgb is None else rgb.ctypes.data, c2w.ctypes.data)
        assert rc == 0 or not check, f"map/engine error flags 0x{rc:x}"
        return rc

    def phase_integrate(self, depth, rgb, c2w):
        depth = np.ascontiguousarray(depth, np.float32)
        c2w = np.ascontiguousarray(c2w, np.float32)
        rgb = None if rgb is None else np.ascontiguousarray(rgb, np.uint8)
        rc = self.L.emu_phase_integrate(self.h, depth.ctypes.data, None if rgb is None else rgb.ctypes.data, c2w.ctypes.data)
        assert rc == 0, f"map/engine error flags 0x{rc:x}"

    def phase_keys(self, depth, c2w):
        depth = np.ascontiguousarray(depth, np.float32)
        c2w = np.ascontiguousarray(c2w, np.float32)
        rc = self.L.emu_phase_keys(self.h, depth.ctypes.data, c2w.ctypes.data)
        assert rc == 0, f"map error flags 0x{rc:x}"

    def phase_mc(self):
        rc = self.L.emu_phase_mc(self.h)
        assert rc == 0, f"map/engine error flags 0x{rc:x}"

    num_visible = property(lambda s: s.L.emu_num_visible(s.h))
    last_updates = property(lambda s: s.L.emu_last_updates(s.h))
    last_triangles = property(lambda s: s.L.emu_last_triangles(s.h))
    num_blocks = property(lambda s: s.L.emu_num_blocks(s.h))

    def visible_keys(self):
        out = np.zeros((max(self.num_visible, 1), 3), np.int32)
        self.L.emu_visible_keys(self.h, out.ctypes.data)
        return out[:self.num_visible]

    def all_keys(self):
        out = np.zeros((max(self.num_blocks, 1), 3), np.int32)
        self.L.emu_all_keys(self.h, out.ctypes.data)
        return out[:self.num_blocks]

    def get_blocks(self, keys):
        keys = np.ascontiguousarray(keys, np.int32).reshape(-1, 3)
        n = len(keys)
        sdf = np.zeros((n, 512), np.float32); w = np.zeros((n, 512), np.float32); rgb = np.zeros((n, 512, 3), np.uint8)
        found = np.zeros(n, np.uint8); neg = np.zeros(n, np.int32)
        self.L.emu_get_blocks(self.h, keys.ctypes.data, n, sdf.ctypes.data, w.ctypes.data, rgb.ctypes.data, found.ctypes.data, neg.ctypes.data)
        return sdf, w, rgb, found.astype(bool), neg

    # -- out-of-core tier (csrc/vh_stream.cu) -----------------------------------------------------
    def far_blocks(self, c2w):
        c2w = np.ascontiguousarray(c2w, np.float32)
        self.L.emu_far_blocks.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        n = self.L.emu_far_blocks(self.h, c2w.ctypes.data, None, 0)
        out = np.zeros((max(n, 1), 3), np.int32)
        self.L.emu_far_blocks(self.h, c2w.ctypes.data, out.ctypes.data, n)
        return out[:n]

    def evict_blocks(self, keys):
        keys = np.ascontiguousarray(keys, np.int32).reshape(-1, 3)
        n = len(keys)
        sdf = np.zeros((n, 512), np.float32); w = np.zeros((n, 512), np.float32); rgb = np.zeros((n, 512, 3), np.uint8)
        found = np.zeros(n, np.uint8)
        self.L.emu_evict_blocks.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 4
        released = self.L.emu_evict_blocks(self.h, keys.ctypes.data, n, sdf.ctypes.data, w.ctypes.data, rgb.ctypes.data, found.ctypes.data)
        assert released >= 0, "key list and table disagree"
        return sdf, w, rgb, found.astype(bool), released

    def upload_blocks(self, keys, sdf, w, rgb, check=True):
        keys = np.ascontiguousarray(keys, np.int32).reshape(-1, 3)
        sdf = np.ascontiguousarray(sdf, np.float32); w = np.ascontiguousarray(w, np.float32); rgb = np.ascontiguousarray(rgb, np.uint8)
        self.L.emu_upload_blocks.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 3
        rc = self.L.emu_upload_blocks(self.h, keys.ctypes.data, len(keys), sdf.ctypes.data, w.ctypes.data, rgb.ctypes.data)
        if check:
            assert rc == 0, f"map error flags 0x{rc:x}"
        return rc

    def rebuild_table(self):
        self.L.emu_rebuild_table.argtypes = [C.c_void_p]
        n = self.L.emu_rebuild_table(self.h)
        assert n >= 0
        return n

    free_slots = property(lambda s: s.L.emu_free_slots(s.h))

    def full_map_mc(self):
        """vh_extract_mesh(VH_MESH_FULL_MAP): every allocated block meshed against the whole map; returns the triangle count"""
        n = self.L.emu_full_map_mc(self.h)
        assert n >= 0, "triangle arena overflow"
        return n

    def block_triangles(self, keys, full_map=False):
        keys = np.ascontiguousarray(keys, np.int32).reshape(-1, 3)
        n = self.L.emu_block_triangles(self.h, keys.ctypes.data, len(keys), None, None, int(full_map))
        assert n >= 0
        xyz = np.zeros((max(n, 1), 3, 3), np.float32); rgb = np.zeros((max(n, 1), 3, 3), np.uint8)
        self.L.emu_block_triangles(self.h, keys.ctypes.data, len(keys), xyz.ctypes.data, rgb.ctypes.data, int(full_map))
        return xyz[:n], rgb[:n]


def mesh_order(keys, blocks_per_chunk=8):
    """tsdf2mesh's block order (tsdf.cu:1786-1806): chunks x, y, z ascending, then the block inside the chunk"""
    keys = np.asarray(keys, np.int64).reshape(-1, 3)
    ch = np.floor_divide(keys, blocks_per_chunk)
    order = np.lexsort((keys[:, 2], keys[:, 1], keys[:, 0], ch[:, 2], ch[:, 1], ch[:, 0]))
    return keys[order].astype(np.int32)


class EmuGroup:
    """N emulated ranks of one sharded map (vh_shard_connect / vh_integrate_sharded): phase 1 on every rank, barrier,
    phase 2 on every rank with the peers' tables and planes in reach."""

    def __init__(self, make_params, n, **kw):
        self.ranks = [EmuEngine(make_params(r, n), **kw) for r in range(n)]
        arr = (C.c_void_p * n)(*[e.h for e in self.ranks])
        rc = self.ranks[0].L.emu_connect(arr, n)
        assert rc == 0, f"emu_connect failed: {rc}"

    def close(self):
        for e in self.ranks:
            e.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def process_frame(self, depth, rgb, c2w):
        for e in self.ranks:              # allocation revision 2: every rank routes its share of the rays' keys, then the frame barrier
            e.phase_keys(depth, c2w)
        for e in self.ranks:
            e.phase_integrate(depth, rgb, c2w)
        for e in self.ranks:
            e.phase_mc()


def compare_march(params, depth, c2w):
    """step-by-step DDA vs the merge formulation (allocation revision 1) over every sampled ray of one frame:
    returns (mismatching steps, steps compared, non-empty keys)"""
    L = lib()
    L.emu_compare_march.argtypes = [C.c_void_p] * 4
    depth = np.ascontiguousarray(depth, np.float32); c2w = np.ascontiguousarray(c2w, np.float32)
    out = np.zeros(3, np.uint64)
    L.emu_compare_march(C.addressof(params), depth.ctypes.data, c2w.ctypes.data, out.ctypes.data)
    return int(out[0]), int(out[1]), int(out[2])
