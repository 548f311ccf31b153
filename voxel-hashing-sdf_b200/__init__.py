"""voxel-hashing-sdf_b200 — B200-native voxel-hashing TSDF engine (host-side Python mirror).

The product is lib/libvhsdf.so: hand-written CUDA for sm_100a behind the C ABI of include/vh_c.h.
This module is a thin ctypes mirror of that ABI used by the tests and by bench.py; it adds no compute.
`TsdfEngine` keeps the names of the reference's C++ entry points (ark::GpuTsdfGenerator::processFrame /
SavePLY, /root/reference/include/tsdf.cuh:604-643); the C++ drop-in classes live in include/tsdf.cuh.

There is NO CPU fallback: importing works anywhere (so the symbol table can be checked on a CPU box),
but creating an engine without the built library or without a B200 raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libvhsdf.so")

VH_MESH_REF_PERSISTENT = 0
VH_MESH_FULL_MAP = 1

STATUS = {0: "VH_OK", 1: "VH_ERR_INVALID", 2: "VH_ERR_NO_DEVICE", 3: "VH_ERR_CUDA", 4: "VH_ERR_TABLE_FULL",
          5: "VH_ERR_POOL_FULL", 6: "VH_ERR_ARENA_FULL", 7: "VH_ERR_IO", 8: "VH_ERR_NOT_FOUND"}

# every symbol include/vh_c.h declares (checked by tests/test_abi.py against the header and the .so)
ABI_SYMBOLS = [
    "vh_last_error", "vh_version", "vh_default_params", "vh_create", "vh_destroy", "vh_reset",
    "vh_integrate", "vh_integrate_async", "vh_integrate_u16_async", "vh_wait_uploads", "vh_sync", "vh_integrate_device",
    "vh_upload_frame", "vh_stage_allocate", "vh_stage_integrate", "vh_stage_marching_cubes", "vh_set_visible",
    "vh_get_stats", "vh_stream", "vh_visible_keys", "vh_allocated_keys", "vh_download_blocks", "vh_voxel_checksum",
    "vh_extract_mesh", "vh_save_ply", "vh_save_ply_binary", "vh_weld_mesh", "vh_host_alloc", "vh_host_free",
    "vh_map_create", "vh_map_destroy", "vh_map_insert", "vh_map_find", "vh_map_erase", "vh_map_size", "vh_map_keys",
    "vh_map_get_view",
    "vh_owner_of_block", "vh_shard_unique_id", "vh_shard_connect", "vh_integrate_sharded", "vh_integrate_sharded_device", "vh_shard_barrier", "vh_shard_gather_mesh", "vh_shard_stats",
    "vh_mesh_order_merge",
    "vh_far_blocks", "vh_blocks_resident", "vh_evict_blocks", "vh_upload_blocks",
]


class VhError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{STATUS.get(code, code)}: {msg}")
        self.code = code


class VhParams(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("min_depth", C.c_float), ("max_depth", C.c_float),
                ("vox_size", C.c_float), ("trunc_margin", C.c_float),
                ("voxels_per_block", C.c_int), ("blocks_per_chunk", C.c_int),
                ("dda_stride", C.c_int), ("max_ray_steps", C.c_int),
                ("chunk_radius", C.c_float), ("max_chunk_num", C.c_int),
                ("num_buckets", C.c_int), ("entries_per_bucket", C.c_int), ("pool_blocks", C.c_int),
                ("use_color", C.c_int), ("mc_per_frame", C.c_int), ("device", C.c_int),
                ("shard_rank", C.c_int), ("shard_count", C.c_int), ("shard_group", C.c_int), ("depth_tile_smem", C.c_int),
                ("tri_arena_bytes", C.c_uint64)]


class VhStats(C.Structure):
    _fields_ = [("frames", C.c_uint64), ("visible_blocks", C.c_uint32), ("allocated_blocks", C.c_uint32),
                ("voxel_updates", C.c_uint64), ("voxel_updates_total", C.c_uint64),
                ("triangles", C.c_uint64), ("arena_triangles", C.c_uint64),
                ("ms_upload", C.c_float), ("ms_alloc", C.c_float), ("ms_integrate", C.c_float), ("ms_mc", C.c_float),
                ("debug_mismatches", C.c_uint64), ("arena_compactions", C.c_uint64), ("forced_syncs", C.c_uint64), ("culled_blocks", C.c_uint64),
                ("ms_cull", C.c_float), ("reserved_f", C.c_float)]


TRI_DTYPE = np.dtype([("xyz0", np.float32, 3), ("rgb0", np.uint8, 4), ("xyz1", np.float32, 3), ("rgb1", np.uint8, 4),
                      ("xyz2", np.float32, 3), ("rgb2", np.uint8, 4)])
VERT_DTYPE = np.dtype([("xyz", np.float32, 3), ("rgb", np.uint8, 4)])
assert TRI_DTYPE.itemsize == 48 and VERT_DTYPE.itemsize == 16


def build(verbose: bool = False) -> str:
    """Compile lib/libvhsdf.so for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.check_call(["make", "-C", _HERE, "-j8"], stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def load_library():
    """dlopen lib/libvhsdf.so and type its entry points. Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU or PyTorch fallback for this engine)")
    L = C.CDLL(LIB_PATH)
    vp, ip = C.c_void_p, C.c_int
    L.vh_last_error.restype = C.c_char_p
    L.vh_version.restype = C.c_char_p
    L.vh_default_params.argtypes = [C.POINTER(VhParams)]
    L.vh_create.argtypes = [C.POINTER(VhParams), C.POINTER(vp)]
    for name in ("vh_destroy", "vh_reset", "vh_wait_uploads", "vh_sync", "vh_stage_marching_cubes"):
        getattr(L, name).argtypes = [vp]
    for name in ("vh_integrate", "vh_integrate_async", "vh_integrate_device"):
        getattr(L, name).argtypes = [vp, vp, vp, vp]
    L.vh_integrate_u16_async.argtypes = [vp, vp, C.c_double, vp, vp]
    L.vh_upload_frame.argtypes = [vp, vp, vp]
    L.vh_stage_allocate.argtypes = [vp, vp, vp]
    L.vh_stage_integrate.argtypes = [vp, vp, vp]
    L.vh_set_visible.argtypes = [vp, vp, ip, vp]
    L.vh_get_stats.argtypes = [vp, C.POINTER(VhStats)]
    L.vh_stream.restype = vp; L.vh_stream.argtypes = [vp]
    L.vh_visible_keys.argtypes = [vp, vp, ip, C.POINTER(ip)]
    L.vh_allocated_keys.argtypes = [vp, vp, ip, C.POINTER(ip)]
    L.vh_download_blocks.argtypes = [vp, vp, ip, vp, vp, vp, vp]
    L.vh_far_blocks.argtypes = [vp, vp, vp, ip, vp]
    L.vh_blocks_resident.argtypes = [vp, vp, vp, ip, vp]
    L.vh_evict_blocks.argtypes = [vp, vp, ip, vp, vp, vp, vp]
    L.vh_upload_blocks.argtypes = [vp, vp, ip, vp, vp, vp]
    L.vh_voxel_checksum.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.vh_extract_mesh.argtypes = [vp, ip, vp, C.c_uint64, C.POINTER(C.c_uint64)]
    L.vh_save_ply.argtypes = [vp, C.c_char_p, ip]
    L.vh_save_ply_binary.argtypes = [vp, C.c_char_p, ip]
    L.vh_weld_mesh.argtypes = [vp, ip, vp, C.c_uint64, C.POINTER(C.c_uint64), vp, C.c_uint64, C.POINTER(C.c_uint64)]
    L.vh_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    L.vh_host_free.argtypes = [vp]
    L.vh_map_create.argtypes = [ip, ip, ip, ip, C.POINTER(vp)]
    L.vh_map_destroy.argtypes = [vp]
    L.vh_map_insert.argtypes = [vp, vp, ip, vp]
    L.vh_map_find.argtypes = [vp, vp, ip, vp]
    L.vh_map_erase.argtypes = [vp, vp, ip, vp]
    L.vh_map_size.argtypes = [vp, C.POINTER(ip)]
    L.vh_map_keys.argtypes = [vp, vp, ip, C.POINTER(ip)]
    L.vh_owner_of_block.argtypes = [ip, ip, ip, ip, ip]
    L.vh_shard_unique_id.argtypes = [vp]
    L.vh_shard_connect.argtypes = [vp, vp]
    L.vh_integrate_sharded.argtypes = [vp, vp, vp, vp]
    L.vh_integrate_sharded_device.argtypes = [vp, vp, vp, vp]
    L.vh_shard_barrier.argtypes = [vp]
    L.vh_shard_gather_mesh.argtypes = [vp, ip, vp, C.c_uint64, C.POINTER(C.c_uint64)]
    L.vh_shard_stats.argtypes = [vp, C.POINTER(VhStats)]
    L.vh_mesh_order_merge.argtypes = [ip, vp, vp, ip, vp, vp]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise VhError(rc, load_library().vh_last_error().decode(errors="replace"))


def default_params(**kw) -> VhParams:
    p = VhParams()
    _check(load_library().vh_default_params(C.byref(p)))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(f"vh_params has no field {k}")
        setattr(p, k, v)
    return p


def blocks_resident(params, c2w, keys):
    """reference residency rule (chunk cube + sphere around the frustum centre) for arbitrary block keys; host arithmetic only"""
    keys = np.ascontiguousarray(keys, np.int32).reshape(-1, 3)
    c2w = np.ascontiguousarray(c2w, np.float32)
    out = np.zeros(max(len(keys), 1), np.uint8)
    _check(load_library().vh_blocks_resident(C.byref(params), _ptr(c2w), _ptr(keys), len(keys), _ptr(out)))
    return out[:len(keys)].astype(bool)


def owner_of_block(x: int, y: int, z: int, shard_count: int, shard_group: int = 8) -> int:
    """which shard of a multi-GPU map owns block (x, y, z); ownership is hashed per cube of shard_group^3 blocks"""
    return load_library().vh_owner_of_block(int(x), int(y), int(z), int(shard_count), int(shard_group))


def mesh_order_merge(parts, blocks_per_chunk: int = 8):
    """parts: list of int32[n_i, 3] block-key arrays, each already in mesh order -> (part, index) arrays of the merged order"""
    parts = [np.ascontiguousarray(p, np.int32).reshape(-1, 3) for p in parts]
    n = sum(len(p) for p in parts)
    ptrs = (C.c_void_p * len(parts))(*[p.ctypes.data for p in parts])
    counts = np.array([len(p) for p in parts], np.int32)
    op, oi = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32)
    _check(load_library().vh_mesh_order_merge(len(parts), C.cast(ptrs, C.c_void_p), counts.ctypes.data, blocks_per_chunk, op.ctypes.data, oi.ctypes.data))
    return op[:n], oi[:n]


def params_for_scene(scene, **kw) -> VhParams:
    return default_params(width=scene.width, height=scene.height, fx=scene.fx, fy=scene.fy, cx=scene.cx, cy=scene.cy, **kw)


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return a
    return a.ctypes.data


class TsdfEngine:
    """Mirror of ark::GpuTsdfGenerator over the C ABI (one map resident on one B200)."""

    def __init__(self, params: VhParams):
        self.L = load_library()
        self.params = params
        self.h = C.c_void_p()
        _check(self.L.vh_create(C.byref(params), C.byref(self.h)))

    # -- lifetime -------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.L.vh_destroy(self.h)
            self.h = None

    Shutdown = close                                    # GpuTsdfGenerator::Shutdown, tsdf.cuh:622

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def reset(self):
        _check(self.L.vh_reset(self.h))

    # -- hot path -------------------------------------------------------------------------------
    def processFrame(self, depth, rgb, c2w):
        """GpuTsdfGenerator::processFrame (tsdf.cuh:610): host buffers, synchronous."""
        depth, rgb, c2w = self._host(depth, rgb, c2w)
        _check(self.L.vh_integrate(self.h, _ptr(depth), _ptr(rgb), _ptr(c2w)))

    process_frame = processFrame

    def integrate_async(self, depth, rgb, c2w):
        """Enqueue one frame; buffers (numpy arrays or raw pinned addresses) must stay alive until sync()."""
        _check(self.L.vh_integrate_async(self.h, _ptr(depth), _ptr(rgb), _ptr(c2w)))

    def integrate_u16_async(self, depth_u16, depth_scale, rgb, c2w):
        """u16 depth samples (e.g. millimetres with depth_scale 0.001), converted on the GPU; buffers must stay alive until sync()"""
        self._keep = (depth_u16, rgb, c2w)
        _check(self.L.vh_integrate_u16_async(self.h, _ptr(depth_u16), float(depth_scale), _ptr(rgb), _ptr(c2w)))

    def integrate_device(self, d_depth: int, d_rgb, c2w):
        _check(self.L.vh_integrate_device(self.h, d_depth, d_rgb, _ptr(c2w)))

    def sync(self):
        _check(self.L.vh_sync(self.h))

    def wait_uploads(self):
        _check(self.L.vh_wait_uploads(self.h))

    def _host(self, depth, rgb, c2w):
        depth = np.ascontiguousarray(depth, np.float32)
        c2w = np.ascontiguousarray(c2w, np.float32)
        assert depth.size == self.params.width * self.params.height and c2w.size == 16
        if rgb is not None:
            rgb = np.ascontiguousarray(rgb, np.uint8)
            assert rgb.size == depth.size * 3
        return depth, rgb, c2w

    # -- multi-GPU (one engine per process/GPU; see include/vh_c.h) --------------------------------
    @staticmethod
    def shard_unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        _check(load_library().vh_shard_unique_id(buf))
        return bytes(buf)

    def shard_connect(self, unique_id: bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        _check(self.L.vh_shard_connect(self.h, buf))

    def integrate_sharded(self, depth, rgb, c2w):
        """collective; depth/rgb may be None on ranks > 0"""
        if depth is not None and not isinstance(depth, int):
            depth, rgb, c2w = self._host(depth, rgb, c2w)
        elif c2w is not None:
            c2w = np.ascontiguousarray(c2w, np.float32)
        self._keep = (depth, rgb, c2w)
        _check(self.L.vh_integrate_sharded(self.h, _ptr(depth), _ptr(rgb), _ptr(c2w)))

    def integrate_sharded_device(self, d_depth, d_rgb, c2w):
        """collective; device pointers (ints) of a frame resident in rank 0's HBM, None on ranks > 0"""
        c2w = np.ascontiguousarray(c2w, np.float32)
        self._keep = (c2w,)
        _check(self.L.vh_integrate_sharded_device(self.h, d_depth, d_rgb, _ptr(c2w)))

    def shard_barrier(self):
        _check(self.L.vh_shard_barrier(self.h))

    def shard_triangles(self, mode=VH_MESH_REF_PERSISTENT):
        """collective; the merged soup on rank 0, empty arrays elsewhere"""
        n = C.c_uint64()
        _check(self.L.vh_shard_gather_mesh(self.h, mode, None, 0, C.byref(n)))
        buf = np.zeros(max(n.value, 1), TRI_DTYPE)
        _check(self.L.vh_shard_gather_mesh(self.h, mode, _ptr(buf), max(n.value, 1), C.byref(n)))
        buf = buf[:n.value]
        xyz = np.stack([buf["xyz0"], buf["xyz1"], buf["xyz2"]], 1) if n.value else np.zeros((0, 3, 3), np.float32)
        rgb = np.stack([buf["rgb0"][:, :3], buf["rgb1"][:, :3], buf["rgb2"][:, :3]], 1) if n.value else np.zeros((0, 3, 3), np.uint8)
        return xyz, rgb

    def shard_stats(self) -> VhStats:
        s = VhStats()
        _check(self.L.vh_shard_stats(self.h, C.byref(s)))
        return s

    # -- stages ---------------------------------------------------------------------------------
    def upload_frame(self, depth, rgb=None):
        depth = np.ascontiguousarray(depth, np.float32)
        rgb = None if rgb is None else np.ascontiguousarray(rgb, np.uint8)
        _check(self.L.vh_upload_frame(self.h, _ptr(depth), _ptr(rgb)))

    def stage_allocate(self, c2w, d_depth=None):
        c2w = np.ascontiguousarray(c2w, np.float32)
        _check(self.L.vh_stage_allocate(self.h, d_depth, _ptr(c2w)))

    def stage_integrate(self, d_depth=None, d_rgb=None):
        _check(self.L.vh_stage_integrate(self.h, d_depth, d_rgb))

    def stage_marching_cubes(self):
        _check(self.L.vh_stage_marching_cubes(self.h))

    def set_visible(self, keys, c2w):
        keys = np.ascontiguousarray(keys, np.int32).reshape(-1, 3)
        c2w = np.ascontiguousarray(c2w, np.float32)
        _check(self.L.vh_set_visible(self.h, _ptr(keys), len(keys), _ptr(c2w)))

    # -- inspection -----------------------------------------------------------------------------
    def stats(self) -> VhStats:
        s = VhStats()
        _check(self.L.vh_get_stats(self.h, C.byref(s)))
        return s

    @property
    def stream(self) -> int:
        return self.L.vh_stream(self.h)

    def visible_keys(self):
        n = C.c_int()
        _check(self.L.vh_visible_keys(self.h, None, 0, C.byref(n)))
        out = np.zeros((max(n.value, 1), 3), np.int32)
        _check(self.L.vh_visible_keys(self.h, _ptr(out), n.value, C.byref(n)))
        return out[:n.value]

    def allocated_keys(self):
        n = C.c_int()
        _check(self.L.vh_allocated_keys(self.h, None, 0, C.byref(n)))
        out = np.zeros((max(n.value, 1), 3), np.int32)
        _check(self.L.vh_allocated_keys(self.h, _ptr(out), n.value, C.byref(n)))
        return out[:n.value]

    def download_blocks(self, keys, want_rgb=True):
        keys = np.ascontiguousarray(keys, np.int32).reshape(-1, 3)
        n = len(keys)
        sdf = np.zeros((n, 512), np.float32)
        w = np.zeros((n, 512), np.float32)
        rgb = np.zeros((n, 512, 3), np.uint8) if want_rgb else None
        found = np.zeros(n, np.uint8)
        _check(self.L.vh_download_blocks(self.h, _ptr(keys), n, _ptr(sdf), _ptr(w), _ptr(rgb), _ptr(found)))
        return sdf, w, rgb, found.astype(bool)

    # -- out-of-core tier (between frames) ----------------------------------------------------------
    def far_blocks(self, c2w):
        """allocated blocks whose chunk fails the reference's residency rule for this pose (tsdf.cu:166-187,300-312)"""
        c2w = np.ascontiguousarray(c2w, np.float32)
        n = C.c_int(0)
        _check(self.L.vh_far_blocks(self.h, _ptr(c2w), None, 0, C.byref(n)))
        out = np.zeros((max(n.value, 1), 3), np.int32)
        _check(self.L.vh_far_blocks(self.h, _ptr(c2w), _ptr(out), n.value, C.byref(n)))
        return out[:n.value]

    def evict_blocks(self, keys, want_rgb=True):
        """download the blocks, then release them from the map: (sdf, weight, rgb, found)"""
        keys = np.ascontiguousarray(keys, np.int32).reshape(-1, 3)
        n = len(keys)
        sdf = np.zeros((n, 512), np.float32)
        w = np.zeros((n, 512), np.float32)
        rgb = np.zeros((n, 512, 3), np.uint8) if want_rgb else None
        found = np.zeros(n, np.uint8)
        _check(self.L.vh_evict_blocks(self.h, _ptr(keys), n, _ptr(sdf), _ptr(w), _ptr(rgb), _ptr(found)))
        return sdf, w, rgb, found.astype(bool)

    def upload_blocks(self, keys, sdf, weight, rgb=None):
        keys = np.ascontiguousarray(keys, np.int32).reshape(-1, 3)
        sdf = np.ascontiguousarray(sdf, np.float32); weight = np.ascontiguousarray(weight, np.float32)
        rgb = None if rgb is None else np.ascontiguousarray(rgb, np.uint8)
        assert sdf.size == len(keys) * 512 and weight.size == len(keys) * 512 and (rgb is None or rgb.size == len(keys) * 512 * 3)
        _check(self.L.vh_upload_blocks(self.h, _ptr(keys), len(keys), _ptr(sdf), _ptr(weight), _ptr(rgb)))

    def checksum(self):
        ss, sw, no, nn = C.c_double(), C.c_double(), C.c_uint64(), C.c_uint64()
        _check(self.L.vh_voxel_checksum(self.h, C.byref(ss), C.byref(sw), C.byref(no), C.byref(nn)))
        return dict(sum_sdf=ss.value, sum_w=sw.value, n_observed=no.value, n_negative=nn.value)

    # -- mesh -----------------------------------------------------------------------------------
    def triangles(self, mode=VH_MESH_REF_PERSISTENT):
        """Ordered triangle soup (tsdf2mesh order), voxel-index units: (xyz float32[n,3,3], rgb uint8[n,3,3])."""
        n = C.c_uint64()
        _check(self.L.vh_extract_mesh(self.h, mode, None, 0, C.byref(n)))
        buf = np.zeros(max(n.value, 1), TRI_DTYPE)
        if n.value:
            _check(self.L.vh_extract_mesh(self.h, mode, _ptr(buf), n.value, C.byref(n)))
        buf = buf[:n.value]
        xyz = np.stack([buf["xyz0"], buf["xyz1"], buf["xyz2"]], 1) if n.value else np.zeros((0, 3, 3), np.float32)
        rgb = np.stack([buf["rgb0"][:, :3], buf["rgb1"][:, :3], buf["rgb2"][:, :3]], 1) if n.value else np.zeros((0, 3, 3), np.uint8)
        return xyz, rgb

    def weld(self, mode=VH_MESH_REF_PERSISTENT):
        nv, nf = C.c_uint64(), C.c_uint64()
        _check(self.L.vh_weld_mesh(self.h, mode, None, 0, C.byref(nv), None, 0, C.byref(nf)))
        verts = np.zeros(max(nv.value, 1), VERT_DTYPE)
        faces = np.zeros((max(nf.value, 1), 3), np.int32)
        _check(self.L.vh_weld_mesh(self.h, mode, _ptr(verts), nv.value, C.byref(nv), _ptr(faces), nf.value, C.byref(nf)))
        return verts[:nv.value], faces[:nf.value]

    def SavePLY(self, path: str, mode=VH_MESH_REF_PERSISTENT):
        """GpuTsdfGenerator::SavePLY (tsdf.cuh:628): ASCII PLY with exact-xyz vertex dedupe."""
        _check(self.L.vh_save_ply(self.h, path.encode(), mode))

    save_ply = SavePLY

    def save_ply_binary(self, path: str, mode=VH_MESH_REF_PERSISTENT):
        _check(self.L.vh_save_ply_binary(self.h, path.encode(), mode))


class BlockHashMap:
    """Mirror of vhashing::HashTable<int3, ...> bulk operations over vh_map_* (vhashing.h:531-603)."""

    def __init__(self, num_buckets: int, entries_per_bucket: int, num_blocks: int, device: int = 0):
        self.L = load_library()
        self.h = C.c_void_p()
        _check(self.L.vh_map_create(num_buckets, entries_per_bucket, num_blocks, device, C.byref(self.h)))

    def close(self):
        if getattr(self, "h", None):
            self.L.vh_map_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def insert(self, keys):
        keys = np.ascontiguousarray(keys, np.int32).reshape(-1, 3)
        out = np.full(len(keys), -1, np.int32)
        _check(self.L.vh_map_insert(self.h, _ptr(keys), len(keys), _ptr(out)))
        return out

    def find(self, keys):
        keys = np.ascontiguousarray(keys, np.int32).reshape(-1, 3)
        out = np.full(len(keys), -1, np.int32)
        _check(self.L.vh_map_find(self.h, _ptr(keys), len(keys), _ptr(out)))
        return out

    def erase(self, keys):
        keys = np.ascontiguousarray(keys, np.int32).reshape(-1, 3)
        out = np.zeros(len(keys), np.int32)
        _check(self.L.vh_map_erase(self.h, _ptr(keys), len(keys), _ptr(out)))
        return out

    def __len__(self):
        n = C.c_int()
        _check(self.L.vh_map_size(self.h, C.byref(n)))
        return n.value

    def keys(self):
        n = C.c_int()
        _check(self.L.vh_map_keys(self.h, None, 0, C.byref(n)))
        out = np.zeros((max(n.value, 1), 3), np.int32)
        _check(self.L.vh_map_keys(self.h, _ptr(out), n.value, C.byref(n)))
        return out[:n.value]
