"""ctypes bindings for the CPU oracle (libvh_oracle.so) and the emulated reference (oracle/_ref).

TEST INFRASTRUCTURE. Imported only by tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py. The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class VoParams(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("min_depth", C.c_float), ("max_depth", C.c_float),
                ("vox_size", C.c_float), ("trunc_margin", C.c_float),
                ("voxels_per_block", C.c_int), ("blocks_per_chunk", C.c_int),
                ("dda_stride", C.c_int), ("max_ray_steps", C.c_int),
                ("chunk_radius", C.c_float), ("max_chunk_num", C.c_int),
                ("use_color", C.c_int), ("run_mc", C.c_int), ("num_threads", C.c_int)]


def build(force: bool = False) -> str:
    so = os.path.join(HERE, "libvh_oracle.so")
    src = os.path.join(HERE, "vh_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "libvh_oracle.so"], stdout=subprocess.DEVNULL)
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp = C.c_void_p
        L.vo_default_params.argtypes = [C.POINTER(VoParams)]
        L.vo_create.restype = vp; L.vo_create.argtypes = [C.POINTER(VoParams)]
        L.vo_destroy.argtypes = [vp]
        L.vo_process_frame.argtypes = [vp, vp, vp, vp]
        L.vo_begin_frame.argtypes = [vp, vp]
        L.vo_stage_allocate.argtypes = [vp, vp]
        L.vo_stage_integrate.restype = C.c_longlong; L.vo_stage_integrate.argtypes = [vp, vp, vp]
        L.vo_stage_mc.restype = C.c_longlong; L.vo_stage_mc.argtypes = [vp]
        L.vo_num_visible.argtypes = [vp]
        L.vo_last_updates.restype = C.c_longlong; L.vo_last_updates.argtypes = [vp]
        L.vo_last_triangles.restype = C.c_longlong; L.vo_last_triangles.argtypes = [vp]
        L.vo_streamed_blocks.restype = C.c_longlong; L.vo_streamed_blocks.argtypes = [vp]
        L.vo_last_times.argtypes = [vp, vp]
        L.vo_visible_keys.argtypes = [vp, vp, C.c_int]
        L.vo_num_blocks.restype = C.c_longlong; L.vo_num_blocks.argtypes = [vp]
        L.vo_all_keys.restype = C.c_longlong; L.vo_all_keys.argtypes = [vp, vp, C.c_longlong]
        L.vo_get_block.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp]
        L.vo_get_blocks.argtypes = [vp, vp, C.c_longlong, vp, vp, vp, vp]
        L.vo_voxel_checksum.argtypes = [vp, vp, vp, vp, vp]
        L.vo_triangles.restype = C.c_longlong; L.vo_triangles.argtypes = [vp, vp, vp, C.c_longlong]
        L.vo_visible_tri_counts.argtypes = [vp, vp]
        L.vo_frame2base.argtypes = [C.POINTER(VoParams), vp, C.c_int, C.c_int, C.c_float, vp]
        L.vo_base2cam.argtypes = [vp, vp, vp]
        L.vo_cam2frame.argtypes = [C.POINTER(VoParams), vp, vp]
        L.vo_vertex_interp.argtypes = [vp, vp, C.c_float, C.c_float, vp]
        L.vo_block_hash.restype = C.c_ulonglong; L.vo_block_hash.argtypes = [C.c_int] * 3
        L.vo_block_is_candidate.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int]
        L.vo_block_in_frustum.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int]
        _lib = L
    return _lib


def default_params(**kw) -> VoParams:
    p = VoParams()
    lib().vo_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def params_for_scene(scene, **kw) -> VoParams:
    return default_params(width=scene.width, height=scene.height, fx=scene.fx, fy=scene.fy, cx=scene.cx, cy=scene.cy, **kw)


def _ptr(a):
    return None if a is None else a.ctypes.data


class Oracle:
    """CPU restatement of GpuTsdfGenerator::processFrame + tsdf2mesh (see vh_oracle.c header)."""

    def __init__(self, params: VoParams):
        self.L = lib()
        self.params = params
        self.nvox = params.voxels_per_block ** 3
        self.h = self.L.vo_create(C.byref(params))

    def close(self):
        if self.h:
            self.L.vo_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def process_frame(self, depth, rgb, c2w):
        depth = np.ascontiguousarray(depth, np.float32)
        c2w = np.ascontiguousarray(c2w, np.float32)
        rgb = None if rgb is None else np.ascontiguousarray(rgb, np.uint8)
        self.L.vo_process_frame(self.h, _ptr(depth), _ptr(rgb), _ptr(c2w))

    def begin_frame(self, c2w):
        c2w = np.ascontiguousarray(c2w, np.float32)
        self.L.vo_begin_frame(self.h, _ptr(c2w))

    def stage_allocate(self, depth):
        depth = np.ascontiguousarray(depth, np.float32)
        self.L.vo_stage_allocate(self.h, _ptr(depth))

    def stage_integrate(self, depth, rgb):
        depth = np.ascontiguousarray(depth, np.float32)
        rgb = None if rgb is None else np.ascontiguousarray(rgb, np.uint8)
        return self.L.vo_stage_integrate(self.h, _ptr(depth), _ptr(rgb))

    def stage_mc(self):
        return self.L.vo_stage_mc(self.h)

    def full_map_mc(self):
        """meshes every allocated block against the whole map (replaces the per-frame meshes: call it last); returns the triangle count"""
        self.L.vo_full_map_mc.restype = C.c_longlong
        self.L.vo_full_map_mc.argtypes = [C.c_void_p]
        return self.L.vo_full_map_mc(self.h)

    @property
    def num_visible(self):
        return self.L.vo_num_visible(self.h)

    @property
    def last_updates(self):
        return self.L.vo_last_updates(self.h)

    @property
    def last_triangles(self):
        return self.L.vo_last_triangles(self.h)

    @property
    def streamed_blocks(self):
        return self.L.vo_streamed_blocks(self.h)

    def last_times(self):
        t = np.zeros(3, np.float64)
        self.L.vo_last_times(self.h, _ptr(t))
        return dict(alloc=t[0], integrate=t[1], mc=t[2])

    def visible_keys(self):
        n = self.num_visible
        out = np.zeros((max(n, 1), 3), np.int32)
        self.L.vo_visible_keys(self.h, _ptr(out), n)
        return out[:n]

    def all_keys(self):
        n = self.L.vo_num_blocks(self.h)
        out = np.zeros((max(n, 1), 3), np.int32)
        self.L.vo_all_keys(self.h, _ptr(out), n)
        return out[:n]

    def get_blocks(self, keys):
        keys = np.ascontiguousarray(keys, np.int32).reshape(-1, 3)
        n = len(keys)
        sdf = np.zeros((n, self.nvox), np.float32)
        w = np.zeros((n, self.nvox), np.float32)
        rgb = np.zeros((n, self.nvox, 3), np.uint8)
        found = np.zeros(n, np.uint8)
        self.L.vo_get_blocks(self.h, _ptr(keys), n, _ptr(sdf), _ptr(w), _ptr(rgb), _ptr(found))
        return sdf, w, rgb, found.astype(bool)

    def checksum(self):
        ss, sw = C.c_double(), C.c_double()
        no, nn = C.c_longlong(), C.c_longlong()
        self.L.vo_voxel_checksum(self.h, C.byref(ss), C.byref(sw), C.byref(no), C.byref(nn))
        return dict(sum_sdf=ss.value, sum_w=sw.value, n_observed=no.value, n_negative=nn.value)

    def triangles(self):
        n = self.L.vo_triangles(self.h, None, None, 0)
        xyz = np.zeros((max(n, 1), 3, 3), np.float32)
        rgb = np.zeros((max(n, 1), 3, 3), np.uint8)
        self.L.vo_triangles(self.h, _ptr(xyz), _ptr(rgb), n)
        return xyz[:n], rgb[:n]

    def visible_tri_counts(self):
        out = np.zeros(max(self.num_visible, 1), np.int32)
        self.L.vo_visible_tri_counts(self.h, _ptr(out))
        return out[:self.num_visible]


# ---------------------------------------------------------------------------------------------
def ref_emu_path(vpb: int) -> str:
    return os.path.join(HERE, "_ref", f"libref_emu_vpb{vpb}.so")


def ref_emu_available(vpb: int) -> bool:
    return os.path.exists(ref_emu_path(vpb))


def ref_cuda_path(variant: str) -> str:
    """the reference's own tsdf.cu built by oracle/build_ref_cuda.sh as a real CUDA library (baseline timing on the GPU box)"""
    return os.path.join(HERE, "_ref", f"libref_cuda_vpb{variant}.so")


class RefEmu:
    """The reference's own GpuTsdfGenerator (src/tsdf.cu) running sequentially on the CPU.

    Built by oracle/build_ref.sh from /root/reference; one engine per process (the reference keeps
    its tables in file-scope globals, tsdf.cu:19-20). Constants other than VOXEL_PER_BLOCK are the
    reference's: 8 blocks per chunk, DDA stride 10, 100 ray steps, +-64 chunks, min depth 0.1.
    """

    def __init__(self, scene, vpb: int, vox_size: float, trunc: float, max_depth: float = 10.0, lib_path: str = None):
        L = C.CDLL(lib_path or ref_emu_path(vpb))
        vp = C.c_void_p
        L.ref_create.restype = vp
        L.ref_create.argtypes = [C.c_int] * 2 + [C.c_float] * 7
        L.ref_process_frame.argtypes = [vp] * 4
        L.ref_last_visible_count.argtypes = [vp]
        L.ref_last_streamed_blocks.argtypes = [vp]
        L.ref_last_visible_keys.argtypes = [vp, vp, C.c_int]
        L.ref_get_block.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp]
        L.ref_voxel_checksum.argtypes = [vp] * 5
        L.ref_triangles.restype = C.c_longlong
        L.ref_triangles.argtypes = [vp, vp, vp, C.c_longlong]
        L.ref_save_ply.argtypes = [vp, C.c_char_p]
        L.ref_frame2cam.argtypes = [C.c_int, C.c_int, C.c_float, vp, vp]
        L.ref_cam2frame.argtypes = [vp, vp, vp]
        L.ref_base2cam.argtypes = [vp, vp, vp]
        L.ref_cam2base.argtypes = [vp, vp, vp]
        L.ref_vertex_interp.argtypes = [vp, vp, C.c_float, C.c_float, vp]
        L.ref_block_hash.restype = C.c_ulonglong
        L.ref_block_hash.argtypes = [C.c_int] * 3
        assert L.ref_voxels_per_block() == vpb
        self.L = L
        self.nvox = vpb ** 3
        self.h = L.ref_create(scene.width, scene.height, scene.fx, scene.fy, scene.cx, scene.cy, max_depth, vox_size, trunc)

    def process_frame(self, depth, rgb, c2w):
        depth = np.ascontiguousarray(depth, np.float32)
        rgb = np.ascontiguousarray(rgb, np.uint8)
        c2w = np.ascontiguousarray(c2w, np.float32)
        self.L.ref_process_frame(self.h, _ptr(depth), _ptr(rgb), _ptr(c2w))

    @property
    def num_visible(self):
        return self.L.ref_last_visible_count(self.h)

    @property
    def streamed_blocks(self):
        return self.L.ref_last_streamed_blocks(self.h)

    def visible_keys(self):
        n = self.num_visible
        out = np.zeros((max(n, 1), 3), np.int32)
        m = self.L.ref_last_visible_keys(self.h, _ptr(out), n)
        assert m == n
        return out[:n]

    def get_blocks(self, keys):
        keys = np.ascontiguousarray(keys, np.int32).reshape(-1, 3)
        n = len(keys)
        sdf = np.zeros((n, self.nvox), np.float32)
        w = np.zeros((n, self.nvox), np.float32)
        rgb = np.zeros((n, self.nvox, 3), np.uint8)
        found = np.zeros(n, bool)
        for i, (x, y, z) in enumerate(keys):
            found[i] = bool(self.L.ref_get_block(self.h, int(x), int(y), int(z), _ptr(sdf[i]), _ptr(w[i]), _ptr(rgb[i])))
        return sdf, w, rgb, found

    def checksum(self):
        ss, sw = C.c_double(), C.c_double()
        no, nn = C.c_longlong(), C.c_longlong()
        self.L.ref_voxel_checksum(self.h, C.byref(ss), C.byref(sw), C.byref(no), C.byref(nn))
        return dict(sum_sdf=ss.value, sum_w=sw.value, n_observed=no.value, n_negative=nn.value)

    def triangles(self):
        n = self.L.ref_triangles(self.h, None, None, 0)
        xyz = np.zeros((max(n, 1), 3, 3), np.float32)
        rgb = np.zeros((max(n, 1), 3, 3), np.uint8)
        self.L.ref_triangles(self.h, _ptr(xyz), _ptr(rgb), n)
        return xyz[:n], rgb[:n]

    def save_ply(self, path: str):
        self.L.ref_save_ply(self.h, path.encode())


class RefCuda(RefEmu):
    """The reference's own CUDA build (oracle/build_ref_cuda.sh) on a real GPU: processFrame unchanged, results read out of
    its host chunk store. The per-frame key hook of the emulated build does not exist here (key_heap lives on the device)."""

    def __init__(self, scene, variant: str, vox_size: float, trunc: float, max_depth: float = 10.0):
        super().__init__(scene, 8 if variant.startswith("8") else 5, vox_size, trunc, max_depth, lib_path=ref_cuda_path(variant))

    def visible_keys(self):
        raise NotImplementedError("the CUDA build keeps key_heap on the device")

    def triangle_count(self):
        return int(self.L.ref_triangles(self.h, None, None, 0))
