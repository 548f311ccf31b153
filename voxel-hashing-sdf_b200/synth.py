"""Synthetic depth sequences for the voxel-hashing TSDF path.

The reference ships no data (its scene0220_02/ holds only the YAML), so every parity vector and
every benchmark frame is generated here: an axis-aligned box room (optionally with spheres inside),
seen by a pinhole camera moving on a horizontal circle about the room centre and looking outward.

Conventions match what the reference's loader hands to processFrame
(/root/reference/src/SaveFrame.cpp:174-180, src/PointCloudGenerator.cpp:118-125):
  depth  float32[H, W] metres, z-depth along the optical axis, quantised to millimetres, 0 = invalid
  rgb    uint8[H, W, 3]
  c2w    float32[16] row-major camera->world, columns = (right, down, forward), last row 0 0 0 1
Intrinsics follow scene0220_02.yaml:11-14 scaled to the image width: fx = fy = 577*W/640.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np


@dataclass
class Scene:
    width: int = 640
    height: int = 480
    room: tuple = (8.0, 6.0, 3.0)          # extent along x, y, z (z is up)
    room_min: tuple = (0.0, 0.0, 0.0)      # minimum corner (ScanNet-like positive octant by default)
    spheres: tuple = ()                    # ((cx, cy, cz, r), ...) in world coordinates
    n_frames: int = 100
    phase: float = 0.0                     # trajectory phase shift in radians (config 5)
    radius_frac: float = 0.25              # circle radius as a fraction of min(Lx, Ly)
    holes: float = 0.0                     # fraction of pixels zeroed (seeded)
    seed: int = 220
    color: bool = False                    # False: rgb = 0 (parity runs); True: checkerboard
    fx: float = field(init=False)
    fy: float = field(init=False)
    cx: float = field(init=False)
    cy: float = field(init=False)

    def __post_init__(self):
        self.fx = self.fy = 577.0 * self.width / 640.0
        self.cx = self.width / 2.0
        self.cy = self.height / 2.0

    # -- camera ---------------------------------------------------------------------------------
    def pose(self, i: int) -> np.ndarray:
        """Row-major 4x4 camera-to-world of frame i as float32[16]."""
        lx, ly, lz = self.room
        mx, my, mz = self.room_min
        th = 2.0 * math.pi * i / self.n_frames + self.phase
        c, s = math.cos(th), math.sin(th)
        r = self.radius_frac * min(lx, ly)
        pos = np.array([mx + lx / 2 + r * c, my + ly / 2 + r * s, mz + lz / 2])
        fwd = np.array([c, s, 0.0])
        down = np.array([0.0, 0.0, -1.0])
        right = np.cross(down, fwd)
        m = np.eye(4)
        m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = right, down, fwd, pos
        return m.astype(np.float32).reshape(16)

    # -- images ---------------------------------------------------------------------------------
    def frame(self, i: int):
        """Returns (depth float32[H,W], rgb uint8[H,W,3], c2w float32[16]) for frame i."""
        c2w32 = self.pose(i)
        m = c2w32.astype(np.float64).reshape(4, 4)
        rot, org = m[:3, :3], m[:3, 3]
        u = np.arange(self.width, dtype=np.float64)
        v = np.arange(self.height, dtype=np.float64)
        uu, vv = np.meshgrid(u, v)
        dcam = np.stack([(uu - self.cx) / self.fx, (vv - self.cy) / self.fy, np.ones_like(uu)], -1)
        d = dcam @ rot.T                                           # world direction per unit z-depth
        lo = np.array(self.room_min, dtype=np.float64)
        hi = lo + np.array(self.room, dtype=np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            t_hi = np.where(d > 0, (hi - org) / d, np.inf)
            t_lo = np.where(d < 0, (lo - org) / d, np.inf)
        t = np.minimum(t_hi, t_lo).min(-1)                         # exit distance of the box, in z-depth units
        for (sx, sy, sz, sr) in self.spheres:
            oc = org - np.array([sx, sy, sz])
            a = (d * d).sum(-1)
            b = 2.0 * (d @ oc)
            cc = oc @ oc - sr * sr
            disc = b * b - 4 * a * cc
            hit = disc > 0
            ts = np.where(hit, (-b - np.sqrt(np.where(hit, disc, 0.0))) / (2 * a), np.inf)
            ts = np.where(ts > 1e-6, ts, np.inf)
            t = np.minimum(t, ts)
        depth = (np.round(t * 1000.0) / 1000.0).astype(np.float32)  # u16-millimetre PNG path
        if self.holes > 0:
            rng = np.random.RandomState(self.seed + 7919 * i)
            depth[rng.random_sample(depth.shape) < self.holes] = 0.0
        if self.color:
            p = org + d * t[..., None]
            chk = (np.floor(p[..., 0] * 2) + np.floor(p[..., 1] * 2) + np.floor(p[..., 2] * 2)).astype(np.int64) & 1
            rgb = np.empty((self.height, self.width, 3), np.uint8)
            rgb[..., 0] = np.where(chk == 1, 220, 40)
            rgb[..., 1] = (np.clip(p[..., 2] / max(self.room[2], 1e-6), 0, 1) * 255).astype(np.uint8)
            rgb[..., 2] = np.where(chk == 1, 60, 200)
        else:
            rgb = np.zeros((self.height, self.width, 3), np.uint8)
        return np.ascontiguousarray(depth), np.ascontiguousarray(rgb), c2w32


# BASELINE.json configs as concrete parameter sets. trunc = 5 x voxel unless the
# config names it; hash shape = num_buckets x 4 entries as in the reference (tsdf.cu:1488).
CONFIGS = {
    "C1": dict(scene=dict(width=640, height=480, room=(8.0, 6.0, 3.0), n_frames=100),
               vox_size=0.01, trunc=0.05, num_buckets=1 << 20, max_depth=10.0),
    "C2": dict(scene=dict(width=640, height=480, room=(8.0, 6.0, 3.0), n_frames=500),
               vox_size=0.005, trunc=0.025, num_buckets=1 << 20, max_depth=10.0),
    "C3": dict(scene=dict(width=1280, height=720, room=(8.0, 6.0, 3.0), n_frames=100),
               vox_size=0.004, trunc=0.03, num_buckets=1 << 24, max_depth=10.0),
    "C4": dict(scene=dict(width=640, height=480, room=(10.0, 10.0, 3.0), room_min=(-5.0, -5.0, -1.5), n_frames=100),
               vox_size=0.002, trunc=0.01, num_buckets=1 << 24, max_depth=10.0),
}


def make_scene(name: str, **over) -> Scene:
    kw = dict(CONFIGS[name]["scene"])
    kw.update(over)
    return Scene(**kw)
