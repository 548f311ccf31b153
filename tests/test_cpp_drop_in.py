"""GPU (B200): the reference-facing C++ layer — ark::GpuTsdfGenerator (include/tsdf.cuh), vhashing::HashTable
(include/vhashing.h), ark::PointCloudGenerator / ark::SaveFrame and the headless load_frames — driven the way the
reference's own callers drive them, checked against the CPU oracle."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from test_cpp_host import CPP, write_scannet_layout
from util import key_set

pytestmark = pytest.mark.gpu
BIN = os.path.join(CPP, "bin")


def need(name):
    path = os.path.join(BIN, name)
    if not os.access(path, os.X_OK):
        subprocess.check_call(["make", "-C", CPP], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return path


def read_tsdf_dump(path):
    raw = open(path, "rb").read()
    head = np.frombuffer(raw, np.float32, 8)
    n = int(head[0])
    rec = np.dtype([("key", np.int32, 3), ("sdf", np.float32, 512), ("w", np.float32, 512)])
    blocks = np.frombuffer(raw, rec, n, 32)
    return head, blocks


def read_ply(path):
    lines = open(path).read().split("\n")
    nv = int([l for l in lines if l.startswith("element vertex")][0].split()[-1])
    nf = int([l for l in lines if l.startswith("element face")][0].split()[-1])
    body = lines[lines.index("end_header") + 1:]
    verts = np.array([[float(x) for x in l.split()] for l in body[:nv]]).reshape(nv, 6)
    faces = np.array([[int(x) for x in l.split()[1:]] for l in body[nv:nv + nf]]).reshape(nf, 3)
    return verts, faces


def test_vhashing_drop_in():
    out = subprocess.run([need("test_vhashing")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "VHASHING_OK" in out.stdout, out.stdout + out.stderr


def test_gpu_tsdf_generator_drop_in(vh, ob, synth, tmp_path):
    """reference ctor + processFrame + SavePLY + SaveTSDF from C++, against the oracle on the same frames"""
    sc = synth.Scene(width=320, height=240, room=(5.0, 4.0, 2.6), n_frames=30, spheres=((3.9, 2.0, 1.0, 0.5),), color=True, holes=0.02)
    kw = dict(vox_size=0.02, trunc_margin=0.1, max_depth=3.5)
    n = 4
    frames = [sc.frame(i) for i in range(n)]
    path = tmp_path / "frames.bin"
    with open(path, "wb") as f:
        f.write(np.array([sc.width, sc.height, n], np.int32).tobytes())
        f.write(np.array([sc.fx, sc.fy, sc.cx, sc.cy, kw["max_depth"], kw["vox_size"], kw["trunc_margin"]], np.float32).tobytes())
        for d, rgb, c2w in frames:
            f.write(d.tobytes()); f.write(rgb.tobytes()); f.write(c2w.tobytes())
    out = subprocess.run([need("test_drop_in"), str(path), str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "DROP_IN_OK" in out.stdout, out.stdout + out.stderr
    o = ob.Oracle(ob.params_for_scene(sc, voxels_per_block=8, use_color=1, **kw))
    per_frame = re.findall(r"frame (\d+) visible (\d+) updates (\d+) triangles (\d+)", out.stdout)
    assert len(per_frame) == n
    for (d, rgb, c2w), rec in zip(frames, per_frame):
        o.process_frame(d, rgb, c2w)
        assert (int(rec[1]), int(rec[2]), int(rec[3])) == (o.num_visible, o.last_updates, o.last_triangles)
    head, blocks = read_tsdf_dump(tmp_path / "tsdf.bin")
    keys = o.all_keys()
    assert key_set(blocks["key"]) == key_set(keys) and np.isclose(head[6], kw["vox_size"]) and np.isclose(head[7], kw["trunc_margin"])
    sdf, w, _, _ = o.get_blocks(blocks["key"])
    assert np.array_equal(blocks["sdf"], sdf) and np.array_equal(blocks["w"], w)
    # PLY: exact-xyz vertex dedupe, vertices * vox_size, ASCII at ostream default precision (6 significant digits)
    xyz, trgb = o.triangles()
    verts, faces = read_ply(tmp_path / "model.ply")
    assert len(faces) == len(xyz)
    uniq = np.unique(xyz.reshape(-1, 3), axis=0)
    assert len(verts) == len(uniq)
    soup = xyz.reshape(-1, 3) * np.float32(kw["vox_size"])
    assert np.allclose(verts[faces.reshape(-1), :3], soup, rtol=1e-5, atol=1e-7)       # the PLY is ASCII at 6 significant digits
    m = re.search(r"vertices (\d+) faces (\d+)", out.stdout)
    assert (int(m.group(1)), int(m.group(2))) == (len(verts), len(faces))


def test_headless_load_frames(vh, ob, synth, tmp_path):
    """the reference's executable flow: SaveFrame::frameLoad -> PointCloudGenerator::PushFrame -> SavePly"""
    sc = synth.Scene(width=640, height=480, room=(4.0, 3.0, 2.5), n_frames=12, spheres=((2.8, 1.5, 1.0, 0.4),), color=True)
    folder = str(tmp_path) + "/"
    ids = [(j, 20 * j) for j in range(3)]                      # frame ids 0, 20, 40 as main.cpp:247-253 walks them
    write_scannet_layout(folder, synth, sc, ids)
    yaml = tmp_path / "settings.yaml"
    yaml.write_text("%YAML:1.0\n# synthetic\nCamera.fx: {fx}\nCamera.fy: {fy}\nCamera.cx: {cx}\nCamera.cy: {cy}\nCamera.width: 640\nCamera.height: 480\n"
                    "DepthMapFactor: 1000.0\nMaxDepth: 3.0\nVoxel.Origin.x: 0.5\nVoxel.Origin.y: 0.0\nVoxel.Origin.z: -0.5\nVoxel.Size: 0.04\n"
                    "Voxel.TruncMargin: 0.2\nVoxel.Dim.x: 8\nVoxel.Dim.y: 8\nVoxel.Dim.z: 8\n".format(fx=sc.fx, fy=sc.fy, cx=sc.cx, cy=sc.cy))
    ply, poses = tmp_path / "model.ply", tmp_path / "poses.txt"
    out = subprocess.run([need("load_frames"), folder, str(yaml), "--out", str(ply), "--poses", str(poses), "--buckets", "65536", "--pool", "65536"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "frames: 3" in out.stdout            # ids 60..140 are missing: the loop stops after 5 misses (main.cpp:249)
    used = np.loadtxt(poses).reshape(-1, 17)
    assert [int(r[0]) for r in used] == [0, 20, 40]
    import cv2
    o = ob.Oracle(ob.params_for_scene(sc, voxels_per_block=8, use_color=1, vox_size=0.04, trunc_margin=0.2, max_depth=3.0))
    for (j, fid), row in zip(ids, used):
        mm = cv2.imread(folder + f"depth/{fid}.png", -1)
        depth = (mm.astype(np.float32).astype(np.float64) * 0.001).astype(np.float32)
        rgb = sc.frame(j)[1]
        o.process_frame(depth, rgb, row[1:].astype(np.float32))     # the exact poses the facade handed to processFrame
    xyz, _ = o.triangles()
    verts, faces = read_ply(ply)
    assert len(faces) == len(xyz) > 0
    assert np.allclose(verts[faces.reshape(-1), :3], xyz.reshape(-1, 3) * np.float32(0.04), rtol=1e-5, atol=1e-7)       # the PLY is ASCII at 6 significant digits
    assert f"allocated blocks {len(o.all_keys())}" in out.stdout
