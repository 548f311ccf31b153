"""CPU: the C++ host side above the C ABI — the drop-in headers compile with the reference's spellings, and the
OpenCV-free frame I/O (include/SaveFrame.h, include/vh_image_io.h) reads and writes the reference's on-disk layout
(/root/reference/src/SaveFrame.cpp:120-218) identically to OpenCV (cv2 is the independent decoder here)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

CPP = os.path.join(ROOT, "tests", "cpp")


@pytest.fixture(scope="module")
def cpp_bins(vh):
    vh.build()
    subprocess.check_call(["make", "-C", CPP], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return os.path.join(CPP, "bin")


def write_scannet_layout(folder, synth, sc, ids, ppm=False):
    """frames of a synthetic scene in the layout SaveFrame::frameLoad expects (depth u16 mm PNG, RGB image, 4x4 pose text)."""
    import cv2
    for sub in ("RGB", "depth", "tcw"):
        os.makedirs(os.path.join(folder, sub), exist_ok=True)
    out = {}
    for j, fid in ids:
        d, rgb, c2w = sc.frame(j)
        mm = np.round(d.astype(np.float64) * 1000.0).astype(np.uint16)
        assert cv2.imwrite(os.path.join(folder, "depth", f"{fid}.png"), mm)
        if ppm:
            with open(os.path.join(folder, "RGB", f"{fid}.ppm"), "wb") as f:
                f.write(b"P6\n%d %d\n255\n" % (sc.width, sc.height)); f.write(rgb.tobytes())
        else:
            assert cv2.imwrite(os.path.join(folder, "RGB", f"{fid}.png"), rgb[..., ::-1])      # cv2 takes BGR; file holds RGB
        np.savetxt(os.path.join(folder, "tcw", f"{fid}.txt"), c2w.reshape(4, 4), fmt="%.9g")
        out[fid] = (mm, rgb, c2w)
    return out


def test_cpp_programs_build(cpp_bins):
    for name in ("test_drop_in", "test_vhashing", "load_frames", "test_frame_io"):
        assert os.access(os.path.join(cpp_bins, name), os.X_OK), name


@pytest.mark.parametrize("ppm", [False, True])
def test_frame_io_matches_opencv(cpp_bins, synth, tmp_path, ppm):
    import cv2
    sc = synth.Scene(width=640, height=480, room=(4.0, 3.0, 2.5), n_frames=10, color=True, holes=0.03)
    folder = str(tmp_path) + "/"
    frames = write_scannet_layout(folder, synth, sc, [(2, 0)], ppm=ppm)
    out = subprocess.check_output([os.path.join(cpp_bins, "test_frame_io"), folder], text=True)
    assert "FRAME_IO_OK" in out
    raw = open(folder + "dump.bin", "rb").read()
    dw, dh, cw, ch = np.frombuffer(raw, np.int32, 4)
    assert (dw, dh, cw, ch) == (640, 480, 640, 480)
    off = 16
    depth = np.frombuffer(raw, np.float32, dw * dh, off).reshape(dh, dw); off += 4 * dw * dh
    rgb = np.frombuffer(raw, np.uint8, cw * ch * 3, off).reshape(ch, cw, 3); off += cw * ch * 3
    c2w = np.frombuffer(raw, np.float32, 16, off)
    mm, rgb_in, c2w_in = frames[0]
    # convertTo(CV_32FC1) then *= 0.001 (SaveFrame.cpp:174-180)
    assert np.array_equal(depth, (mm.astype(np.float32).astype(np.float64) * 0.001).astype(np.float32))
    assert np.array_equal(rgb, rgb_in)
    assert np.allclose(c2w, c2w_in, atol=2e-6)          # inv(inv(pose)): upstream of the engine boundary, not bit-exact
    # frameWrite output is readable by OpenCV and holds the same frame
    back = cv2.imread(folder + "depth/1.png", -1)
    assert back.dtype == np.uint16 and np.array_equal(back, mm)
    assert np.array_equal(cv2.imread(folder + "RGB/1.png", cv2.IMREAD_COLOR)[..., ::-1], rgb_in)
    assert np.allclose(np.loadtxt(folder + "tcw/1.txt").reshape(16), c2w_in, atol=4e-6)
