"""Mesh export pinned against the REFERENCE'S OWN SavePLY / tsdf2mesh (src/tsdf.cu:1697-1708, :1760-1888): the golden fixtures
hold the bytes of the PLY file the emulated reference wrote for its final map (tests/golden/make_golden.py).
  CPU: the rules the other tests restate (vertex ids in order of first appearance over the ordered soup, exact-xyz welding, first
       colour wins, positions scaled by the voxel size, default ostream formatting) reproduce that file byte for byte; and, where
       oracle/_ref exists, the emulated reference run live writes the same bytes as the fixture.
  GPU: vh_save_ply (GpuTsdfGenerator::SavePLY drop-in: ordered gather + GPU weld + ASCII writer) writes the same bytes, and
       vh_weld_mesh returns the reference's vertex numbering."""
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from conftest import ROOT
from util import CASES, engine_params, load_golden

NAMES = ["g8_color_holes", "g8_negative_coords", "g5_reference_defaults"]


def fmt(v):
    """C++ `ostream << float` with default precision: %g"""
    return "%g" % float(v)


def ply_from_soup(xyz, rgb, vox_size):
    """tsdf2mesh restated: weld on exact xyz (first appearance, first colour), scale, write (tsdf.cu:1810-1821, :1866-1886)"""
    pts = xyz.reshape(-1, 3) + np.float32(0.0)                      # -0 and +0 are the same vertex (float ==)
    cols = rgb.reshape(-1, 3)
    uniq, first, inv = np.unique(pts, axis=0, return_index=True, return_inverse=True)
    rank = np.empty(len(uniq), np.int64); rank[np.argsort(first, kind="stable")] = np.arange(len(uniq))
    faces = rank[inv.reshape(-1)].reshape(-1, 3)
    order = np.sort(first)
    verts = pts[order] * np.float32(vox_size)
    vcol = cols[order]
    out = ["ply", "format ascii 1.0", "comment stanford bunny", f"element vertex {len(verts)}", "property float x", "property float y", "property float z",
           "property uchar red", "property uchar green", "property uchar blue", f"element face {len(faces)}", "property list uchar int vertex_index", "end_header"]
    out += [f"{fmt(p[0])} {fmt(p[1])} {fmt(p[2])} {int(c[0])} {int(c[1])} {int(c[2])}" for p, c in zip(verts, vcol)]
    out += [f"3 {f[0]} {f[1]} {f[2]}" for f in faces]
    return ("\n".join(out) + "\n").encode(), verts, vcol, faces


@pytest.mark.parametrize("name", NAMES)
def test_restated_mesh_assembly_reproduces_the_reference_ply(name):
    g = load_golden(name)
    ref = bytes(g["ply"])
    mine, verts, _, faces = ply_from_soup(g["tri_xyz"], g["tri_rgb"], CASES[name]["vox_size"])
    assert len(faces) == len(g["tri_xyz"]) and f"element face {len(faces)}".encode() in ref
    if mine != ref:
        a, b = mine.split(b"\n"), ref.split(b"\n")
        bad = [i for i, (x, y) in enumerate(zip(a, b)) if x != y]
        raise AssertionError(f"{name}: {len(bad)} of {len(b)} lines differ, first: line {bad[0]}: {a[bad[0]]!r} vs reference {b[bad[0]]!r}" if bad else f"{name}: length {len(a)} vs {len(b)} lines")


LIVE = """
import importlib, sys, os, tempfile, numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + '/tests')
from oracle import binding as ob
from util import CASES, load_golden
synth = importlib.import_module('voxel-hashing-sdf_b200.synth')
case = CASES[{name!r}]
sc = synth.Scene(**case['scene'])
ref = ob.RefEmu(sc, case['vpb'], case['vox_size'], case['trunc'], case['max_depth'])
for i in range(case['frames']):
    ref.process_frame(*sc.frame(i))
with tempfile.TemporaryDirectory() as td:
    ref.save_ply(os.path.join(td, 'm.ply'))
    live = open(os.path.join(td, 'm.ply'), 'rb').read()
assert live == bytes(load_golden({name!r})['ply']), 'the reference run live does not write the committed PLY'
print('PLY_OK', len(live))
"""


@pytest.mark.parametrize("name", ["g8_color_holes", "g5_reference_defaults"])
def test_fixture_ply_is_what_the_reference_writes(ob, name):
    if not ob.ref_emu_available(CASES[name]["vpb"]):
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    out = subprocess.run([sys.executable, "-c", textwrap.dedent(LIVE.format(root=ROOT, name=name))], capture_output=True, text=True, timeout=900)
    assert "PLY_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["g8_color_holes", "g8_negative_coords"])
def test_engine_save_ply_writes_the_reference_file(name, vh, synth, tmp_path):
    case, g = CASES[name], load_golden(name)
    sc = synth.Scene(**case["scene"])
    with vh.TsdfEngine(engine_params(vh, sc, case)) as eng:
        for i in range(case["frames"]):
            eng.processFrame(*sc.frame(i))
        ply = tmp_path / "model.ply"
        eng.SavePLY(str(ply))
        mine, ref = ply.read_bytes(), bytes(g["ply"])
        if mine != ref:
            a, b = mine.split(b"\n"), ref.split(b"\n")
            bad = [i for i, (x, y) in enumerate(zip(a, b)) if x != y]
            raise AssertionError(f"{name}: {len(a)} vs {len(b)} lines, first difference: line {bad[0] if bad else -1}: {a[bad[0]] if bad else b''!r} vs reference {b[bad[0]] if bad else b''!r}")
        # the welded mesh through the ABI: the reference's numbering and vertices
        _, verts_ref, vcol_ref, faces_ref = ply_from_soup(g["tri_xyz"], g["tri_rgb"], case["vox_size"])
        verts, faces = eng.weld()
        assert np.array_equal(faces.reshape(-1, 3), faces_ref)
        assert np.array_equal(verts["xyz"], verts_ref) and np.array_equal(verts["rgb"][:, :3], vcol_ref)
