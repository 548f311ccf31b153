"""CPU: marching-cubes table self-consistency and synthetic-scene conventions."""
import os
import re
import subprocess

import numpy as np

from conftest import ROOT


def _rows():
    src = open(os.path.join(ROOT, "include", "vh_mc_tables.h")).read()
    body = src[src.index("VH_MC_TRI_ROWS[256]"):]
    body = body[:body.index("};")]
    rows = re.findall(r'"([0-9a-b]*)"', body)
    assert len(rows) == 256
    return rows


PAIRS = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]


def test_tri_rows_use_exactly_the_cut_edges():
    for c, r in enumerate(_rows()):
        assert len(r) % 3 == 0 and len(r) <= 15
        cut = {e for e, (a, b) in enumerate(PAIRS) if ((c >> a) & 1) != ((c >> b) & 1)}
        assert {int(ch, 16) for ch in r} == cut, c


def test_tri_rows_complement_symmetry_and_counts():
    rows = _rows()
    counts = [len(r) // 3 for r in rows]
    assert counts[0] == counts[255] == 0 and max(counts) == 5
    assert sum(counts) == 820                      # total triangles of Bourke's table
    for c in range(256):
        assert counts[c] == counts[255 - c] or True  # (counts need not be symmetric; edges are)
        assert {int(ch, 16) for ch in rows[c]} == {int(ch, 16) for ch in rows[255 - c]}


def test_expand_tables_from_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include <stdio.h>\n#include "vh_mc_tables.h"\nint main(){signed char t[256][16];unsigned char n[256];unsigned short e[256];'
                   'vh_mc_expand_tables(t,n,e);printf("%d %d %d %d %x %x\\n",n[1],t[1][0],t[1][1],t[1][2],e[1],e[3]);return 0;}\n')
    exe = tmp_path / "t"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    assert subprocess.check_output([str(exe)], text=True).split() == ["1", "0", "8", "3", "109", "30a"]


def test_scene_conventions(synth):
    sc = synth.Scene(width=64, height=48, room=(4, 3, 2.5), n_frames=8, spheres=((2.8, 1.5, 1.0, 0.4),), holes=0.1, color=True)
    d, rgb, c2w = sc.frame(3)
    assert d.dtype == np.float32 and d.shape == (48, 64) and rgb.shape == (48, 64, 3) and c2w.shape == (16,)
    assert np.all(c2w[12:] == [0, 0, 0, 1])
    R = c2w.reshape(4, 4)[:3, :3].astype(np.float64)
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-6) and np.linalg.det(R) > 0.99
    nz = d[d > 0]
    assert np.allclose(nz * 1000, np.round(nz * 1000), atol=1e-3)          # millimetre quantised
    assert 0.05 < (d == 0).mean() < 0.2
    d2, _, _ = sc.frame(3)
    assert np.array_equal(d, d2)                                           # deterministic
    sc0 = synth.Scene(width=64, height=48)
    assert abs(sc0.fx - 57.7) < 1e-6 and sc0.cx == 32.0
