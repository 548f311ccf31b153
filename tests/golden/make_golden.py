"""Generate tests/golden/*.npz from the REFERENCE ITSELF (its src/tsdf.cu run under the CPU emulation of
oracle/build_ref.sh). Run here, where /root/reference exists:

    oracle/build_ref.sh && python tests/golden/make_golden.py

Each fixture holds the synthetic-scene parameters (inputs are regenerated deterministically from them), the
visible block list of every frame, every stored block of the final map (keys, sdf, weight, rgb — only blocks that
were ever visible), the final ordered triangle soup and the bytes of the PLY file the reference's own SavePLY / tsdf2mesh
(src/tsdf.cu:1697-1708, :1760-1888) writes for that map. One subprocess per case: the reference keeps its tables in
file-scope globals (tsdf.cu:19-20), so one engine per process.
"""
import importlib
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from cases import CASES  # noqa: E402


def run_case(name):
    from oracle import binding as ob
    synth = importlib.import_module("voxel-hashing-sdf_b200.synth")
    case = CASES[name]
    sc = synth.Scene(**case["scene"])
    ref = ob.RefEmu(sc, case["vpb"], case["vox_size"], case["trunc"], case["max_depth"])
    vis, seen, streamed = [], {}, []
    for i in range(case["frames"]):
        d, rgb, c2w = sc.frame(i)
        ref.process_frame(d, rgb, c2w)
        k = ref.visible_keys()
        vis.append(k)
        streamed.append(ref.streamed_blocks)
        for key in map(tuple, k):
            seen[key] = True
    keys = np.array(sorted(seen.keys()), np.int32).reshape(-1, 3)
    sdf, w, rgb, found = ref.get_blocks(keys)
    assert found.all()
    xyz, trgb = ref.triangles()
    out = dict(keys=keys, sdf=sdf, weight=w, rgb=rgb, tri_xyz=xyz, tri_rgb=trgb, streamed=np.array(streamed, np.int64),
               n_frames=np.array(case["frames"]))
    for i, k in enumerate(vis):
        out[f"visible_{i}"] = k
    import tempfile
    with tempfile.TemporaryDirectory() as td:          # the reference's own mesh export: vertex welding, numbering, ASCII formatting
        ref.save_ply(os.path.join(td, "model.ply"))
        out["ply"] = np.frombuffer(open(os.path.join(td, "model.ply"), "rb").read(), np.uint8)
    cs = ref.checksum()
    out["checksum"] = np.array([cs["sum_sdf"], cs["sum_w"], cs["n_observed"], cs["n_negative"]], np.float64)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "blocks", len(keys), "triangles", len(xyz), "visible", [len(v) for v in vis], "->", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(sys.argv[1])
    else:
        for name in CASES:
            subprocess.check_call([sys.executable, os.path.abspath(__file__), name], stdout=None)
