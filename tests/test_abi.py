"""CPU: the C-ABI library builds for sm_100a, loads without a GPU, exports every symbol include/vh_c.h declares,
and the ctypes mirrors of its structs have the C sizes. No compute calls here."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def built(vh):
    vh.build()
    return vh


def header_symbols():
    src = open(os.path.join(ROOT, "include", "vh_c.h")).read()
    return sorted(set(re.findall(r"VH_API\s+[\w\s\*]+?\b(vh_\w+)\s*\(", src)))


def test_header_and_binding_agree(vh):
    assert header_symbols() == sorted(vh.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol(built):
    out = subprocess.check_output(["nm", "-D", "--defined-only", built.LIB_PATH], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    missing = [s for s in header_symbols() if s not in exported]
    assert not missing, missing
    extra = [s for s in exported if s.startswith("vh_") and s not in header_symbols()]
    assert not extra, f"undeclared exports: {extra}"
    lib = built.load_library()
    for s in header_symbols():
        assert getattr(lib, s) is not None


def test_library_is_sm100a_only(built):
    out = subprocess.check_output(["cuobjdump", "-lelf", built.LIB_PATH], text=True)
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_struct_sizes_match_c(built, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "vh_c.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(vh_params), sizeof(vh_stats),'
                   ' sizeof(vh_triangle), sizeof(vh_vertex), sizeof(vh_map_view));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    assert sizes[0] == C.sizeof(built.VhParams)
    assert sizes[1] == C.sizeof(built.VhStats)
    assert sizes[2] == built.TRI_DTYPE.itemsize == 48
    assert sizes[3] == built.VERT_DTYPE.itemsize == 16


def test_defaults_are_the_reference_constants(built):
    p = built.default_params()
    assert (p.blocks_per_chunk, p.dda_stride, p.max_ray_steps, p.max_chunk_num) == (8, 10, 100, 128)   # tsdf.cuh:41-43, tsdf.cu:13,2156
    assert abs(p.min_depth - 0.1) < 1e-7 and p.chunk_radius == 4.0 and p.entries_per_bucket == 4      # tsdf.cu:1318, tsdf.cuh:44, tsdf.cu:1488
    assert p.voxels_per_block == 8


def test_no_gpu_means_loud_failure(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(built.VhError) as ei:
        built.TsdfEngine(built.default_params())
    assert ei.value.code == 2 and "no CPU fallback" in str(ei.value)
