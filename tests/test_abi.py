"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every declared symbol."""
import ctypes as C
import os
import re

import pytest

from moonshine_b200 import hydra

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    return set(re.findall(r"MSNE_API\s+[^;(]*?\b(HdMoonshine\w+)\s*\(", txt))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(hydra.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = C.CDLL(hydra.LIB_PATH)
    declared = _declared("moonshine.h") | _declared("moonshine_ext.h")
    assert len(_declared("moonshine.h")) == 24  # hydra/moonshine.h:72-95
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/*.h but not exported"
    assert declared == set(hydra.SIGNATURES), "python binding table and headers disagree"


def test_struct_layouts_match_reference_header():
    # hydra/moonshine.h:17-69 — sizes the reference's C++ delegate is compiled against
    assert C.sizeof(hydra.F32x3) == 12 and C.sizeof(hydra.F32x4) == 16 and C.sizeof(hydra.Mat3x4) == 48
    assert C.sizeof(hydra.Geometry) == 12 and hydra.Geometry.sampled.offset == 8
    assert C.sizeof(hydra.Lens) == 48 and C.sizeof(hydra.Material) == 24 and C.sizeof(hydra.Extent2D) == 8
    assert C.sizeof(hydra.MsneSpecConstants) == 28  # pipeline.zig:319-327


def test_stats_struct_matches_the_header_field_for_field():
    """MsneStats grows with the library: the ctypes mirror must list the header's fields in the header's order and width."""
    txt = open(os.path.join(ROOT, "include", "moonshine_ext.h")).read()
    body = re.search(r"typedef struct MsneStats \{(.*?)\} MsneStats;", txt, re.S).group(1)
    declared = re.findall(r"^\s*(uint64_t|double)\s+(\w+);", body, re.M)
    ctype = {"uint64_t": C.c_uint64, "double": C.c_double}
    assert [(n, ctype[t]) for t, n in declared] == list(hydra.MsneStats._fields_)
    assert C.sizeof(hydra.MsneStats) == 8 * len(declared)


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU"):
        hydra.HdMoonshine()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "moonshine_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h", ".sh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                assert "liboracle" not in src and "oracle/" not in src.replace("oracle/bvh.h)", "").replace("oracle/binding.py", ""), f
