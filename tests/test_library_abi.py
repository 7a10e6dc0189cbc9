"""CPU-side checks of the boundary: the shared library loads and exports every symbol that
include/pmr_b200.h declares, and the product package never touches oracle/."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pmr_b200.h")).read()
    return sorted(set(re.findall(r"PMR_API[^;(]*?\b(pmr_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for s in ("pmr_create", "pmr_destroy", "pmr_rasterize_forward", "pmr_rasterize_backward",
              "pmr_interpolate_forward", "pmr_rasterize_interpolate_forward",
              "pmr_rasterize_interpolate_backward", "pmr_rasterize_clip_space_host"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    from pytorch_mesh_renderer_b200 import _lib, build
    build.build()
    lib = _lib.load()
    for s in declared_symbols():
        assert hasattr(lib, s), s
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared_symbols()
    assert lib.pmr_version() >= 100


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import pytorch_mesh_renderer_b200 as m
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.rasterize_clip_space(torch.zeros(1, 3, 4), torch.zeros(1, 3, 2), torch.zeros(1, 3, dtype=torch.int32),
                               4, 4, torch.zeros(2))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pytorch_mesh_renderer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "raster_oracle" not in text or f in ("raster_math.cuh",), f


def test_argument_errors_keep_reference_types():
    import torch
    import pytorch_mesh_renderer_b200 as m
    v = torch.zeros(1, 3, 4); a = torch.zeros(1, 3, 2); t = torch.zeros(1, 3, dtype=torch.int32)
    with pytest.raises(ValueError, match="Image width must be > 0"):
        m.rasterize_clip_space(v, a, t, 0, 4, torch.zeros(2))
    with pytest.raises(ValueError, match="must be 3D"):
        m.rasterize_clip_space(v[0], a, t, 4, 4, torch.zeros(2))


def test_both_build_recipes_compile_every_source():
    """build.py (what __graft_entry__.build() runs) and csrc/Makefile list the same .cu files: all of them."""
    import glob
    from pytorch_mesh_renderer_b200 import build
    csrc = os.path.join(ROOT, "pytorch_mesh_renderer_b200", "csrc")
    on_disk = sorted(os.path.basename(p) for p in glob.glob(os.path.join(csrc, "*.cu")))
    assert sorted(build.SOURCES) == on_disk
    makefile = open(os.path.join(csrc, "Makefile")).read()
    listed = re.search(r"^SRCS\s*:=\s*(.*)$", makefile, re.M).group(1).split()
    assert sorted(listed) == on_disk
