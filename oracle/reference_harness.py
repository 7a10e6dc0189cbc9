"""Runs the UNMODIFIED reference (kernel from oracle/_ref, Python layers imported from
/root/reference) so the oracle can be pinned against it and golden vectors generated.

TEST INFRASTRUCTURE ONLY.  /root/reference exists in the build container, not on the GPU box:
`available()` says whether the Python layers can be imported; `kernel()` only needs the
prebuilt oracle/_ref/rasterize_triangles_cpp.so (which does travel).
"""
import os
import sys

REFERENCE_ROOT = "/root/reference"
_kernel = None
_rast_module = None


def kernel():
    """The reference's compiled `rasterize_triangles_cpp` module, or None."""
    global _kernel
    if _kernel is None:
        from . import build_ref
        _kernel = build_ref.load()
        if _kernel is not None:
            sys.modules.setdefault("rasterize_triangles_cpp", _kernel)
    return _kernel


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "mesh_renderer")) and kernel() is not None


def rasterize_module():
    """The reference module src.mesh_renderer.rasterize with USE_CPP_RASTERIZER switched on.

    The package attribute of that name is shadowed by a function
    (src/mesh_renderer/__init__.py:2), so the module is fetched through sys.modules
    (SURVEY.md F9).
    """
    global _rast_module
    if _rast_module is None:
        assert available(), "reference sources or compiled kernel missing"
        sys.dont_write_bytecode = True
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        import importlib
        importlib.import_module("src.mesh_renderer.rasterize")
        mod = sys.modules["src.mesh_renderer.rasterize"]
        mod.USE_CPP_RASTERIZER = True
        _rast_module = mod
    return _rast_module


def camera_utils():
    rasterize_module()
    return sys.modules["src.common.camera_utils"]
