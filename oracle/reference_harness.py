"""Runs the UNMODIFIED reference (kernel from oracle/_ref, Python layers imported from
/root/reference or from their verbatim staging in oracle/_ref/pysrc) so the oracle can be pinned against
it, golden vectors generated, and the reference arm of bench.py timed on the stock code path.

TEST INFRASTRUCTURE ONLY.  /root/reference exists in the build container, not on the GPU box; both the
prebuilt oracle/_ref/rasterize_triangles_cpp.so and oracle/_ref/pysrc/ (oracle/build_ref.py) travel there.
`available()` says whether the Python layers and the kernel can be imported.
"""
import os
import sys

def _reference_root():
    if os.path.isdir(os.path.join("/root/reference", "src", "mesh_renderer")):
        return "/root/reference"
    from . import build_ref
    return build_ref.stage_python() or "/root/reference"


REFERENCE_ROOT = _reference_root()
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


def import_reference(native_module):
    """The reference's src.mesh_renderer.{rasterize, rasterize_triangles_ext}, unmodified, with `native_module`
    standing where ext.py:3 does `import rasterize_triangles_cpp`, and USE_CPP_RASTERIZER on.  The modules are
    imported once; which native module the name resolves to is (re)bound on every call, so the same reference
    code can be run on the reference's own kernel and on this repository's drop-in.
    Returns (rasterize module, ext module)."""
    import importlib
    sys.dont_write_bytecode = True
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    sys.modules["rasterize_triangles_cpp"] = native_module
    importlib.import_module("src.mesh_renderer.rasterize")
    rast = sys.modules["src.mesh_renderer.rasterize"]       # the package attribute is a function
    rast.USE_CPP_RASTERIZER = True
    ext = importlib.import_module("src.mesh_renderer.rasterize_triangles_ext")
    ext.rasterize_triangles_cpp = native_module
    return rast, ext


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "mesh_renderer", "rasterize.py")) and kernel() is not None


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
        _rast_module, _ = import_reference(kernel())
    return _rast_module


def camera_utils():
    rasterize_module()
    return sys.modules["src.common.camera_utils"]
