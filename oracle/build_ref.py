"""Build the reference's own C++ rasterizer kernel into oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

The source is compiled where it lies (/root/reference/src/mesh_renderer/kernels/
rasterize_triangles.cpp, the file behind `rasterize_triangles_cpp.forward/backward`,
PYBIND11_MODULE at :421-424); nothing is copied into this repository. The output
`oracle/_ref/rasterize_triangles_cpp.so` is git-ignored but travels to the GPU box with the
gpurun snapshot. It is used (a) to pin the plain-C restatement in oracle/raster_oracle.c,
(b) to generate tests/golden/*.npz, and (c) as the `cpu_baseline` / `--impl reference` arm of
bench.py. The product path never loads it.

Flags: plain `-O3`, x86-64 baseline (no -march/-mfma/-ffast-math), exactly what
torch.utils.cpp_extension would use for the reference's setup.py, so that GCC cannot contract
a*b+c into an FMA.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src/mesh_renderer/kernels/rasterize_triangles.cpp"
OUT_DIR = os.path.join(HERE, "_ref")
OUT_SO = os.path.join(OUT_DIR, "rasterize_triangles_cpp.so")


def ref_so_path():
    return OUT_SO


def build(force=False, verbose=False):
    """Returns the .so path, or None when the reference source is not present (GPU box)."""
    if os.path.exists(OUT_SO) and not force:
        if not os.path.exists(REF_SRC) or os.path.getmtime(OUT_SO) >= os.path.getmtime(REF_SRC):
            return OUT_SO
    if not os.path.exists(REF_SRC):
        return OUT_SO if os.path.exists(OUT_SO) else None
    import torch
    from torch.utils import cpp_extension
    os.makedirs(OUT_DIR, exist_ok=True)
    inc = []
    for p in cpp_extension.include_paths():
        inc += ["-isystem", p]
    inc += ["-isystem", sysconfig.get_paths()["include"]]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    cmd = ["g++", "-O3", "-std=c++17", "-fPIC", "-shared", "-w",
           "-DTORCH_EXTENSION_NAME=rasterize_triangles_cpp",
           "-DTORCH_API_INCLUDE_EXTENSION_H",
           "-D_GLIBCXX_USE_CXX11_ABI=%d" % abi,
           *inc, REF_SRC, "-o", OUT_SO,
           "-L" + libdir, "-Wl,-rpath," + libdir,
           "-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return OUT_SO


# The reference's Python layers of the path (and what they import), staged beside the kernel so that the
# reference arm of bench.py and the drop-in tests run the STOCK code on the GPU box, where /root/reference does
# not exist.  Like the .so they are git-ignored: nothing of the reference enters this repository's history.
REF_ROOT = "/root/reference"
PY_DIR = os.path.join(OUT_DIR, "pysrc")
PY_FILES = ["src/__init__.py", "src/common/__init__.py", "src/common/camera_utils.py", "src/common/meshes.py",
            "src/common/obj_utils.py", "src/common/shapes.py", "src/common/debug_utils.py",
            "src/mesh_renderer/__init__.py", "src/mesh_renderer/rasterize.py",
            "src/mesh_renderer/rasterize_triangles_ext.py", "src/mesh_renderer/rasterize_triangles_python.py",
            "src/mesh_renderer/render.py"]


def stage_python(force=False):
    """Copies the files above to oracle/_ref/pysrc/ (verbatim).  Returns the directory to put on sys.path, or
    None when neither the reference nor an earlier staging is present."""
    import shutil
    if os.path.isdir(os.path.join(REF_ROOT, "src", "mesh_renderer")):
        for rel in PY_FILES:
            src, dst = os.path.join(REF_ROOT, rel), os.path.join(PY_DIR, rel)
            if not os.path.exists(src):
                continue
            if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(src, dst)
    return PY_DIR if os.path.exists(os.path.join(PY_DIR, "src", "mesh_renderer", "rasterize.py")) else None


def load():
    """Import the built module (torch must be imported first so libtorch is resolvable)."""
    import importlib.util
    import torch  # noqa: F401
    path = build()
    if path is None or not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("rasterize_triangles_cpp", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True)
    print("reference kernel:", p)
    print("reference python layers:", stage_python(force="--force" in sys.argv))
