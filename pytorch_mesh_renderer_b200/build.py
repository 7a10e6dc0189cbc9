"""Compiles csrc/*.cu into pytorch_mesh_renderer_b200/libpmr_b200.so for sm_100a (nvcc
cross-compiles without a GPU).  The flags are the ones in csrc/Makefile; -fmad=false,
-prec-div=true and -ftz=false are part of the numerical contract."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpmr_b200.so")
SOURCES = ["c_api.cu", "raster_forward.cu", "raster_backward.cu", "vertex_stage.cu", "shade.cu", "mesh_normals.cu", "peer_exchange.cu"]
HEADERS = ["pmr_internal.cuh", "raster_math.cuh", "shade_math.cuh", "tensor_maps.cuh", os.path.join("..", "..", "include", "pmr_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-extended-lambda"]


def is_stale():
    if not os.path.exists(LIB):
        return True
    built = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > built for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """One nvcc per translation unit, side by side, then one link: same flags as csrc/Makefile."""
    if not force and not is_stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "nvcc")
    objdir = os.path.join(CSRC, "_obj")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd, cwd=CSRC)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        objects = list(pool.map(compile_one, SOURCES))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB, *objects, "-lcudart"]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
