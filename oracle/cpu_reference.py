"""CPU arm of bench.py: the reference's own implementation of the path, timed on host cores.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/__init__.py).

What runs per image is the reference's STOCK code path: its own
`rasterize_clip_space` (src/mesh_renderer/rasterize.py:66-152, unmodified, USE_CPP_RASTERIZER on) over its own
`BarycentricRasterizer` (rasterize_triangles_ext.py) over its compiled C++ kernel
`rasterize_triangles_cpp.forward/backward` -- the kernel built from
/root/reference/src/mesh_renderer/kernels/rasterize_triangles.cpp and the Python layers staged verbatim under
oracle/_ref/ by oracle/build_ref.py (git-ignored; they travel to the GPU box with the snapshot).  Only when the
staged Python layers are missing is the op chain of rasterize.py:118-150 restated below around the reference
kernel (same ops, one image at a time); when oracle/_ref is absent altogether the plain-C oracle stands in
(kind "port").

The kernel is single-threaded and holds the GIL, so parallelism is over images: one worker
process per host core, torch.set_num_threads(1) each.
"""
import os
import time

import numpy as np

_state = {}


def _workload(name, size_override=None):
    from pytorch_mesh_renderer_b200 import synthetic as S
    if name == "c1":
        return S.cube_test_scene()
    if name == "c2":
        return S.sphere_views(159, 158, 64, size_override or 512)
    if name == "c3":
        return S.sphere_views(708, 707, 16, size_override or 1024)
    if name == "c4":
        return S.sphere_views(224, 223, 256, size_override or 512)
    if name == "c5":
        return S.occlusion_soup(32, size_override or 2048)
    raise ValueError(name)


def _init(workload_name, use_reference_kernel):
    import torch
    torch.set_num_threads(1)
    sc = _workload(workload_name)
    _state["sc"] = sc
    _state["torch"] = torch
    K = None
    if use_reference_kernel:
        from oracle import reference_harness as rh
        K = rh.kernel()
    _state["K"] = K
    _state["rast"] = None
    if K is not None:
        try:
            if rh.available():
                _state["rast"] = rh.rasterize_module()      # the reference's own rasterize.py on its own kernel
        except Exception:                                    # noqa: BLE001 -- fall back to the restated op chain
            _state["rast"] = None
    if K is not None:
        class KernelFn(torch.autograd.Function):
            @staticmethod
            def forward(ctx, vertices, triangles, width, height):
                ids, bary, z = K.forward(vertices, triangles, width, height)
                ctx.save_for_backward(vertices, triangles, ids, bary)
                ctx.mark_non_differentiable(ids, z)
                return ids, bary.detach(), z

            @staticmethod
            def backward(ctx, _a, g_bary, _b):
                vertices, triangles, ids, bary = ctx.saved_tensors
                return K.backward(g_bary.contiguous(), vertices, triangles, ids, bary)[0], None, None, None
        _state["fn"] = KernelFn
    _one_image(0)          # first call pays torch's lazy initialisation; keep it out of the timings
    return os.getpid()


def _one_image(b):
    """rasterize_clip_space forward + backward for image b of the workload; returns pixels done."""
    sc, torch = _state["sc"], _state["torch"]
    W, H = sc["width"], sc["height"]
    A = sc["attributes"].shape[2]
    b = b % sc["clip_vertices"].shape[0]
    g = np.random.default_rng(1000 + b).standard_normal((H, W, A), dtype=np.float32)
    if _state["K"] is None:
        from oracle import oracle
        oracle.rasterize_clip_space(sc["clip_vertices"][b:b + 1], sc["attributes"][b:b + 1], sc["triangles"],
                                    W, H, sc["background"], grad_out=g[None])
        return W * H
    v = torch.from_numpy(sc["clip_vertices"][b]).requires_grad_(True)
    a = torch.from_numpy(sc["attributes"][b]).requires_grad_(True)
    t = torch.from_numpy(sc["triangles"])
    bg = torch.from_numpy(sc["background"])
    if _state["rast"] is not None:
        out = _state["rast"].rasterize_clip_space(v[None], a[None], t, W, H, bg)
        out.backward(torch.from_numpy(g)[None])
        return W * H
    ids, bary, _ = _state["fn"].apply(v, t, W, H)
    # rasterize.py:118-150, one image
    corner_ids = torch.index_select(t, 0, ids.reshape(-1).long())
    corners = a[corner_ids.long()]                                   # [P,3,A]
    weights = bary.reshape(-1, 3)
    image = torch.sum(corners * weights.unsqueeze(2), dim=1)
    alpha = torch.clamp(torch.sum(2.0 * weights, dim=1), 0.0, 1.0).unsqueeze(1)
    out = alpha * image + (1.0 - alpha) * bg
    out.backward(torch.from_numpy(g).reshape(-1, A))
    return W * H


def _noop(_):
    return 0


class CpuReference:
    """Pool of worker processes, one per host core, each holding the workload."""

    def __init__(self, workload_name, cores=None):
        import multiprocessing as mp
        from oracle import build_ref
        self.cores = int(cores or os.cpu_count() or 1)
        so = build_ref.build()
        build_ref.stage_python()
        self.kind = "reference" if so and os.path.exists(so) else "port"
        ctx = mp.get_context("spawn")
        self.pool = ctx.Pool(self.cores, initializer=_init, initargs=(workload_name, self.kind == "reference"))
        self.pool.map(_noop, range(self.cores * 2))
        self.workload_name = workload_name

    def warm(self):
        self.pool.map(_one_image, range(self.cores), chunksize=1)

    def timed(self, n_images):
        """Runs n_images images across the pool; returns (pixels, seconds)."""
        t0 = time.perf_counter()
        px = sum(self.pool.map(_one_image, range(n_images), chunksize=1))
        return px, time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()
