"""Drop-in for the reference's native module `rasterize_triangles_cpp` (PYBIND11_MODULE at
src/mesh_renderer/kernels/rasterize_triangles.cpp:421-424; called from rasterize_triangles_ext.py:39,56):

    forward(vertices [V,4] f32, triangles [T,3] i32, image_width, image_height)
        -> [px_triangle_ids [H,W] i32, px_barycentric_coords [H,W,3] f32, z_buffer [H,W] f32]
    backward(df_dbarycentric_coords [H,W,3], vertices, triangles, px_triangle_ids, px_barycentric_coords)
        -> [df_dvertices [V,4] f32]                       (K.cpp:302-307, :131-137)

Same names, argument order (width before height, outputs [height, width]), list return values and
RuntimeError on a wrong scalar type as the reference's accessor<> gives (K.cpp:323-328).  The work runs in
libpmr_b200.so on the tensors' CUDA device; CPU tensors (what the reference's own tests pass) are processed on the
current CUDA device and the results returned on the CPU.  With

    sys.modules["rasterize_triangles_cpp"] = pytorch_mesh_renderer_b200.rasterize_triangles_cpp

in place before the import, the reference's UNMODIFIED rasterize_triangles_ext.py / rasterize.py run on this
library (tests/test_gpu_reference_dropin.py).  backward() sums in the reference's order (K.cpp:156-157, :232-269)
unless `pytorch_mesh_renderer_b200.set_backward_mode('atomic')` selected the throughput mode.
"""
import torch

from . import ops
from . import rasterize_triangles_ext as _ext

__all__ = ["forward", "backward"]

_default_mode = "ordered"       # the reference's kernel is sequential: its bits are the contract of this module


def _device_of(t):
    return t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device())


def _mode():
    mode = _ext.get_backward_mode()
    return mode if _ext.mode_was_set_explicitly() else _default_mode


def forward(vertices, triangles, image_width, image_height):
    if vertices.dim() != 2 or vertices.shape[1] != 4:
        raise RuntimeError("vertices must have shape [vertex_count, 4], got %s" % (tuple(vertices.shape),))
    dev = _device_of(vertices)
    ids, bary, z = ops.rasterize_forward(vertices.detach().to(dev).unsqueeze(0), triangles.to(dev),
                                         image_width, image_height)
    out = [ids[0], bary[0], z[0]]
    if not vertices.is_cuda:
        out = [t.cpu() for t in out]
    return out


def backward(df_dbarycentric_coords, vertices, triangles, px_triangle_ids, px_barycentric_coords):
    dev = _device_of(vertices)
    lift = lambda t: t.detach().to(dev).unsqueeze(0)
    dv = ops.rasterize_backward(lift(df_dbarycentric_coords), lift(vertices), triangles.to(dev),
                                lift(px_triangle_ids), lift(px_barycentric_coords), _mode())[0]
    return [dv if vertices.is_cuda else dv.cpu()]
