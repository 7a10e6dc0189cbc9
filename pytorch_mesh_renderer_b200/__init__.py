"""B200-native implementation of pytorch_mesh_renderer's barycentric rasterization hot path.

Public surface = the reference's for this path (src/mesh_renderer/__init__.py:2,
rasterize.py, rasterize_triangles_ext.py): `rasterize`, `rasterize_clip_space`,
`rasterize_barycentric`, `BarycentricRasterizer`, plus `render` / `tone_mapper` (the package exports of
src/mesh_renderer/__init__.py:1) as a device-aware caller.  Rasterization and interpolation run in hand-written CUDA
(csrc/, built into libpmr_b200.so for sm_100a) behind the C ABI of include/pmr_b200.h.
"""
from .rasterize import rasterize, rasterize_barycentric, rasterize_clip_space
from .render import phong_shader, render, tone_mapper
from .rasterize_triangles_ext import (BarycentricRasterizer, RasterizeInterpolate, backward_mode,
                                      get_backward_mode, set_backward_mode)

__version__ = "0.1.0"
name = "pytorch_mesh_renderer_b200"
