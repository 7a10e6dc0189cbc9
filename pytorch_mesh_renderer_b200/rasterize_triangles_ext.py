"""autograd glue for the CUDA rasterizer -- the counterpart of the reference's
src/mesh_renderer/rasterize_triangles_ext.py:6-63 (class and argument order kept verbatim).

`BarycentricRasterizer.apply(clip_space_vertices, triangles, image_width, image_height)` returns
`(px_triangle_ids, px_barycentric_coords, z_buffer)`; backward maps d(barycentrics) to
d(clip-space vertices) (x, y, w columns; z column zero) and returns
`(df_dvertices, zeros_like(triangles), None, None)` like the reference (ext.py:63).

Additions over the reference: a leading batch dimension is accepted ([B,V,4] -> [B,H,W,...]),
and the accumulation order of backward is selectable (`set_backward_mode`).
"""
import contextlib
import os

import torch

from . import ops

_backward_mode = os.environ.get("PMR_BACKWARD_MODE", "atomic")
_mode_explicit = "PMR_BACKWARD_MODE" in os.environ


def set_backward_mode(mode):
    """'atomic' (throughput; fp32 sums in arbitrary order) or 'ordered' (the reference's
    summation order, bit-reproducible; see include/pmr_b200.h)."""
    global _backward_mode, _mode_explicit
    ops.mode_code(mode)
    _backward_mode = mode
    _mode_explicit = True


def get_backward_mode():
    return _backward_mode


def mode_was_set_explicitly():
    """False while the mode is the built-in default (nobody called set_backward_mode, no PMR_BACKWARD_MODE)."""
    return _mode_explicit


@contextlib.contextmanager
def backward_mode(mode):
    global _mode_explicit
    previous, was_explicit = get_backward_mode(), _mode_explicit
    set_backward_mode(mode)
    try:
        yield
    finally:
        set_backward_mode(previous)
        _mode_explicit = was_explicit


class BarycentricRasterizer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, clip_space_vertices, triangles, image_width, image_height):
        """Rasterize clip-space (xyzw) vertices [V,4] (or [B,V,4]) with int32 triangles [T,3].

        Returns px_triangle_ids [H,W] int32 (0 also where empty), px_barycentric_coords [H,W,3]
        (zeros where empty) and z_buffer [H,W] (1.0 where empty); see
        rasterize_triangles.cpp:275-301 for the conventions.
        """
        batched = clip_space_vertices.dim() == 3
        v = clip_space_vertices if batched else clip_space_vertices.unsqueeze(0)
        ids, bary, z = ops.rasterize_forward(v, triangles, image_width, image_height)
        ctx.save_for_backward(v, triangles, ids, bary)
        ctx.batched = batched
        ctx.mode = get_backward_mode()
        ctx.mark_non_differentiable(ids)
        ctx.set_materialize_grads(False)      # no zero-filled [H,W] gradients for unused outputs
        if not batched:
            ids, bary, z = ids[0], bary[0], z[0]
        return ids, bary, z

    @staticmethod
    def backward(ctx, _, df_dbarycentric_coords, __):
        v, triangles, ids, bary = ctx.saved_tensors
        if df_dbarycentric_coords is None:
            return torch.zeros_like(v if ctx.batched else v[0]), torch.zeros_like(triangles), None, None
        g = df_dbarycentric_coords if ctx.batched else df_dbarycentric_coords.unsqueeze(0)
        df_dvertices = ops.rasterize_backward(g, v, triangles, ids, bary, ctx.mode)
        if not ctx.batched:
            df_dvertices = df_dvertices[0]
        return df_dvertices, torch.zeros_like(triangles), None, None


class RasterizeInterpolate(torch.autograd.Function):
    """rasterize_clip_space (rasterize.py:66-152) as ONE differentiable op: the rasterizer, the
    corner-attribute gather, the barycentric weighting, alpha and the background blend run in a
    single fused kernel pass forward and a single one backward, instead of the reference's Python
    loop over images plus eight torch ops (and their index_put_ backward)."""

    @staticmethod
    def forward(ctx, clip_space_vertices, attributes, triangles, image_width, image_height, background_value):
        image, ids, bary, z = ops.rasterize_interpolate_forward(
            clip_space_vertices, attributes, triangles, background_value, image_width, image_height)
        ctx.save_for_backward(clip_space_vertices, attributes, triangles, ids, bary)
        ctx.mode = get_backward_mode()
        ctx.mark_non_differentiable(ids, bary, z)     # the buffers are by-products here
        ctx.set_materialize_grads(False)
        return image, ids, bary, z

    @staticmethod
    def backward(ctx, grad_image, _ids, _bary, _z):
        v, a, triangles, ids, bary = ctx.saved_tensors
        need_v, need_a, _, _, _, need_bg = ctx.needs_input_grad
        dv = da = d_bg = None
        if grad_image is None:
            return None, None, None, None, None, None
        if need_v or need_a:
            dv, da = ops.rasterize_interpolate_backward(grad_image.contiguous(), v, a, triangles, ids, bary,
                                                        ctx.mode, need_vertices=need_v, need_attributes=need_a)
        if need_bg:
            # d out / d background = 1 - alpha  (rasterize.py:145-150); off the hot path.
            alpha = torch.clamp(torch.sum(2.0 * bary, dim=3, keepdim=True), 0.0, 1.0)
            d_bg = (grad_image * (1.0 - alpha)).sum(dim=(0, 1, 2))
        return dv, da, None, None, None, d_bg
