"""Entry points of the reference's src/mesh_renderer/rasterize.py, same names, argument order
and error behaviour, running on the CUDA kernels of libpmr_b200.so.

  rasterize_barycentric(clip_space_vertices, triangles, image_width, image_height)   rast.py:15-25
  rasterize(world_space_vertices, attributes, triangles, camera_matrices, W, H, bg)  rast.py:27-63
  rasterize_clip_space(clip_space_vertices, attributes, triangles, W, H, bg)         rast.py:66-152

There is no USE_CPP_RASTERIZER switch and no Python rasterizer: the CUDA path is the only path.
Tensors that live on the CPU (the reference's tests pass CPU tensors) are moved to the current
CUDA device, processed there and the results moved back; gradients flow through the copies.
"""
import torch

from . import camera_utils
from .rasterize_triangles_ext import BarycentricRasterizer, RasterizeInterpolate


def _compute_device(*tensors):
    for t in tensors:
        if isinstance(t, torch.Tensor) and t.is_cuda:
            return t.device
    if not torch.cuda.is_available():
        raise RuntimeError("pytorch_mesh_renderer_b200 needs a CUDA device; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def rasterize_barycentric(clip_space_vertices, triangles, image_width, image_height):
    """Unbatched [V,4] (the reference's form) or batched [B,V,4] clip-space vertices ->
    (px_triangle_ids, px_barycentric_coords, z_buffer)."""
    home = clip_space_vertices.device
    dev = _compute_device(clip_space_vertices, triangles)
    ids, bary, z = BarycentricRasterizer.apply(
        clip_space_vertices.to(dev), triangles.to(dev), image_width, image_height)
    if home != dev:
        ids, bary, z = ids.to(home), bary.to(home), z.to(home)
    return ids, bary, z


def rasterize(world_space_vertices, attributes, triangles, camera_matrices, image_width, image_height,
              background_value):
    """Applies the model-view-projection matrices [B,4,4] to world-space vertices [B,V,3] and calls
    rasterize_clip_space (rast.py:60-63)."""
    dev = _compute_device(world_space_vertices, attributes, camera_matrices)
    home = world_space_vertices.device
    clip_space_vertices = camera_utils.transform_homogeneous(
        camera_matrices.to(dev), world_space_vertices.to(dev))
    out = rasterize_clip_space(clip_space_vertices, attributes, triangles, image_width, image_height,
                               background_value)
    return out.to(home) if home != out.device else out


def rasterize_clip_space(clip_space_vertices, attributes, triangles, image_width, image_height,
                         background_value, return_buffers=False):
    """Rasterizes clip-space vertices [B,V,4] and interpolates attributes [B,V,A] perspective-correctly.

    Returns the attribute image [B,H,W,A]; pixels outside all triangles take background_value [A].
    With return_buffers=True also returns (px_triangle_ids, px_barycentric_coords, z_buffer).
    Raises ValueError for non-positive image sizes or a vertex buffer that is not 3-D
    (rast.py:98-103).
    """
    if not image_width > 0:
        raise ValueError("Image width must be > 0.")
    if not image_height > 0:
        raise ValueError("Image height must be > 0.")
    if len(clip_space_vertices.shape) != 3:
        raise ValueError("The vertex buffer must be 3D.")

    home = clip_space_vertices.device
    dev = _compute_device(clip_space_vertices, attributes)
    background = torch.as_tensor(background_value)
    background = background.to(device=dev, dtype=torch.float32)       # render.py:197 passes int64
    image, ids, bary, z = RasterizeInterpolate.apply(
        clip_space_vertices.to(dev), attributes.to(dev), triangles.to(dev),
        int(image_width), int(image_height), background)
    if home != dev:
        image = image.to(home)
    if return_buffers:
        if home != dev:
            ids, bary, z = ids.to(home), bary.to(home), z.to(home)
        return image, (ids, bary, z)
    return image
