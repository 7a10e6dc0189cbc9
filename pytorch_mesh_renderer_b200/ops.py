"""Tensor-level wrappers over the C ABI (include/pmr_b200.h): argument checks in the
reference's terms, output allocation on the inputs' CUDA device, launch on torch's current
stream.  No arithmetic happens in Python."""
import os

import torch

from . import _lib

# Range check of the triangle indices on every call (one min/max reduction and a device synchronisation):
# off by default like the reference's native layer (K.cpp:331-337 indexes straight into the accessor); the
# reference's Python layer raises IndexError for such input (rast.py:118-132), which this reproduces when on.
_check_indices = os.environ.get("PMR_CHECK_INDICES", "0") not in ("", "0")


def set_index_checks(on):
    """Turns the per-call range check of `triangles` against [0, vertex_count) on or off (debugging aid)."""
    global _check_indices
    _check_indices = bool(on)


def _check_mesh(v, t, a=None, last=4):
    """Shapes the kernels rely on; out-of-range sizes would make them gather out of bounds (a sticky CUDA fault)."""
    if v.dim() != 3 or v.shape[-1] != last:
        raise ValueError("vertices must have shape [batch_size, vertex_count, %d], got %s" % (last, tuple(v.shape)))
    if t.dim() != 2 or t.shape[1] != 3:
        raise ValueError("triangles must have shape [triangle_count, 3], got %s" % (tuple(t.shape),))
    if a is not None and (a.dim() != 3 or tuple(a.shape[:2]) != tuple(v.shape[:2])):
        raise ValueError("attributes must have shape [batch_size, vertex_count, attribute_count] matching the "
                         "vertices %s, got %s" % (tuple(v.shape[:2]), tuple(a.shape)))
    if _check_indices and t.numel():
        lo, hi = int(t.min()), int(t.max())
        if lo < 0 or hi >= v.shape[1]:
            raise IndexError("triangle vertex index out of range: [%d, %d] with %d vertices" % (lo, hi, v.shape[1]))


def _check_buffers(i, b, B):
    if i.dim() != 3 or i.shape[0] != B or b.dim() != 4 or tuple(b.shape) != tuple(i.shape) + (3,):
        raise ValueError("px_triangle_ids [B,H,W] / px_barycentric_coords [B,H,W,3] do not match the batch: %s, %s"
                         % (tuple(i.shape), tuple(b.shape)))


_MODE_NAMES = {"atomic": _lib.BACKWARD_ATOMIC, "ordered": _lib.BACKWARD_ORDERED}


def _require(t, dtype, what):
    # The reference's accessor<> raises RuntimeError for a wrong scalar type
    # (rasterize_triangles.cpp:323-328); keep type and wording.
    if t.dtype != dtype:
        names = {torch.float32: "Float", torch.int32: "Int"}
        raise RuntimeError("expected scalar type %s but found %s for %s"
                           % (names[dtype], str(t.dtype).replace("torch.", ""), what))
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor at this level" % what)
    return t.contiguous()


def _aligned(t):
    if t.data_ptr() % 16:
        t = t.clone(memory_format=torch.contiguous_format)
    return t


def mode_code(mode):
    if isinstance(mode, str):
        if mode not in _MODE_NAMES:
            raise ValueError("backward mode must be 'atomic' or 'ordered', got %r" % (mode,))
        return _MODE_NAMES[mode]
    return int(mode)


def rasterize_forward(vertices, triangles, image_width, image_height):
    """vertices [B,V,4] f32, triangles [T,3] i32 -> ids [B,H,W] i32, bary [B,H,W,3], z [B,H,W]."""
    v = _aligned(_require(vertices, torch.float32, "vertices"))
    t = _require(triangles, torch.int32, "triangles")
    _check_mesh(v, t)
    B, V, _ = v.shape
    W, H = int(image_width), int(image_height)
    dev = v.device
    ids = torch.empty((B, max(H, 0), max(W, 0)), dtype=torch.int32, device=dev)
    bary = torch.empty((B, max(H, 0), max(W, 0), 3), dtype=torch.float32, device=dev)
    z = torch.empty((B, max(H, 0), max(W, 0)), dtype=torch.float32, device=dev)
    ctx = _lib.context(dev.index)
    with torch.cuda.device(dev):
        rc = _lib.load().pmr_rasterize_forward(ctx, _lib.ptr(v), _lib.ptr(t), B, V, t.shape[0], W, H,
                                               _lib.ptr(ids), _lib.ptr(bary), _lib.ptr(z), _lib.stream_ptr(dev))
    _lib.check(ctx, rc)
    return ids, bary, z


def rasterize_backward(df_dbary, vertices, triangles, ids, bary, mode):
    """-> df_dvertices [B,V,4]."""
    v = _aligned(_require(vertices, torch.float32, "vertices"))
    t = _require(triangles, torch.int32, "triangles")
    g = _require(df_dbary, torch.float32, "df_dbarycentric_coords")
    i = _require(ids, torch.int32, "px_triangle_ids")
    b = _require(bary, torch.float32, "px_barycentric_coords")
    _check_mesh(v, t)
    B, V, _ = v.shape
    _check_buffers(i, b, B)
    if tuple(g.shape) != tuple(b.shape):
        raise ValueError("df_dbarycentric_coords must have the shape of the barycentric buffer %s" % (tuple(b.shape),))
    H, W = i.shape[1], i.shape[2]
    out = torch.empty((B, V, 4), dtype=torch.float32, device=v.device)
    ctx = _lib.context(v.device.index)
    with torch.cuda.device(v.device):
        rc = _lib.load().pmr_rasterize_backward(ctx, _lib.ptr(g), _lib.ptr(v), _lib.ptr(t), _lib.ptr(i), _lib.ptr(b),
                                                B, V, t.shape[0], W, H, _lib.ptr(out), mode_code(mode),
                                                _lib.stream_ptr(v.device))
    _lib.check(ctx, rc)
    return out


def interpolate_forward(attributes, triangles, ids, bary, background):
    a = _require(attributes, torch.float32, "attributes")
    t = _require(triangles, torch.int32, "triangles")
    i = _require(ids, torch.int32, "px_triangle_ids")
    b = _require(bary, torch.float32, "px_barycentric_coords")
    bg = _require(background, torch.float32, "background_value")
    _check_mesh(a, t, last=a.shape[-1] if a.dim() == 3 else 0)
    B, V, A = a.shape
    _check_buffers(i, b, B)
    if bg.numel() != A:
        raise ValueError("background_value must have %d entries" % A)
    H, W = i.shape[1], i.shape[2]
    out = torch.empty((B, H, W, A), dtype=torch.float32, device=a.device)
    ctx = _lib.context(a.device.index)
    with torch.cuda.device(a.device):
        rc = _lib.load().pmr_interpolate_forward(ctx, _lib.ptr(a), _lib.ptr(t), _lib.ptr(i), _lib.ptr(b), _lib.ptr(bg),
                                                 B, V, t.shape[0], A, W, H, _lib.ptr(out), _lib.stream_ptr(a.device))
    _lib.check(ctx, rc)
    return out


def rasterize_interpolate_forward(vertices, attributes, triangles, background, image_width, image_height):
    """Fused rasterize_clip_space forward -> image [B,H,W,A], ids, bary, z."""
    v = _aligned(_require(vertices, torch.float32, "clip_space_vertices"))
    a = _require(attributes, torch.float32, "attributes")
    t = _require(triangles, torch.int32, "triangles")
    bg = _require(background, torch.float32, "background_value")
    _check_mesh(v, t, a)
    B, V, _ = v.shape
    A = a.shape[2]
    if bg.numel() != A:
        raise ValueError("background_value must have %d entries" % A)
    W, H = int(image_width), int(image_height)
    dev = v.device
    ids = torch.empty((B, H, W), dtype=torch.int32, device=dev)
    bary = torch.empty((B, H, W, 3), dtype=torch.float32, device=dev)
    z = torch.empty((B, H, W), dtype=torch.float32, device=dev)
    image = torch.empty((B, H, W, A), dtype=torch.float32, device=dev)
    ctx = _lib.context(dev.index)
    with torch.cuda.device(dev):
        rc = _lib.load().pmr_rasterize_interpolate_forward(
            ctx, _lib.ptr(v), _lib.ptr(a), _lib.ptr(t), _lib.ptr(bg), B, V, t.shape[0], A, W, H,
            _lib.ptr(ids), _lib.ptr(bary), _lib.ptr(z), _lib.ptr(image), _lib.stream_ptr(dev))
    _lib.check(ctx, rc)
    return image, ids, bary, z


def rasterize_interpolate_backward(grad_image, vertices, attributes, triangles, ids, bary, mode,
                                   need_vertices=True, need_attributes=True):
    """-> (d_vertices [B,V,4] or None, d_attributes [B,V,A] or None)."""
    g = _require(grad_image, torch.float32, "grad_output")
    v = _aligned(_require(vertices, torch.float32, "clip_space_vertices"))
    a = _require(attributes, torch.float32, "attributes")
    t = _require(triangles, torch.int32, "triangles")
    i = _require(ids, torch.int32, "px_triangle_ids")
    b = _require(bary, torch.float32, "px_barycentric_coords")
    _check_mesh(v, t, a)
    B, V, A = a.shape
    _check_buffers(i, b, B)
    if tuple(g.shape) != tuple(i.shape) + (A,):
        raise ValueError("grad_output must have shape %s, got %s" % (tuple(i.shape) + (A,), tuple(g.shape)))
    H, W = i.shape[1], i.shape[2]
    dev = v.device
    if need_vertices and need_attributes:
        # one allocation: the library then clears both gradients with one memset
        both = torch.empty((B * V * (4 + A),), dtype=torch.float32, device=dev)
        dv, da = both[:B * V * 4].view(B, V, 4), both[B * V * 4:].view(B, V, A)
    else:
        dv = torch.empty((B, V, 4), dtype=torch.float32, device=dev) if need_vertices else None
        da = torch.empty((B, V, A), dtype=torch.float32, device=dev) if need_attributes else None
    ctx = _lib.context(dev.index)
    with torch.cuda.device(dev):
        rc = _lib.load().pmr_rasterize_interpolate_backward(
            ctx, _lib.ptr(g), _lib.ptr(v), _lib.ptr(a), _lib.ptr(t), _lib.ptr(i), _lib.ptr(b),
            B, V, t.shape[0], A, W, H, _lib.ptr(dv), _lib.ptr(da), mode_code(mode), _lib.stream_ptr(dev))
    _lib.check(ctx, rc)
    return dv, da


def transform_forward(matrices, vertices, shared):
    """clip[b][v] = M_b (x, y, z, 1): matrices [B,4,4], vertices [V,3] (shared) or [B,V,3] -> [B,V,4]."""
    m = _require(matrices, torch.float32, "matrices")
    w = _require(vertices, torch.float32, "vertices")
    B = m.shape[0]
    V = w.shape[0] if shared else w.shape[1]
    clip = torch.empty((B, V, 4), dtype=torch.float32, device=m.device)
    ctx = _lib.context(m.device.index)
    with torch.cuda.device(m.device):
        rc = _lib.load().pmr_transform_forward(ctx, _lib.ptr(m), _lib.ptr(w), B, V, int(bool(shared)), _lib.ptr(clip),
                                               _lib.stream_ptr(m.device))
    _lib.check(ctx, rc)
    return clip


def transform_backward(matrices, d_clip, shared):
    """Gradient of transform_forward with respect to the vertices: [V,3] (summed over views) or [B,V,3]."""
    m = _require(matrices, torch.float32, "matrices")
    g = _aligned(_require(d_clip, torch.float32, "d_clip_vertices"))
    B, V, _ = g.shape
    out = torch.empty((V, 3) if shared else (B, V, 3), dtype=torch.float32, device=m.device)
    ctx = _lib.context(m.device.index)
    with torch.cuda.device(m.device):
        rc = _lib.load().pmr_transform_backward(ctx, _lib.ptr(m), _lib.ptr(g), B, V, int(bool(shared)), _lib.ptr(out),
                                                _lib.stream_ptr(m.device))
    _lib.check(ctx, rc)
    return out


def vertex_incidence(triangles, vertex_count):
    """Topology table of the vertex-normal kernels: (offsets int32 [V+1], incidence int32 [3T])."""
    t = _require(triangles, torch.int32, "triangles")
    V, T = int(vertex_count), t.shape[0]
    offsets = torch.empty((V + 1,), dtype=torch.int32, device=t.device)
    incidence = torch.empty((3 * T,), dtype=torch.int32, device=t.device)
    ctx = _lib.context(t.device.index)
    with torch.cuda.device(t.device):
        rc = _lib.load().pmr_vertex_incidence(ctx, _lib.ptr(t), T, V, _lib.ptr(offsets), _lib.ptr(incidence),
                                              _lib.stream_ptr(t.device))
    _lib.check(ctx, rc)
    return offsets, incidence


def vertex_normals_forward(vertices, triangles, offsets, incidence):
    """vertices [B,V,3] -> (normals [B,V,3], raw un-normalised sums [B,V,3])."""
    v = _require(vertices, torch.float32, "vertices")
    t = _require(triangles, torch.int32, "triangles")
    B, V, _ = v.shape
    normals, raw = torch.empty_like(v), torch.empty_like(v)
    ctx = _lib.context(v.device.index)
    with torch.cuda.device(v.device):
        rc = _lib.load().pmr_vertex_normals_forward(ctx, _lib.ptr(v), _lib.ptr(t), _lib.ptr(offsets),
                                                    _lib.ptr(incidence), B, V, t.shape[0], _lib.ptr(raw),
                                                    _lib.ptr(normals), _lib.stream_ptr(v.device))
    _lib.check(ctx, rc)
    return normals, raw


def vertex_normals_backward(grad_normals, raw, vertices, triangles, offsets, incidence):
    """-> d_vertices [B,V,3]."""
    g = _require(grad_normals, torch.float32, "grad_output")
    v = _require(vertices, torch.float32, "vertices")
    t = _require(triangles, torch.int32, "triangles")
    B, V, _ = v.shape
    grad_raw, d_vertices = torch.empty_like(v), torch.empty_like(v)
    ctx = _lib.context(v.device.index)
    with torch.cuda.device(v.device):
        rc = _lib.load().pmr_vertex_normals_backward(ctx, _lib.ptr(g), _lib.ptr(raw), _lib.ptr(v), _lib.ptr(t),
                                                     _lib.ptr(offsets), _lib.ptr(incidence), B, V, t.shape[0],
                                                     _lib.ptr(grad_raw), _lib.ptr(d_vertices),
                                                     _lib.stream_ptr(v.device))
    _lib.check(ctx, rc)
    return d_vertices


def shade_diffuse_forward(pixels, light_positions, light_intensities, ambient):
    """pixels [B,H,W,A>=9] -> RGBA [B,H,W,4] (rows flipped), diffuse + ambient Phong terms."""
    px = _require(pixels, torch.float32, "pixels")
    lp = _require(light_positions, torch.float32, "light_positions")
    li = _require(light_intensities, torch.float32, "light_intensities")
    am = _require(ambient, torch.float32, "ambient_color") if ambient is not None else None
    B, H, W, A = px.shape
    rgba = torch.empty((B, H, W, 4), dtype=torch.float32, device=px.device)
    ctx = _lib.context(px.device.index)
    with torch.cuda.device(px.device):
        rc = _lib.load().pmr_shade_diffuse_forward(ctx, _lib.ptr(px), _lib.ptr(lp), _lib.ptr(li), _lib.ptr(am), B,
                                                   lp.shape[1], A, W, H, _lib.ptr(rgba), _lib.stream_ptr(px.device))
    _lib.check(ctx, rc)
    return rgba


def shade_diffuse_backward(grad_rgba, pixels, light_positions, light_intensities, ambient):
    """-> d_pixels [B,H,W,A]."""
    g = _aligned(_require(grad_rgba, torch.float32, "grad_output"))
    px = _require(pixels, torch.float32, "pixels")
    lp = _require(light_positions, torch.float32, "light_positions")
    li = _require(light_intensities, torch.float32, "light_intensities")
    am = _require(ambient, torch.float32, "ambient_color") if ambient is not None else None
    B, H, W, A = px.shape
    d_pixels = torch.empty_like(px)
    ctx = _lib.context(px.device.index)
    with torch.cuda.device(px.device):
        rc = _lib.load().pmr_shade_diffuse_backward(ctx, _lib.ptr(g), _lib.ptr(px), _lib.ptr(lp), _lib.ptr(li),
                                                    _lib.ptr(am), B, lp.shape[1], A, W, H, _lib.ptr(d_pixels),
                                                    _lib.stream_ptr(px.device))
    _lib.check(ctx, rc)
    return d_pixels


def shade_phong_forward(pixels, light_positions, light_intensities, ambient, camera_position, shininess):
    """pixels [B,H,W,12|13] -> (RGBA [B,H,W,4] rows flipped, norm2 [B,L]); diffuse + ambient + specular."""
    px = _require(pixels, torch.float32, "pixels")
    lp = _require(light_positions, torch.float32, "light_positions")
    li = _require(light_intensities, torch.float32, "light_intensities")
    am = _require(ambient, torch.float32, "ambient_color") if ambient is not None else None
    cam = _require(camera_position, torch.float32, "camera_position")
    sh = _require(shininess, torch.float32, "shininess_coefficients") if shininess is not None else None
    B, H, W, A = px.shape
    L = lp.shape[1]
    rgba = torch.empty((B, H, W, 4), dtype=torch.float32, device=px.device)
    norm2 = torch.empty((B, L), dtype=torch.float32, device=px.device)
    ctx = _lib.context(px.device.index)
    with torch.cuda.device(px.device):
        rc = _lib.load().pmr_shade_phong_forward(ctx, _lib.ptr(px), _lib.ptr(lp), _lib.ptr(li), _lib.ptr(am), _lib.ptr(cam),
                                                 _lib.ptr(sh), B, L, A, W, H, _lib.ptr(norm2), _lib.ptr(rgba),
                                                 _lib.stream_ptr(px.device))
    _lib.check(ctx, rc)
    return rgba, norm2


def shade_phong_backward(grad_rgba, pixels, light_positions, light_intensities, ambient, camera_position, shininess,
                         norm2):
    """-> d_pixels [B,H,W,A]."""
    g = _aligned(_require(grad_rgba, torch.float32, "grad_output"))
    px = _require(pixels, torch.float32, "pixels")
    lp = _require(light_positions, torch.float32, "light_positions")
    li = _require(light_intensities, torch.float32, "light_intensities")
    am = _require(ambient, torch.float32, "ambient_color") if ambient is not None else None
    cam = _require(camera_position, torch.float32, "camera_position")
    sh = _require(shininess, torch.float32, "shininess_coefficients") if shininess is not None else None
    n2 = _require(norm2, torch.float32, "norm2")
    B, H, W, A = px.shape
    L = lp.shape[1]
    d_pixels = torch.empty_like(px)
    sums = torch.empty((B, L), dtype=torch.float32, device=px.device)
    ctx = _lib.context(px.device.index)
    with torch.cuda.device(px.device):
        rc = _lib.load().pmr_shade_phong_backward(ctx, _lib.ptr(g), _lib.ptr(px), _lib.ptr(lp), _lib.ptr(li), _lib.ptr(am),
                                                  _lib.ptr(cam), _lib.ptr(sh), _lib.ptr(n2), B, L, A, W, H,
                                                  _lib.ptr(sums), _lib.ptr(d_pixels), _lib.stream_ptr(px.device))
    _lib.check(ctx, rc)
    return d_pixels


def render_diffuse_forward(vertices, attributes, triangles, background, light_positions, light_intensities, ambient,
                           image_width, image_height):
    """Fused rasterize + interpolate + diffuse/ambient lighting -> (RGBA [B,H,W,4] rows flipped, ids, bary, z)."""
    v = _aligned(_require(vertices, torch.float32, "clip_space_vertices"))
    a = _require(attributes, torch.float32, "attributes")
    t = _require(triangles, torch.int32, "triangles")
    bg = _require(background, torch.float32, "background_value")
    lp = _require(light_positions, torch.float32, "light_positions")
    li = _require(light_intensities, torch.float32, "light_intensities")
    am = _require(ambient, torch.float32, "ambient_color") if ambient is not None else None
    B, V, _ = v.shape
    if a.shape[2] != 9:
        raise ValueError("the fused render path takes 9 attribute channels [normal, position, diffuse colour]")
    W, H = int(image_width), int(image_height)
    dev = v.device
    ids = torch.empty((B, H, W), dtype=torch.int32, device=dev)
    bary = torch.empty((B, H, W, 3), dtype=torch.float32, device=dev)
    z = torch.empty((B, H, W), dtype=torch.float32, device=dev)
    rgba = torch.empty((B, H, W, 4), dtype=torch.float32, device=dev)
    ctx = _lib.context(dev.index)
    with torch.cuda.device(dev):
        rc = _lib.load().pmr_render_diffuse_forward(
            ctx, _lib.ptr(v), _lib.ptr(a), _lib.ptr(t), _lib.ptr(bg), _lib.ptr(lp), _lib.ptr(li), _lib.ptr(am),
            B, V, t.shape[0], lp.shape[1], W, H, _lib.ptr(ids), _lib.ptr(bary), _lib.ptr(z), _lib.ptr(rgba),
            _lib.stream_ptr(dev))
    _lib.check(ctx, rc)
    return rgba, ids, bary, z


def render_diffuse_backward(grad_rgba, vertices, attributes, triangles, background, light_positions,
                            light_intensities, ambient, ids, bary, need_vertices=True, need_attributes=True):
    """-> (d_vertices [B,V,4] or None, d_attributes [B,V,9] or None)."""
    g = _aligned(_require(grad_rgba, torch.float32, "grad_output"))
    v = _aligned(_require(vertices, torch.float32, "clip_space_vertices"))
    a = _require(attributes, torch.float32, "attributes")
    t = _require(triangles, torch.int32, "triangles")
    bg = _require(background, torch.float32, "background_value")
    lp = _require(light_positions, torch.float32, "light_positions")
    li = _require(light_intensities, torch.float32, "light_intensities")
    am = _require(ambient, torch.float32, "ambient_color") if ambient is not None else None
    i = _require(ids, torch.int32, "px_triangle_ids")
    b = _require(bary, torch.float32, "px_barycentric_coords")
    B, V, _ = v.shape
    H, W = i.shape[1], i.shape[2]
    dev = v.device
    dv = torch.empty((B, V, 4), dtype=torch.float32, device=dev) if need_vertices else None
    da = torch.empty((B, V, 9), dtype=torch.float32, device=dev) if need_attributes else None
    ctx = _lib.context(dev.index)
    with torch.cuda.device(dev):
        rc = _lib.load().pmr_render_diffuse_backward(
            ctx, _lib.ptr(g), _lib.ptr(v), _lib.ptr(a), _lib.ptr(t), _lib.ptr(bg), _lib.ptr(lp), _lib.ptr(li),
            _lib.ptr(am), _lib.ptr(i), _lib.ptr(b), B, V, t.shape[0], lp.shape[1], W, H, _lib.ptr(dv), _lib.ptr(da),
            _lib.stream_ptr(dev))
    _lib.check(ctx, rc)
    return dv, da
