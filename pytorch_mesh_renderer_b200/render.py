"""Direct caller of the hot path: `render` / `phong_shader` / `tone_mapper` with the signatures of the
reference's src/mesh_renderer/render.py (render :16-228, phong_shader :231-386, tone_mapper :389-419),
device-aware (everything stays on the device of `vertices`; the reference allocates on the CPU, SURVEY
F11).  Rasterization and attribute interpolation run in the CUDA kernels of libpmr_b200.  The lighting of
a `render` call is one kernel each way without specular colours (`shade_diffuse`: the reference's own cube test
and BASELINE config c1) and two each way with them (`shade_phong`: the per-(image, light) normalisation over all
pixels of render.py:347-353 is a reduction between the passes) -- csrc/shade.cu; callers that need gradients
with respect to the lights, the camera or a per-image shininess use the torch-op `phong_shader` below.  Everything is checked against outputs
of the unmodified reference (tests/golden/render_*.npz) and the reference's PNG fixtures.
"""
import torch
import torch.nn.functional as F

from . import camera_utils, ops
from .rasterize import rasterize


class _ShadeDiffuse(torch.autograd.Function):
    """normalize(normals) + diffuse/ambient Phong + mask + flip of render.py:201-228 / :231-386 as one
    kernel each way; differentiable with respect to the pixel attributes."""

    @staticmethod
    def forward(ctx, pixels, light_positions, light_intensities, ambient_color):
        rgba = ops.shade_diffuse_forward(pixels, light_positions, light_intensities, ambient_color)
        ctx.save_for_backward(pixels, light_positions, light_intensities)
        ctx.ambient = ambient_color
        return rgba

    @staticmethod
    def backward(ctx, grad_rgba):
        pixels, light_positions, light_intensities = ctx.saved_tensors
        d_pixels = ops.shade_diffuse_backward(grad_rgba.contiguous(), pixels, light_positions, light_intensities,
                                              ctx.ambient)
        return d_pixels, None, None, None


def shade_diffuse(pixels, light_positions, light_intensities, ambient_color=None):
    """RGBA [B,H,W,4] (rows flipped like phong_shader's result) from the interpolated attribute image
    `pixels` [B,H,W,A>=9] = [normal, world position, diffuse colour, ...]; gradients flow to `pixels` only."""
    return _ShadeDiffuse.apply(pixels.contiguous(), light_positions.contiguous().float(),
                               light_intensities.contiguous().float(),
                               ambient_color.contiguous().float() if ambient_color is not None else None)

class _ShadePhong(torch.autograd.Function):
    """The same with the specular branch of render.py:326-372 (two kernels each way: the reference normalises the
    reflection . view products by their L2 norm over the whole image)."""

    @staticmethod
    def forward(ctx, pixels, light_positions, light_intensities, ambient_color, camera_position, shininess):
        rgba, norm2 = ops.shade_phong_forward(pixels, light_positions, light_intensities, ambient_color,
                                              camera_position, shininess)
        ctx.save_for_backward(pixels, light_positions, light_intensities, camera_position, norm2)
        ctx.ambient, ctx.shininess = ambient_color, shininess
        return rgba

    @staticmethod
    def backward(ctx, grad_rgba):
        pixels, light_positions, light_intensities, camera_position, norm2 = ctx.saved_tensors
        d_pixels = ops.shade_phong_backward(grad_rgba.contiguous(), pixels, light_positions, light_intensities,
                                            ctx.ambient, camera_position, ctx.shininess, norm2)
        return d_pixels, None, None, None, None, None


def shade_phong(pixels, light_positions, light_intensities, camera_position, ambient_color=None, shininess=None):
    """RGBA [B,H,W,4] (rows flipped) from `pixels` [B,H,W,12] = [normal, position, diffuse, specular] with
    `shininess` [B] (or a scalar), or [B,H,W,13] with the exponent in channel 12; gradients flow to `pixels` only."""
    B = pixels.shape[0]
    if pixels.shape[3] == 12:
        shininess = torch.as_tensor(shininess, dtype=torch.float32, device=pixels.device).reshape(-1).expand(B).contiguous()
    else:
        shininess = None
    return _ShadePhong.apply(pixels.contiguous(), light_positions.contiguous().float(),
                             light_intensities.contiguous().float(),
                             ambient_color.contiguous().float() if ambient_color is not None else None,
                             camera_position.contiguous().float(), shininess)


class _RenderDiffuse(torch.autograd.Function):
    """rasterize_clip_space + shade_diffuse as ONE pass each way: the resolve kernel lights the nine interpolated
    channels while they are still in registers, the backward kernel recomputes them from the corner attributes --
    the [B,H,W,9] attribute image and its gradient are never written (render.py:183-228)."""

    @staticmethod
    def forward(ctx, clip_vertices, attributes, triangles, background, light_positions, light_intensities,
                ambient_color, image_width, image_height):
        rgba, ids, bary, _z = ops.render_diffuse_forward(clip_vertices, attributes, triangles, background,
                                                         light_positions, light_intensities, ambient_color,
                                                         image_width, image_height)
        ctx.save_for_backward(clip_vertices, attributes, triangles, background, light_positions, light_intensities,
                              ids, bary)
        ctx.ambient = ambient_color
        return rgba

    @staticmethod
    def backward(ctx, grad_rgba):
        v, a, t, bg, lp, li, ids, bary = ctx.saved_tensors
        need_v, need_a = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dv = da = None
        if need_v or need_a:
            dv, da = ops.render_diffuse_backward(grad_rgba.contiguous(), v, a, t, bg, lp, li, ctx.ambient, ids, bary,
                                                 need_vertices=need_v, need_attributes=need_a)
        return dv, da, None, None, None, None, None, None, None


def render_diffuse_clip_space(clip_vertices, attributes, triangles, light_positions, light_intensities,
                              image_width, image_height, ambient_color=None, background_value=None):
    """RGBA [B,H,W,4] of clip-space meshes whose nine per-vertex attributes are [normal, world position, diffuse
    colour]: rasterize + interpolate + diffuse/ambient Phong in one fused pass each way (gradients: clip vertices
    and attributes, accumulated atomically)."""
    device = clip_vertices.device
    if background_value is None:
        background_value = torch.full((9,), -1.0, device=device)                       # render.py:197
    return _RenderDiffuse.apply(clip_vertices.contiguous().float(), attributes.contiguous().float(),
                                triangles.contiguous(), background_value.contiguous().float(),
                                light_positions.contiguous().float(), light_intensities.contiguous().float(),
                                ambient_color.contiguous().float() if ambient_color is not None else None,
                                int(image_width), int(image_height))


# lights a fused-kernel call can take (csrc/shade.cu kMaxLights)
_MAX_FUSED_LIGHTS = 16


def _per_image(value, batch, what, device):
    """float | 0-D tensor | [batch] tensor -> float32 [batch] on `device` (render.py:126-146)."""
    if isinstance(value, (int, float)):
        return torch.full((batch,), float(value), dtype=torch.float32, device=device)
    value = value.to(device=device, dtype=torch.float32)
    if value.dim() == 0:
        return value.expand(batch).clone()
    if list(value.shape) != [batch]:
        raise ValueError("%s must be a float, a 0D tensor, or a 1D tensor with shape [batch_size]." % what)
    return value


def _per_image_vec3(value, batch, what, device):
    value = value.to(device)
    if list(value.shape) == [3]:
        return value.unsqueeze(0).expand(batch, 3)
    if list(value.shape) != [batch, 3]:
        raise ValueError("%s must have shape [batch_size, 3] or [3]." % what)
    return value


def render(vertices, triangles, normals, diffuse_colors, camera_position, camera_lookat, camera_up,
           light_positions, light_intensities, image_width, image_height, specular_colors=None,
           shininess_coefficients=None, ambient_color=None, fov_y=40.0, near_clip=0.01, far_clip=10.0):
    """Phong-shaded render of a batch of meshes -> RGBA [batch, height, width, 4].

    Arguments, shapes and ValueErrors follow render.py:16-181.  Per-vertex attributes are packed as
    [normals, positions, diffuse(, specular(, shininess))] (A = 9 / 12 / 13), interpolated by the
    rasterizer with background -1, then lit per pixel.
    """
    if vertices.dim() != 3 or vertices.shape[-1] != 3:
        raise ValueError("Vertices must have shape [batch_size, vertex_count, 3].")
    batch = vertices.shape[0]
    if normals.dim() != 3 or normals.shape[-1] != 3:
        raise ValueError("Normals must have shape [batch_size, vertex_count, 3].")
    if light_positions.dim() != 3 or light_positions.shape[-1] != 3:
        raise ValueError("light_positions must have shape [batch_size, light_count, 3].")
    if light_intensities.dim() != 3 or light_intensities.shape[-1] != 3:
        raise ValueError("light_intensities must have shape [batch_size, light_count, 3].")
    if diffuse_colors.dim() != 3 or diffuse_colors.shape[-1] != 3:
        raise ValueError("diffuse_colors must have shape [batch_size, vertex_count, 3].")
    if ambient_color is not None and list(ambient_color.shape) != [batch, 3]:
        raise ValueError("ambient_color must have shape [batch_size, 3].")
    if (specular_colors is None) != (shininess_coefficients is None):
        raise ValueError("Specular colors and shininess coefficients must be supplied together.")

    home = vertices.device
    if home.type != "cuda" and not torch.cuda.is_available():
        raise RuntimeError("pytorch_mesh_renderer_b200 needs a CUDA device; there is no CPU fallback")
    device = home if home.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
    to = lambda t: t.to(device) if t is not None else None
    vertices, normals, diffuse_colors = to(vertices), to(normals), to(diffuse_colors)
    light_positions, light_intensities, ambient_color = to(light_positions), to(light_intensities), to(ambient_color)
    camera_position = _per_image_vec3(camera_position, batch, "camera_position", device)
    camera_lookat = _per_image_vec3(camera_lookat, batch, "camera_lookat", device)
    camera_up = _per_image_vec3(camera_up, batch, "camera_up", device)
    fov_y = _per_image(fov_y, batch, "fov_y", device)
    near_clip = _per_image(near_clip, batch, "near_clip", device)
    far_clip = _per_image(far_clip, batch, "far_clip", device)

    attributes = [normals, vertices, diffuse_colors]
    per_vertex_shininess = False
    if specular_colors is not None:
        specular_colors = to(specular_colors)
        if isinstance(shininess_coefficients, (int, float)):
            shininess_coefficients = torch.tensor(float(shininess_coefficients), dtype=torch.float32)
        shininess_coefficients = to(shininess_coefficients)
        if specular_colors.dim() != 3:
            raise ValueError("The specular colors must have shape [batch_size, vertex_count, 3].")
        if shininess_coefficients.dim() > 2:
            raise ValueError("The shininess coefficients must have shape at most [batch_size, vertex_count].")
        attributes.append(specular_colors)
        if shininess_coefficients.dim() == 2:
            per_vertex_shininess = True
            attributes.append(shininess_coefficients.unsqueeze(2))
    vertex_attributes = torch.cat(attributes, 2)

    view = camera_utils.look_at(camera_position, camera_lookat, camera_up)
    projection = camera_utils.perspective(image_width / image_height, fov_y, near_clip, far_clip)
    clip_from_world = torch.matmul(projection, view)

    background = torch.full((vertex_attributes.shape[2],), -1.0, device=device)        # render.py:197
    lights_need_grad = light_positions.requires_grad or light_intensities.requires_grad or (
        ambient_color is not None and ambient_color.requires_grad)
    fusable = not lights_need_grad and light_positions.shape[1] <= _MAX_FUSED_LIGHTS
    from .rasterize_triangles_ext import get_backward_mode
    if specular_colors is None and fusable and get_backward_mode() == "atomic":
        # one fused pass each way: the attribute image is never materialised
        clip = camera_utils.transform_homogeneous(clip_from_world, vertices)
        image = render_diffuse_clip_space(clip, vertex_attributes, triangles.to(device), light_positions,
                                          light_intensities, image_width, image_height, ambient_color, background)
        return image.to(home) if home != device else image

    pixels = rasterize(vertices, vertex_attributes, triangles.to(device), clip_from_world, image_width,
                       image_height, background)

    if specular_colors is None and fusable:
        image = shade_diffuse(pixels, light_positions, light_intensities, ambient_color)
        return image.to(home) if home != device else image
    if (specular_colors is not None and fusable and not camera_position.requires_grad
            and (per_vertex_shininess or not shininess_coefficients.requires_grad)
            and (per_vertex_shininess or shininess_coefficients.numel() in (1, batch))):
        image = shade_phong(pixels, light_positions, light_intensities, camera_position, ambient_color,
                            None if per_vertex_shininess else shininess_coefficients)
        return image.to(home) if home != device else image

    pixel_normals = F.normalize(pixels[..., 0:3], p=2, dim=3)
    pixel_positions = pixels[..., 3:6]
    pixel_diffuse = pixels[..., 6:9]
    pixel_specular = pixel_shininess = None
    if specular_colors is not None:
        pixel_specular = pixels[..., 9:12]
        pixel_shininess = pixels[..., 12] if per_vertex_shininess else shininess_coefficients.reshape(-1, 1, 1)
    # background pixels carry diffuse = -1 in every channel (render.py:215)
    mask = (pixel_diffuse >= 0.0).any(dim=3).to(torch.float32)
    image = phong_shader(pixel_normals, mask, pixel_positions, light_positions, light_intensities, pixel_diffuse,
                         camera_position if specular_colors is not None else None, pixel_specular,
                         pixel_shininess, ambient_color)
    return image.to(home) if home != device else image


def phong_shader(normals, alphas, pixel_positions, light_positions, light_intensities, diffuse_colors=None,
                 camera_position=None, specular_colors=None, shininess_coefficients=None, ambient_color=None):
    """Per-pixel Phong lighting of rasterized buffers -> RGBA [batch, height, width, 4], flipped
    vertically (row 0 of the result is the top of the image).  Mirrors render.py:231-386, including
    its per-(image, light) L2 normalisation of the specular dot products over all pixels (:347-353)."""
    batch, height, width = normals.shape[:3]
    n = normals.reshape(batch, 1, -1, 3)                                   # [B,1,P,3]
    pos = pixel_positions.reshape(batch, 1, -1, 3)
    kd = diffuse_colors.reshape(batch, -1, 3)
    to_light = F.normalize(light_positions.unsqueeze(2) - pos, p=2, dim=3)               # [B,L,P,3]
    n_dot_l = torch.clamp((n * to_light).sum(3), 0.0, 1.0)                               # [B,L,P]
    rgb = (kd.unsqueeze(1) * n_dot_l.unsqueeze(3) * light_intensities.unsqueeze(2)).sum(1)
    if ambient_color is not None:
        rgb = ambient_color.unsqueeze(1) * kd + rgb
    if camera_position is not None:
        ks = specular_colors.reshape(batch, -1, 3)
        mirror = F.normalize(2.0 * n_dot_l.unsqueeze(3) * n - to_light, p=2, dim=3)
        to_camera = F.normalize(camera_position.reshape(batch, 1, 3) - pos[:, 0], p=2, dim=2)       # [B,P,3]
        r_dot_v = (mirror * to_camera.unsqueeze(1)).sum(3)                                         # [B,L,P]
        r_dot_v = torch.clamp(F.normalize(r_dot_v, p=2, dim=2), 0.0, 1.0)
        r_dot_v = torch.where(n_dot_l != 0.0, r_dot_v, torch.zeros_like(r_dot_v))
        shininess = shininess_coefficients.unsqueeze(1)                    # broadcasts over [B,L,H,W]
        specularity = torch.pow(r_dot_v.reshape(batch, -1, height, width), shininess).reshape(batch, -1, height * width, 1)
        rgb = rgb + (ks.unsqueeze(1) * specularity * light_intensities.unsqueeze(2)).sum(1)
    rgb = rgb.reshape(batch, height, width, 3)
    alpha = alphas.reshape(batch, height, width, 1)
    rgb = torch.where(alpha > 0.5, rgb, torch.zeros_like(rgb))
    return torch.flip(torch.cat([rgb, alpha], 3), dims=[1])


def tone_mapper(image, gamma):
    """Per-image gamma correction scaled so that the image maximum becomes 1, clipped to [0, 1]
    (render.py:389-419)."""
    corrected = torch.pow(image, gamma)
    peak = corrected.reshape(image.shape[0], -1).max(1).values.reshape(-1, 1, 1, 1)
    return torch.clamp(corrected / peak, 0.0, 1.0)
