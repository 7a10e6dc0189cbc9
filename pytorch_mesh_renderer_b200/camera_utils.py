"""Device-aware counterpart of the one camera helper on the hot path's doorstep:
transform_homogeneous (reference src/common/camera_utils.py:142-170), which the reference allocates on
the CPU regardless of its inputs.  World -> clip space runs in one CUDA kernel
(csrc/vertex_stage.cu); its backward sums the per-view clip-space gradients of a shared mesh on the
fly, which is the on-device reduction the multi-GPU path all-reduces afterwards."""
import torch

from . import ops


class _TransformVertices(torch.autograd.Function):
    """matrices [B,4,4] x vertices ([V,3] shared by all views, or [B,V,3]) -> clip [B,V,4]."""

    @staticmethod
    def forward(ctx, matrices, vertices, exchange=None):
        shared = vertices.dim() == 2
        ctx.save_for_backward(matrices, vertices)
        ctx.shared = shared
        ctx.exchange = exchange if shared else None
        return ops.transform_forward(matrices, vertices, shared)

    @staticmethod
    def backward(ctx, d_clip):
        matrices, vertices = ctx.saved_tensors
        d_matrices = d_vertices = None
        if ctx.needs_input_grad[1]:
            if ctx.exchange is not None:
                # multi-GPU: partial over the local views + sum over the ranks through peer memory, fused
                d_vertices = ctx.exchange.reduce(matrices, ops._aligned(d_clip.contiguous()))
            else:
                d_vertices = ops.transform_backward(matrices, d_clip.contiguous(), ctx.shared)
        if ctx.needs_input_grad[0]:
            # camera fitting (reference example4): d M_b = sum_v d_clip[b,v] (x) (x, y, z, 1); off the hot path
            w = vertices.unsqueeze(0).expand(matrices.shape[0], -1, -1) if ctx.shared else vertices
            hom = torch.cat([w, torch.ones_like(w[..., :1])], 2)
            d_matrices = torch.einsum("bvi,bvj->bij", d_clip, hom)
        return d_matrices, d_vertices, None


def _on_device(*tensors):
    for t in tensors:
        if t.is_cuda:
            return t.device
    if not torch.cuda.is_available():
        raise RuntimeError("pytorch_mesh_renderer_b200 needs a CUDA device; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def transform_homogeneous(matrices, vertices):
    """Applies batched 4x4 homogeneous transforms to xyz vertices: (M V^T)^T with w = 1.

    matrices [B,4,4], vertices [B,N,3] -> [B,N,4].  Raises ValueError on wrong rank like
    camera_utils.py:159-164.  CPU tensors are processed on the current CUDA device and returned
    on the CPU.
    """
    if len(matrices.shape) != 3:
        raise ValueError("matrices must have 3 dimensions (missing batch dimension?)")
    if len(vertices.shape) != 3:
        raise ValueError("vertices must have 3 dimensions (missing batch dimension?)")
    home = vertices.device
    dev = _on_device(matrices, vertices)
    out = _TransformVertices.apply(matrices.to(dev).float(), vertices.to(dev).float())
    return out.to(home) if home != dev else out


def transform_shared_mesh(matrices, vertices, exchange=None):
    """One mesh [V,3] seen by B views: matrices [B,4,4] -> clip [B,V,4]; the gradient with respect to
    `vertices` comes back as [V,3], already summed over the views -- and, with `exchange` (a
    distributed.SharedGradientExchange), over the views of ALL ranks."""
    if len(matrices.shape) != 3 or len(vertices.shape) != 2:
        raise ValueError("expected matrices [B,4,4] and vertices [V,3]")
    dev = _on_device(matrices, vertices)
    return _TransformVertices.apply(matrices.to(dev).float(), vertices.to(dev).float(), exchange)


# ---------------------------------------------------------------------------------------------
# Camera builders (reference src/common/camera_utils.py:10-139), device-aware: every tensor is
# created on the device of the inputs, no numpy round trips, no host synchronisation.
# ---------------------------------------------------------------------------------------------

def euler_matrices(angles):
    """XYZ Tait-Bryan rotations, [N,3] angles in radians -> [N,4,4] (camera_utils.py:10-42)."""
    sx, sy, sz = torch.sin(angles).unbind(1)
    cx, cy, cz = torch.cos(angles).unbind(1)
    zero, one = torch.zeros_like(sx), torch.ones_like(sx)
    rows = [cz * cy, cz * sy * sx - cx * sz, sz * sx + cz * cx * sy, zero,
            cy * sz, cz * cx + sz * sy * sx, cx * sz * sy - cz * sx, zero,
            -sy, cy * sx, cy * cx, zero,
            zero, zero, zero, one]
    return torch.stack(rows, 1).reshape(-1, 4, 4)


def look_at(eye, center, world_up):
    """gluLookAt: world -> eye space matrices [N,4,4] from [N,3] eye, gaze target and up vector
    (camera_utils.py:45-96).  Degenerate inputs (eye == center, up parallel to the gaze) raise
    ValueError; the check costs one device-to-host read, like the reference's numpy asserts."""
    forward = center - eye
    forward_norm = torch.linalg.norm(forward, dim=1, keepdim=True)
    side = torch.cross(forward / forward_norm, world_up, dim=-1)
    side_norm = torch.linalg.norm(side, dim=1, keepdim=True)
    if bool((forward_norm <= 1e-6).any()) or bool((side_norm <= 1e-6).any()):
        raise ValueError("Camera matrix is degenerate (eye and center coincide, or up is parallel to the gaze).")
    forward = forward / forward_norm
    side = side / side_norm
    up = torch.cross(side, forward, dim=-1)
    rotation = torch.stack([side, up, -forward], 1)                       # [N,3,3]
    translation = -torch.einsum("nij,nj->ni", rotation, eye)
    top = torch.cat([rotation, translation.unsqueeze(2)], 2)               # [N,3,4]
    bottom = torch.zeros_like(top[:, :1, :])
    bottom[:, 0, 3] = 1.0
    return torch.cat([top, bottom], 1)


def perspective(aspect_ratio, fov_y, near_clip, far_clip):
    """gluPerspective: [N] fov (degrees), near, far -> [N,4,4] (camera_utils.py:99-139)."""
    import math
    f = 1.0 / torch.tan(fov_y * (math.pi / 360.0))
    depth = far_clip - near_clip
    m = torch.zeros((fov_y.shape[0], 4, 4), dtype=torch.float32, device=fov_y.device)
    m[:, 0, 0] = f / aspect_ratio
    m[:, 1, 1] = f
    m[:, 2, 2] = -(far_clip + near_clip) / depth
    m[:, 2, 3] = -2.0 * (far_clip * near_clip / depth)
    m[:, 3, 2] = -1.0
    return m
