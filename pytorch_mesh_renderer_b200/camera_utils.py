"""Device-aware counterpart of the one camera helper on the hot path's doorstep:
transform_homogeneous (reference src/common/camera_utils.py:142-170), which the reference
allocates on the CPU regardless of its inputs."""
import torch


def transform_homogeneous(matrices, vertices):
    """Applies batched 4x4 homogeneous transforms to xyz vertices: (M V^T)^T with w = 1.

    matrices [B,4,4], vertices [B,N,3] -> [B,N,4].  Raises ValueError on wrong rank like
    camera_utils.py:159-164.
    """
    if len(matrices.shape) != 3:
        raise ValueError("matrices must have 3 dimensions (missing batch dimension?)")
    if len(vertices.shape) != 3:
        raise ValueError("vertices must have 3 dimensions (missing batch dimension?)")
    ones = torch.ones([vertices.shape[0], vertices.shape[1], 1], dtype=torch.float32, device=vertices.device)
    homogeneous = torch.cat([vertices, ones], 2)
    return torch.matmul(homogeneous, matrices.to(vertices.device).transpose(1, 2))
