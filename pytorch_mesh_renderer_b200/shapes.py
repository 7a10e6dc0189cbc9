"""Counterpart of the reference's src/common/shapes.py: `sphere(radius, resolution)` and `cube(size)` return
(vertices [V,3] float32, triangles [T,3] int32, normals [V,3] float32) on the CPU, element for element what the
reference's generators return (tests/test_mesh_utils.py compares them with the reference's output digests).

The sphere is built with array arithmetic instead of the reference's Python double loop (shapes.py:44-82), but
the construction is the same, including its particulars, which a user switching over would otherwise see as a
different mesh: K = resolution longitudes x K latitude rings + 2 poles (shapes.py:32-36); the last quad of every
ring and the last fan triangle of each pole index the NEXT vertex id rather than wrapping around the ring
(shapes.py:59-82), and the poles sit at (0, +-1, 0) whatever the radius (shapes.py:54-55).
"""
import numpy as np
import torch


def sphere(radius, resolution=25):
    K = int(resolution)
    theta = np.linspace(np.pi / (K + 1), np.pi - np.pi / (K + 1), K, endpoint=True)      # shapes.py:33,47
    phi = np.linspace(0.0, 2.0 * np.pi, K, endpoint=False)                               # shapes.py:48
    st, ct = np.sin(theta)[:, None], np.cos(theta)[:, None]
    ring = np.stack([st * np.sin(phi)[None, :], np.broadcast_to(ct, (K, K)), st * np.cos(phi)[None, :]], axis=2)
    vertices = torch.zeros([K * K + 2, 3], dtype=torch.float32)
    # the product is formed in float64 and rounded once (the reference multiplies a float64 tensor, shapes.py:49-53)
    vertices[:K * K] = torch.from_numpy((radius * ring.reshape(K * K, 3)).astype(np.float32))
    vertices[K * K] = torch.tensor([0.0, 1.0, 0.0])
    vertices[K * K + 1] = torch.tensor([0.0, -1.0, 0.0])

    i, j = np.meshgrid(np.arange(K - 1), np.arange(K), indexing="ij")
    top_left, top_right = i * K + j, i * K + j + 1
    bottom_left, bottom_right = (i + 1) * K + j, (i + 1) * K + j + 1
    quads = np.stack([np.stack([top_left, bottom_left, top_right], axis=2),
                      np.stack([top_right, bottom_left, bottom_right], axis=2)], axis=2)   # [K-1, K, 2, 3]
    k = np.arange(K)
    top = np.stack([np.full(K, K * K), k, k + 1], axis=1)                                   # shapes.py:68-73
    bottom = np.stack([np.full(K, K * K + 1), (K - 1) * K + k + 1, (K - 1) * K + k], axis=1)  # shapes.py:75-80
    triangles = torch.tensor(np.concatenate([quads.reshape(-1, 3), top, bottom], axis=0), dtype=torch.int32)
    normals = torch.nn.functional.normalize(vertices, p=2.0, dim=-1)
    return vertices, triangles, normals


_CUBE_CORNERS = [[-1, -1, 1], [-1, -1, -1], [-1, 1, -1], [-1, 1, 1], [1, -1, 1], [1, -1, -1], [1, 1, -1], [1, 1, 1]]
_CUBE_FACES = [[2, 1, 0], [0, 3, 2], [6, 2, 3], [3, 7, 6], [5, 6, 7], [7, 4, 5],
               [1, 5, 4], [4, 0, 1], [2, 6, 5], [5, 1, 2], [0, 4, 7], [7, 3, 0]]


def cube(size):
    """Axis-aligned cube of side `size` centred on the origin; normals point from the centre to the corners
    (shapes.py:85-117)."""
    vertices = 0.5 * size * torch.tensor(_CUBE_CORNERS, dtype=torch.float32)
    normals = torch.nn.functional.normalize(vertices, p=2.0, dim=-1)
    triangles = torch.tensor(_CUBE_FACES, dtype=torch.int32)
    return vertices, triangles, normals
