"""Counterpart of the reference's src/common/meshes.py: `compute_vertex_normals(vertices, triangles)`, same
name, arguments and result, running on the CUDA kernels of csrc/mesh_normals.cu (include/pmr_b200.h
pmr_vertex_*), differentiable with respect to the vertices.

The reference scatters face normals with three index_add_ calls per mesh in a Python loop over the batch
(meshes.py:19-33).  Here the topology becomes a vertex -> (corner, triangle) table once (cached per
triangles tensor) and both directions are gathers over it: one launch forward, two backward, for the whole
batch, with the sums in the reference's order (bit-identical to the reference on the CPU).

Not reproduced: for a mesh of exactly three triangles the reference's `torch.cross` call, which passes no
`dim`, falls under torch's deprecated rule "first dimension of size 3" and takes the product along the
triangle axis; this module always uses the coordinate axis.

CPU tensors are moved to the current CUDA device and the result moved back (there is no CPU path).
"""
import collections

import torch

from . import ops

_INCIDENCE_CACHE = collections.OrderedDict()
_INCIDENCE_CACHE_SIZE = 16


def _incidence(triangles, vertex_count):
    """(offsets, incidence) of this topology; rebuilt when the triangles tensor is replaced or written to."""
    key = (triangles.data_ptr(), triangles._version, triangles.shape[0], int(vertex_count), triangles.device)
    hit = _INCIDENCE_CACHE.get(key)
    if hit is not None:
        _INCIDENCE_CACHE.move_to_end(key)
        return hit[0], hit[1]
    offsets, incidence = ops.vertex_incidence(triangles, vertex_count)
    # the entry keeps the triangles tensor alive, so its address cannot be reused while the entry exists
    _INCIDENCE_CACHE[key] = (offsets, incidence, triangles)
    while len(_INCIDENCE_CACHE) > _INCIDENCE_CACHE_SIZE:
        _INCIDENCE_CACHE.popitem(last=False)
    return offsets, incidence


class VertexNormals(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vertices, triangles):
        offsets, incidence = _incidence(triangles, vertices.shape[1])
        normals, raw = ops.vertex_normals_forward(vertices, triangles, offsets, incidence)
        ctx.save_for_backward(vertices, triangles, offsets, incidence, raw)
        return normals

    @staticmethod
    def backward(ctx, grad_normals):
        vertices, triangles, offsets, incidence, raw = ctx.saved_tensors
        return ops.vertex_normals_backward(grad_normals.contiguous(), raw, vertices, triangles, offsets,
                                           incidence), None


def compute_vertex_normals(vertices, triangles):
    """Per-vertex normals [batch_size, vertex_count, 3] of a triangle mesh: the (area-weighted) face normals
    summed on their three vertices, then normalised (meshes.py:3-35).

    vertices: float32 [batch_size, vertex_count, 3]; triangles: int32 [triangle_count, 3].
    """
    if vertices.dim() != 3 or vertices.shape[2] != 3:
        raise ValueError("vertices must have shape [batch_size, vertex_count, 3]")
    if triangles.dim() != 2 or triangles.shape[1] != 3:
        raise ValueError("triangles must have shape [triangle_count, 3]")
    home = vertices.device
    if vertices.is_cuda:
        dev = vertices.device
    else:
        if not torch.cuda.is_available():
            raise RuntimeError("pytorch_mesh_renderer_b200 needs a CUDA device; there is no CPU fallback")
        dev = torch.device("cuda", torch.cuda.current_device())
    tri = triangles if triangles.device == dev and triangles.dtype == torch.int32 else \
        triangles.to(device=dev, dtype=torch.int32)
    normals = VertexNormals.apply(vertices.to(dev), tri.contiguous())
    return normals.to(home) if home != dev else normals
