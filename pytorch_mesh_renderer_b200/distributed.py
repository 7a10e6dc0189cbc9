"""Multi-GPU plumbing: one process per GPU, views (images) sharded across ranks.

The forward pass needs no communication (every view is rasterized independently, like the
reference's `for b in range(batch_size)` loop, rasterize.py:112).  The only exchange on the path is
the gradient of parameters that all views share -- the world-space mesh in multi-view fitting
(the `torch.stack([vertices] * n)` pattern of the reference's example7b.py:225): every rank reduces
its local views' clip-space gradients to one world-space [V,3] tensor on the device (autograd of
`transform_homogeneous` + the broadcast), then ONE all-reduce(sum) moves it over NVLink (NCCL).
"""
import torch
import torch.distributed as dist


def shard_views(n_views, rank=None, world_size=None):
    """Contiguous slice of the view (batch) dimension owned by `rank`: range(start, stop).

    Ranks get floor(n/world) or ceil(n/world) views; every view belongs to exactly one rank.
    """
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    base, extra = divmod(int(n_views), int(world_size))
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def all_reduce_gradients(tensors, group=None):
    """Sums each tensor over all ranks with a single collective.

    The tensors (e.g. the [V,3] world-vertex gradient and, when attributes are shared too, the
    [V,A] attribute gradient) are packed into one flat buffer so that exactly one all-reduce is
    issued per step: the message is small (0.6 MB at V = 50 k), so its cost is launch latency, not
    bandwidth.  Works with NCCL (CUDA tensors) and gloo (CPU tensors).  Returns the tensors,
    updated in place.
    """
    tensors = [t for t in tensors if t is not None]
    if not tensors or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return tensors
    if len(tensors) == 1 and tensors[0].is_contiguous():
        dist.all_reduce(tensors[0], op=dist.ReduceOp.SUM, group=group)
        return tensors
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    offset = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[offset:offset + n].view_as(t))
        offset += n
    return tensors


def rasterize_shared_mesh(world_vertices, attributes, triangles, camera_matrices, image_width, image_height,
                          background_value):
    """Renders this rank's views of ONE shared mesh.

    world_vertices [V,3] and attributes [V,A] are shared by all views (and all ranks);
    camera_matrices [B_local,4,4] are this rank's views.  Returns the attribute images
    [B_local,H,W,A].  After `loss.backward()`, `world_vertices.grad` / `attributes.grad` hold this
    rank's partial sums; `all_reduce_gradients([...])` completes them.
    """
    from .camera_utils import transform_shared_mesh
    from .rasterize import rasterize_clip_space
    B = camera_matrices.shape[0]
    clip = transform_shared_mesh(camera_matrices, world_vertices)      # one kernel; backward sums over views
    attrs = attributes.unsqueeze(0).expand(B, -1, -1)
    return rasterize_clip_space(clip, attrs, triangles, image_width, image_height, background_value)
