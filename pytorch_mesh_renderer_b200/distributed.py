"""Multi-GPU plumbing: one process per GPU, views (images) sharded across ranks.

The forward pass needs no communication (every view is rasterized independently, like the
reference's `for b in range(batch_size)` loop, rasterize.py:112).  The only exchange on the path is
the gradient of parameters that all views share -- the world-space mesh in multi-view fitting
(the `torch.stack([vertices] * n)` pattern of the reference's example7b.py:225): every rank reduces
its local views' clip-space gradients to one world-space [V,3] tensor on the device (autograd of
`transform_homogeneous` + the broadcast), then ONE all-reduce(sum) moves it over NVLink (NCCL).
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib


def shard_views(n_views, rank=None, world_size=None, interleaved=False):
    """The views (batch entries) owned by `rank`, as a range: a contiguous slice range(start, stop), or with
    `interleaved` every world_size-th view, range(rank, n_views, world_size).

    Ranks get floor(n/world) or ceil(n/world) views; every view belongs to exactly one rank.  Interleave when
    neighbouring views cost alike (an orbit of cameras: views near the poles of a UV sphere rasterize slivers):
    a step takes as long as the slowest rank, and on c4 over 8 GPUs the contiguous shares differ by 13 %
    (0.66 .. 0.75 ms), the interleaved ones by 0.6 % (profiles/tools/slice_times.py).
    """
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if interleaved:
        return range(int(rank), int(n_views), int(world_size))
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


class SharedGradientExchange:
    """Sum of the shared-mesh gradient over the ranks through peer memory (include/pmr_b200.h
    pmr_transform_backward_exchange, csrc/peer_exchange.cu) instead of kernel + NCCL all-reduce: the backward of
    the vertex stage stores this rank's partial [V,3] into every peer's exchange buffer over NVLink, and a second
    kernel adds the world's partials in rank order.  All ranks end with the bit-identical sum; the step contains
    no collective call and no host synchronisation.

    One process per GPU of one box (CUDA IPC maps the peers' buffers).  `create` returns None when the exchange
    cannot be set up (single rank, CPU group, no peer access, more than 16 ranks): callers then all-reduce with
    `all_reduce_gradients`.  Used by passing it to `camera_utils.transform_shared_mesh(..., exchange=ex)`, whose
    backward then returns the COMPLETE gradient; every rank must run the same sequence of steps.
    """

    MAX_PEERS = 16

    @classmethod
    def create(cls, vertex_count, device, group=None):
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return None
        world = dist.get_world_size(group)
        device = torch.device(device)
        ok = (device.type == "cuda" and world <= cls.MAX_PEERS and dist.get_backend(group) == "nccl"
              and torch.cuda.device_count() >= world)
        ex = None
        if ok:
            try:
                ex = cls(int(vertex_count), device, group)
            except (_lib.PmrError, ValueError):
                ex = None
        # all ranks or none: a rank that failed makes everybody fall back
        votes = torch.tensor([1 if ex is not None else 0], device=device if device.type == "cuda" else "cpu")
        dist.all_reduce(votes, op=dist.ReduceOp.MIN, group=group)
        if int(votes.item()) == 0:
            if ex is not None:
                ex.close(collective=False)
            return None
        return ex

    def __init__(self, vertex_count, device, group=None):
        """Collective over the group.  Every rank goes through the same sequence of collectives whatever fails
        locally (a rank that raised early would leave the others waiting); the failure is raised at the end."""
        self.group, self.device = group, device
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.vertex_count = vertex_count
        self.epoch = 0
        self.opened = []
        self.own = None
        lib = _lib.load()
        self.ctx = _lib.context(device.index)
        failure = None
        nbytes = lib.pmr_peer_exchange_bytes(3 * vertex_count, self.world)
        own, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
        with torch.cuda.device(device):
            if nbytes == 0 or lib.pmr_peer_alloc(self.ctx, nbytes, ctypes.byref(own), handle) != 0:
                failure = (lib.pmr_last_error(self.ctx) or b"exchange buffer could not be allocated").decode()
            else:
                self.own = own
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw if failure is None else None, group=group)
        if failure is None and any(h is None for h in handles):
            failure = "a peer could not allocate its exchange buffer"
        self.peers = (ctypes.c_void_p * self.world)()
        if failure is None:
            for r in range(self.world):
                if r == self.rank:
                    self.peers[r] = own.value
                    continue
                p = ctypes.c_void_p()
                with torch.cuda.device(device):
                    rc = lib.pmr_peer_open(self.ctx, handles[r], ctypes.byref(p))
                if rc != 0:
                    failure = (lib.pmr_last_error(self.ctx) or b"peer buffer could not be mapped").decode()
                    break
                self.opened.append(p)
                self.peers[r] = p.value
        dist.barrier(group=group)          # every buffer is zero-filled and mapped before the first store
        if failure is not None:
            self.close(collective=False)
            raise _lib.PmrError("peer exchange: " + failure)

    def reduce(self, matrices, d_clip):
        """matrices [B,4,4], d_clip [B,V,4] of this rank's views -> d_world [V,3] summed over ALL ranks' views."""
        B, V, _ = d_clip.shape
        if V != self.vertex_count:
            raise ValueError("exchange was created for %d vertices, got %d" % (self.vertex_count, V))
        self.epoch += 1                    # bookkeeping only: the device counts the steps itself (epoch argument 0),
        #                                    so that the call has no per-step argument and replays from a CUDA graph
        out = torch.empty((V, 3), dtype=torch.float32, device=d_clip.device)
        with torch.cuda.device(d_clip.device):
            rc = _lib.load().pmr_transform_backward_exchange(
                self.ctx, _lib.ptr(matrices), _lib.ptr(d_clip), B, V, self.peers, self.rank, self.world, 0,
                _lib.ptr(out), _lib.stream_ptr(d_clip.device))
        _lib.check(self.ctx, rc)
        return out

    def timed_out(self):
        """True if a wait for a peer ever gave up (synchronises the device).  From that moment every reduce() on
        these buffers returns NaN instead of a sum of stale partials; check before the optimizer step when a rank
        may stall for longer than PMR_PEER_WAIT_SECONDS (default 10 s)."""
        status = ctypes.c_int(0)
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            _lib.check(self.ctx, _lib.load().pmr_peer_status(self.ctx, self.own, ctypes.byref(status)))
        return status.value != 0

    def close(self, collective=True):
        if self.own is None and not self.opened:
            return
        lib = _lib.load()
        failed = False
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            if self.own is not None:
                status = ctypes.c_int(0)
                failed = lib.pmr_peer_status(self.ctx, self.own, ctypes.byref(status)) == 0 and status.value != 0
            if collective:
                dist.barrier(group=self.group)     # nobody stores into a buffer that is about to go away
            for p in self.opened:
                lib.pmr_peer_close(self.ctx, p)
            if self.own is not None:
                lib.pmr_peer_free(self.ctx, self.own)
        self.opened, self.own = [], None
        if failed:
            raise _lib.PmrError("peer exchange: a wait for a peer's partial sums gave up during this run; the "
                                "gradients returned from that step on were NaN")


def rasterize_shared_mesh(world_vertices, attributes, triangles, camera_matrices, image_width, image_height,
                          background_value, exchange=None):
    """Renders this rank's views of ONE shared mesh.

    world_vertices [V,3] and attributes [V,A] are shared by all views (and all ranks);
    camera_matrices [B_local,4,4] are this rank's views.  Returns the attribute images
    [B_local,H,W,A].  After `loss.backward()`, `world_vertices.grad` / `attributes.grad` hold this
    rank's partial sums; `all_reduce_gradients([...])` completes them.  With `exchange` (a
    SharedGradientExchange) `world_vertices.grad` is already the sum over all ranks (attributes.grad is not).
    """
    from .camera_utils import transform_shared_mesh
    from .rasterize import rasterize_clip_space
    B = camera_matrices.shape[0]
    clip = transform_shared_mesh(camera_matrices, world_vertices, exchange=exchange)   # backward sums over views
    attrs = attributes.unsqueeze(0).expand(B, -1, -1)
    return rasterize_clip_space(clip, attrs, triangles, image_width, image_height, background_value)
