"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo process group
(view sharding and the single packed all-reduce of shared-parameter gradients)."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_views, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pytorch_mesh_renderer_b200 import distributed as D
    mine = D.shard_views(n_views)
    # Each view contributes a known gradient to the shared [V,3] vertices and [V,A] attributes.
    V, A = 7, 4
    d_world = torch.zeros(V, 3)
    d_attr = torch.zeros(V, A)
    for v in mine:
        g = torch.Generator().manual_seed(100 + v)
        d_world += torch.randn(V, 3, generator=g)
        d_attr += torch.randn(V, A, generator=g)
    # the peer-memory exchange is a CUDA / NCCL feature: on a CPU group every rank gets None (after the same
    # collectives on every rank) and the caller all-reduces
    assert D.SharedGradientExchange.create(V, torch.device("cpu")) is None
    D.all_reduce_gradients([d_world, d_attr])
    out.put((rank, list(mine), d_world.numpy(), d_attr.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_views", [5, 8])
def test_shard_and_all_reduce_world_size_2(n_views):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_views, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    results.sort()
    views = results[0][1] + results[1][1]
    assert sorted(views) == list(range(n_views))                      # every view exactly once
    assert abs(len(results[0][1]) - len(results[1][1])) <= 1          # balanced
    V, A = 7, 4
    want_w, want_a = np.zeros((V, 3), np.float32), np.zeros((V, A), np.float32)
    for v in range(n_views):
        g = torch.Generator().manual_seed(100 + v)
        want_w += torch.randn(V, 3, generator=g).numpy()
        want_a += torch.randn(V, A, generator=g).numpy()
    for _, _, w, a in results:                                         # both ranks hold the full sum
        np.testing.assert_allclose(w, want_w, rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(a, want_a, rtol=1e-6, atol=1e-6)


def test_shard_views_partitions():
    from pytorch_mesh_renderer_b200.distributed import shard_views
    for n in (0, 1, 7, 64, 256):
        for world in (1, 2, 3, 8):
            parts = [list(shard_views(n, r, world)) for r in range(world)]
            assert sum(parts, []) == list(range(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1
            dealt = [list(shard_views(n, r, world, interleaved=True)) for r in range(world)]
            assert sorted(sum(dealt, [])) == list(range(n))                      # a partition as well
            assert max(map(len, dealt)) - min(map(len, dealt)) <= 1
            assert all(v % world == r for r, part in enumerate(dealt) for v in part)
