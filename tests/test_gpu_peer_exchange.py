"""Multi-GPU exchange of the shared-mesh gradient through peer memory (csrc/peer_exchange.cu,
distributed.SharedGradientExchange): one process per GPU, compared with the two-step path it replaces
(transform_backward kernel, then NCCL all-reduce).  Needs at least two GPUs on the box; skipped otherwise.

With two ranks a + b is the same number in either order, so the fused result must equal the NCCL result bit for
bit; with more ranks every rank must hold the SAME bits (rank-order sums) and agree with NCCL within rounding."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, V, B, steps, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    device = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=device)
    from pytorch_mesh_renderer_b200 import distributed as D, ops
    from pytorch_mesh_renderer_b200.camera_utils import transform_shared_mesh
    ex = D.SharedGradientExchange.create(V, device)
    assert ex is not None, "peer exchange could not be set up"
    g = torch.Generator().manual_seed(100 + rank)
    mvp = torch.randn((B, 4, 4), generator=g).to(device)
    world_vertices = torch.randn((V, 3), generator=g).to(device)
    results = []
    for step in range(steps):
        d_clip = torch.randn((B, V, 4), generator=g).to(device)
        # through autograd, as a training step does it
        wv = world_vertices.detach().requires_grad_(True)
        transform_shared_mesh(mvp, wv, exchange=ex).backward(d_clip)
        fused = wv.grad.clone()
        two_step = ops.transform_backward(mvp, d_clip, True)
        dist.all_reduce(two_step)
        if world == 2:
            assert torch.equal(fused, two_step), (rank, step, float((fused - two_step).abs().max()))
        else:
            scale = float(two_step.abs().max())
            assert float((fused - two_step).abs().max()) <= 1e-5 * scale, (rank, step)
        results.append(fused.cpu().numpy())
    assert not ex.timed_out()
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), np.stack(results))
    ex.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("V,B", [(1000, 3), (25124, 8)])
def test_fused_exchange_matches_kernel_plus_nccl(tmp_path, V, B):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)
    mp.spawn(_worker, args=(world, _free_port(), V, B, 5, str(tmp_path)), nprocs=world, join=True)
    first = np.load(os.path.join(str(tmp_path), "rank0.npy"))
    for r in range(1, world):
        assert np.array_equal(first, np.load(os.path.join(str(tmp_path), "rank%d.npy" % r))), r
