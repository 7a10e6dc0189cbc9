"""Host-side logic of distributed.SharedGradientExchange on the CPU, with the library, the process group and the
CUDA context replaced by fakes: the sequence of collectives is the same on every rank whatever fails locally, the
peer table is filled correctly, epochs advance, and failures make `create` fall back on ALL ranks.  (The kernels and
the real CUDA IPC path are covered by tests/test_gpu_peer_exchange.py on a multi-GPU box.)"""
import contextlib
import ctypes

import pytest
import torch

from pytorch_mesh_renderer_b200 import _lib, distributed as D


class FakeLib:
    """Records calls; hands out fake device pointers; can be told to fail."""

    def __init__(self, rank, fail_alloc=False, fail_open_of=None):
        self.rank, self.fail_alloc, self.fail_open_of = rank, fail_alloc, fail_open_of
        self.calls, self.closed, self.freed, self.epochs = [], [], [], []

    def pmr_peer_exchange_bytes(self, n_floats, world):
        return 4352 + 2 * world * ((n_floats + 3) // 4 * 4) * 4

    def pmr_peer_alloc(self, ctx, nbytes, ptr_ref, handle):
        self.calls.append("alloc")
        if self.fail_alloc:
            return -2
        ptr_ref._obj.value = 0x1000 * (self.rank + 1)
        handle.raw = bytes([self.rank + 1]) * 64
        return 0

    def pmr_peer_open(self, ctx, handle, ptr_ref):
        owner = handle[0] - 1
        self.calls.append("open%d" % owner)
        if owner == self.fail_open_of:
            return -2
        ptr_ref._obj.value = 0x1000 * (owner + 1) + 0x100000 * (self.rank + 1)      # the peer's buffer as mapped here
        return 0

    def pmr_peer_close(self, ctx, p):
        self.closed.append(p.value)
        return 0

    def pmr_peer_free(self, ctx, p):
        self.freed.append(p.value)
        return 0

    def pmr_last_error(self, ctx):
        return b"fake failure"

    def pmr_peer_status(self, ctx, own, status_ref):
        status_ref._obj.value = 0
        return 0

    def pmr_transform_backward_exchange(self, ctx, m, g, B, V, peers, rank, world, epoch, out, stream):
        self.epochs.append(epoch)
        self.peers_seen = [peers[r] for r in range(world)]
        return 0


class FakeGroup:
    """A world of `world` ranks run one after the other: collectives are recorded per rank and the gathered objects
    are exchanged through a shared list (rank r's contribution is known before anyone reads, because the ranks'
    contributions are produced by the same deterministic fake)."""

    def __init__(self, world, contributions):
        self.world, self.contributions = world, contributions
        self.log = {r: [] for r in range(world)}


@pytest.fixture
def fake_world(monkeypatch):
    def run(world, fail_alloc_on=(), fail_open=None, device_type="cuda"):
        # what every rank contributes to all_gather_object (None = its allocation failed)
        contributions = [None if r in fail_alloc_on else bytes([r + 1]) * 64 for r in range(world)]
        group = FakeGroup(world, contributions)
        results, libs = {}, {}
        for rank in range(world):
            lib = FakeLib(rank, fail_alloc=rank in fail_alloc_on,
                          fail_open_of=fail_open[1] if fail_open and fail_open[0] == rank else None)
            libs[rank] = lib
            monkeypatch.setattr(_lib, "load", lambda lib=lib: lib)
            monkeypatch.setattr(_lib, "context", lambda index: ctypes.c_void_p(1))
            monkeypatch.setattr(_lib, "stream_ptr", lambda device: ctypes.c_void_p(0))
            monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
            monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
            monkeypatch.setattr(torch.cuda, "device_count", lambda: world)
            monkeypatch.setattr(D.dist, "is_initialized", lambda: True)
            monkeypatch.setattr(D.dist, "get_world_size", lambda g=None: world)
            monkeypatch.setattr(D.dist, "get_rank", lambda g=None, rank=rank: rank)
            monkeypatch.setattr(D.dist, "get_backend", lambda g=None: "nccl")

            def all_gather_object(out, obj, group=None, rank=rank):
                assert obj == contributions[rank]
                group_log = fake.log[rank]
                group_log.append("all_gather_object")
                out[:] = contributions
            fake = group

            def barrier(group=None, rank=rank):
                fake.log[rank].append("barrier")

            def all_reduce(t, op=None, group=None, rank=rank):
                fake.log[rank].append("all_reduce")
                # MIN over the ranks: a rank votes 1 only if nothing failed anywhere it can see
                everyone_ok = not fail_alloc_on and not fail_open
                t.fill_(1 if everyone_ok else 0)
            monkeypatch.setattr(D.dist, "all_gather_object", all_gather_object)
            monkeypatch.setattr(D.dist, "barrier", barrier)
            monkeypatch.setattr(D.dist, "all_reduce", all_reduce)
            real_tensor = torch.tensor
            monkeypatch.setattr(torch, "tensor", lambda data, device=None, **k: real_tensor(data, **k))
            try:
                results[rank] = D.SharedGradientExchange.create(5, torch.device(device_type, rank) if device_type == "cuda"
                                                               else torch.device("cpu"))
            finally:
                monkeypatch.setattr(torch, "tensor", real_tensor)
        return results, libs, group
    return run


def test_successful_setup_fills_the_peer_table_and_counts_epochs(fake_world, monkeypatch):
    results, libs, group = fake_world(3)
    for rank, ex in results.items():
        assert ex is not None
        assert group.log[rank] == ["all_gather_object", "barrier", "all_reduce"]
        assert libs[rank].calls == ["alloc"] + ["open%d" % r for r in range(3) if r != rank]
        table = [ex.peers[r] for r in range(3)]
        assert table[rank] == 0x1000 * (rank + 1)                                  # own allocation at [rank]
        assert all(table[r] == 0x1000 * (r + 1) + 0x100000 * (rank + 1) for r in range(3) if r != rank)
    # steps: the library is told to count the steps on the device (epoch argument 0: nothing in the call changes
    # from step to step, so it replays from a CUDA graph) and sees the same table; the wrapper keeps its own count
    rank, ex = 1, results[1]
    monkeypatch.setattr(_lib, "load", lambda: libs[rank])
    for _ in range(3):
        out = ex.reduce(torch.zeros(2, 4, 4), torch.zeros(2, 5, 4))
        assert out.shape == (5, 3)
    assert libs[rank].epochs == [0, 0, 0] and ex.epoch == 3
    assert libs[rank].peers_seen == [ex.peers[r] for r in range(3)]
    with pytest.raises(ValueError, match="created for 5 vertices"):
        ex.reduce(torch.zeros(2, 4, 4), torch.zeros(2, 6, 4))
    assert ex.timed_out() is False
    monkeypatch.setattr(D.dist, "barrier", lambda group=None: fake_log.append("barrier"))
    fake_log = group.log[rank]
    ex.close()
    assert group.log[rank][-1] == "barrier"            # nobody frees a buffer while a peer may still store into it
    assert sorted(libs[rank].closed) == sorted(ex_ptr for r, ex_ptr in enumerate(libs[rank].peers_seen) if r != rank)
    assert libs[rank].freed == [0x1000 * (rank + 1)]
    ex.close()                                                                     # idempotent
    assert libs[rank].freed == [0x1000 * (rank + 1)]


@pytest.mark.parametrize("kwargs", [dict(fail_alloc_on=(1,)), dict(fail_open=(2, 0))])
def test_a_local_failure_makes_every_rank_fall_back_after_the_same_collectives(fake_world, kwargs):
    results, libs, group = fake_world(3, **kwargs)
    for rank in range(3):
        assert results[rank] is None
        # nobody skipped a collective: a rank that raised early would leave the others waiting
        assert group.log[rank] == ["all_gather_object", "barrier", "all_reduce"], (rank, group.log[rank])
        # whatever was mapped or allocated has been released
        opened = [c for c in libs[rank].calls if c.startswith("open")]
        failed_here = kwargs.get("fail_open") and kwargs["fail_open"][0] == rank
        if "fail_alloc_on" in kwargs:
            assert opened == []                                                    # nobody maps anything
            assert libs[rank].freed == ([] if rank == 1 else [0x1000 * (rank + 1)])
        else:
            assert len(libs[rank].closed) == len(opened) - (1 if failed_here else 0)
            assert libs[rank].freed == [0x1000 * (rank + 1)]


def test_cpu_groups_and_single_ranks_get_none(fake_world, monkeypatch):
    results, libs, group = fake_world(2, device_type="cpu")
    assert results == {0: None, 1: None}
    assert all(libs[r].calls == [] for r in range(2))                              # the library is not touched
    assert all(group.log[r] == ["all_reduce"] for r in range(2))                   # the vote still happens everywhere
    monkeypatch.setattr(D.dist, "is_initialized", lambda: False)
    assert D.SharedGradientExchange.create(5, torch.device("cpu")) is None
