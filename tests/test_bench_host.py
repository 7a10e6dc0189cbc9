"""Host-side pieces of bench.py that run without a GPU: the NUMA placement of the end-to-end leg is best effort and
must never raise or shrink the process to nothing."""
import builtins
import io
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _fake_torch(monkeypatch, domain=0, bus=0x1b, dev=0):
    import torch
    props = types.SimpleNamespace(pci_domain_id=domain, pci_bus_id=bus, pci_device_id=dev)
    monkeypatch.setattr(torch.cuda, "get_device_properties", lambda i: props)


def _fake_sysfs(monkeypatch, numa_node, cpulist):
    real_open = builtins.open

    def fake_open(path, *a, **k):
        if isinstance(path, str) and path.endswith("/numa_node"):
            assert path == "/sys/bus/pci/devices/0000:1b:00.0/numa_node"
            return io.StringIO("%d\n" % numa_node)
        if isinstance(path, str) and path.startswith("/sys/devices/system/node/node"):
            return io.StringIO(cpulist + "\n")
        return real_open(path, *a, **k)
    monkeypatch.setattr(builtins, "open", fake_open)


def test_placement_binds_to_the_gpu_node(monkeypatch):
    import bench
    allowed = set(range(8))
    calls = []
    _fake_torch(monkeypatch)
    _fake_sysfs(monkeypatch, 1, "4-7,64-71")
    monkeypatch.setattr(os, "sched_getaffinity", lambda pid: set(allowed))
    monkeypatch.setattr(os, "sched_setaffinity", lambda pid, cpus: calls.append(set(cpus)))
    text = bench.bind_near_gpu(0)
    assert calls == [{4, 5, 6, 7}] and "node 1" in text and "4 of 8" in text


def test_placement_leaves_things_alone(monkeypatch):
    import bench
    calls = []
    _fake_torch(monkeypatch)
    monkeypatch.setattr(os, "sched_getaffinity", lambda pid: set(range(8)))
    monkeypatch.setattr(os, "sched_setaffinity", lambda pid, cpus: calls.append(set(cpus)))
    _fake_sysfs(monkeypatch, -1, "0-7")                       # virtual machine: no NUMA node reported
    assert "no NUMA node" in bench.bind_near_gpu(0)
    _fake_sysfs(monkeypatch, 0, "0-7")                        # already on the node
    assert "bound to 8 of 8" in bench.bind_near_gpu(0)
    _fake_sysfs(monkeypatch, 1, "7,100-120")                  # one usable cpu: not worth it
    assert "unchanged" in bench.bind_near_gpu(0)
    assert calls == []


def test_placement_never_raises(monkeypatch):
    import bench
    import torch

    def boom(i):
        raise RuntimeError("no driver")
    monkeypatch.setattr(torch.cuda, "get_device_properties", boom)
    assert bench.bind_near_gpu(0).startswith("affinity unchanged")
