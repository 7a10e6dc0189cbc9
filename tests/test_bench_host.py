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


def test_work_counter_matches_the_oracle_counters():
    """bench.count_raster_work (the instrumented pass behind the FP32 roofline) counts the same pixel visits and
    inside-test passes as the oracle's forward pass, triangle for triangle, when nothing is sampled."""
    import bench
    from oracle import oracle
    from pytorch_mesh_renderer_b200 import synthetic as S
    for sc in (S.sphere_views(40, 39, 1, 128), S.occlusion_soup(1, 192, n_triangles=150), S.cube_test_scene(160, 120)):
        cv, tr, W, H = sc["clip_vertices"][0], sc["triangles"], sc["width"], sc["height"]
        counters = oracle.forward(cv, tr, W, H, return_counters=True)[3]
        n_bbox, n_inside, stride = bench.count_raster_work(cv, tr, W, H)
        assert (n_bbox, n_inside, stride) == (int(counters[0]), int(counters[1]), 1)
    # sampled: an estimate, scaled back by the stride
    sc = S.sphere_views(60, 59, 1, 128)
    cv, tr = sc["clip_vertices"][0], sc["triangles"]
    exact = bench.count_raster_work(cv, tr, 128, 128)
    rough = bench.count_raster_work(cv, tr, 128, 128, max_tests=exact[0] // 4)
    assert rough[2] >= 4 and abs(rough[0] - exact[0]) < 0.2 * exact[0] and abs(rough[1] - exact[1]) < 0.25 * exact[1]
    assert bench.count_raster_work(cv, tr[:0], 128, 128) == (0, 0, 1)
