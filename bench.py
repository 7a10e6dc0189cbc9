"""bench.py -- fwd+bwd Mpixels/s of rasterize+interpolate (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2] [--impl b200|reference]

One "step" = one pass of the hot path over one batch: fused rasterize+interpolate forward
(ids, barycentrics, z, attribute image) and its backward (gradients to clip-space vertices and
attributes), A = 9 attributes.  The views of the sphere configs share one world-space mesh, so a step is:
world -> clip for the rank's views (one kernel), rasterize+interpolate forward, backward, reduction of the
per-view clip-space gradients to one world-space [V,3] gradient (one kernel) and, across ranks, the sum of
those partials (pushed into peer memory by that kernel, or ONE NCCL all-reduce).

N = 1 (default): BASELINE.json configs[1] (c2: 50 244-triangle UV sphere, 64 views x 512^2), plus a `configs`
array with the device-timed step of c1, c3, c4 and c5 at full batch (ms per step, both roofline fractions,
parity of the timed mode against the CPU oracle).
N > 1 (one process per GPU under torchrun): BASELINE.json configs[3] (c4: 256 views x 512^2 of ONE shared
99 904-triangle mesh), the 256 views batch-sharded 256/N per GPU: STRONG scaling; value = 256 views' pixels /
max-over-ranks time.  Rank 0 also times the whole of c4 alone first (`single_gpu`), so that the line carries its
own one-GPU reference for the same workload.  `--config c2|c3` under N > 1 keeps the per-GPU batch (weak scaling).

Prints ONE JSON line (rank 0).  --impl reference times the reference's own CPU implementation (its compiled
kernel from oracle/_ref under its own rasterize.py, one worker process per host core) instead.
"""
import argparse
import ctypes
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "fwd+bwd Mpixels/s (rasterize+interpolate)"
UNIT = "Mpixels/s"
A = 9

WORKLOADS = {
    "c1": "c1: 12-triangle cube test scene, batch 2 x 640x480, A=9",
    "c2": "c2: 50244-triangle UV sphere, batch 64 x 512x512, A=9",
    "c3": "c3: 1001112-triangle UV sphere, batch 16 x 1024x1024, A=9",
    "c4": "c4: 99904-triangle UV sphere shared by 256 views x 512x512, A=9",
    "c5": "c5: 2000 large overlapping triangles (depth complexity ~50), batch 32 x 2048x2048, A=9",
}


def claim_stdout():
    """stdout carries exactly ONE JSON line: keep a private handle on it and point fd 1 at stderr, so that
    whatever libraries print there (NCCL's version banner under torchrun, for one) cannot precede the line."""
    out = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    return out


def make_workload(name, batch_override=None):
    from pytorch_mesh_renderer_b200 import synthetic as S
    if name == "c1":
        return S.cube_test_scene()
    if name == "c2":
        return S.sphere_views(159, 158, batch_override or 64, 512)
    if name == "c3":
        return S.sphere_views(708, 707, batch_override or 16, 1024)
    if name == "c4":
        return S.sphere_views(224, 223, batch_override or 256, 512)
    if name == "c5":
        return S.occlusion_soup(batch_override or 32, 2048)
    raise ValueError(name)


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU while the timed region runs, in a separate
    nvidia-smi process (the recipe's clocks line) so that the Python launch loop cannot starve it."""

    QUERY = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index, period_ms=20):
        import subprocess
        self.proc = None
        if index is None:
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                 "-lms", str(period_ms)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def start(self):
        """Blocks until the first sample has arrived (nvidia-smi takes a while to come up), then marks the
        beginning of the timed region."""
        import select
        self.head = ""
        if self.proc is not None:
            ready, _, _ = select.select([self.proc.stdout], [], [], 3.0)
            if ready:
                self.head = self.proc.stdout.readline()
        self.t_begin = time.time()

    def stop(self):
        import datetime
        self.t_end = time.time()
        samples, reasons, max_mhz, nearby = [], set(), None, []
        if self.proc is not None:
            time.sleep(0.03)
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                out = ""
            for line in (self.head + out).splitlines():
                f = [x.strip() for x in line.split(",")]
                if len(f) < 7:
                    continue
                try:
                    stamp = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    mhz = float(f[1])
                    max_mhz = float(f[2])
                except ValueError:
                    continue
                inside = self.t_begin - 0.005 <= stamp <= self.t_end + 0.005
                if not inside:
                    if abs(stamp - self.t_begin) < 0.25 or abs(stamp - self.t_end) < 0.25:
                        nearby.append(mhz)
                    continue
                samples.append(mhz)
                for name, v in zip(self.NAMES, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        window = "timed region"
        if not samples and nearby:          # region shorter than the sampling period: nearest samples
            samples, window = nearby, "within 250 ms of the timed region"
        return {"sm_mhz": statistics.median(samples) if samples else None, "sm_max_mhz": max_mhz,
                "reasons": sorted(reasons), "samples": len(samples), "window": window}


def run_reference(args, rank):
    """--impl reference: the reference's CPU path on all host cores, bounded sample per step."""
    if rank != 0:
        return
    from oracle.cpu_reference import CpuReference
    sc = make_workload(args.config)
    ref = CpuReference(args.config)
    per_step = ref.cores
    for _ in range(max(args.warmup, 1)):
        ref.timed(per_step)
    px = secs = 0.0
    for _ in range(args.steps):
        p, s = ref.timed(per_step)
        px += p
        secs += s
    ref.close()
    value = px / secs / 1e6
    H, W = sc["height"], sc["width"]
    sample = "%d images of %dx%d per step (one per worker process), %d steps" % (per_step, W, H, args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.config], "attributes": A},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": ref.cores, "kind": ref.kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    args.out.write(json.dumps(line) + "\n")
    args.out.flush()


def bind_near_gpu(device_index):
    """Best effort: keep this process on the CPUs of the GPU's NUMA node, so that the pinned host buffers of the
    end-to-end leg are allocated in the memory next to the GPU's PCIe root (first touch).  Returns a description
    for the JSON line; never raises and changes nothing when the topology cannot be read."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return "gpu %s reports no NUMA node; affinity unchanged" % bdf
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        near = allowed & cpus
        if len(near) < 2:
            return "gpu %s on node %d, %d of its cpus allowed; affinity unchanged" % (bdf, node, len(near))
        if near != allowed:
            os.sched_setaffinity(0, near)
        return "gpu %s on node %d; process bound to %d of %d allowed cpus" % (bdf, node, len(near), len(allowed))
    except Exception as e:                      # noqa: BLE001 -- placement is an optimisation, never a failure
        return "affinity unchanged (%s)" % type(e).__name__


def count_raster_work(clip_vertices, triangles, width, height, max_tests=60_000_000):
    """Instrumented pass for the FP32 roofline: how many pixel visits (N_bbox, the clamped pixel boxes
    of K.cpp:367-375) and inside-test passes (N_inside, K.cpp:380) ONE image costs the reference algorithm.
    Host numpy over a strided sample of the triangles (at most `max_tests` pixel visits), scaled back; fp32
    arithmetic in the reference's operation order, but only counts leave this function."""
    import numpy as np
    v = np.asarray(clip_vertices, np.float32)
    t = np.asarray(triangles, np.int64)
    T = t.shape[0]
    if T == 0:
        return 0, 0, 1
    p = v[t]                                                      # [T,3,4]
    x, y, w = p[..., 0], p[..., 1], p[..., 3]
    alive = ~(w < 0).all(1)                                       # K.cpp:339
    front = (w > 0).all(1)
    half_w, half_h = np.float32(0.5 * width), np.float32(0.5 * height)
    with np.errstate(divide="ignore", invalid="ignore"):
        sx = (x / w + np.float32(1.0)) * half_w
        sy = (y / w + np.float32(1.0)) * half_h
    clampi = lambda a, hi: np.clip(np.nan_to_num(a, nan=0.0, posinf=1e9, neginf=-1e9), 0, hi).astype(np.int64)
    left = np.where(front, clampi(np.floor(sx.min(1)), width), 0)
    right = np.where(front, clampi(np.ceil(sx.max(1)), width), width)
    bottom = np.where(front, clampi(np.floor(sy.min(1)), height), 0)
    top = np.where(front, clampi(np.ceil(sy.max(1)), height), height)
    bw = np.where(alive, np.maximum(right - left, 0), 0)
    bh = np.where(alive, np.maximum(top - bottom, 0), 0)
    area = bw * bh
    total = int(area.sum())
    stride = max(1, -(-total // max_tests))
    pick = np.arange(0, T, stride)
    n_bbox = int(area[pick].sum())
    # unnormalised inverse, K.cpp:61-87 (rows are the edge functions), sign-flipped for negative determinants
    x0, x1, x2 = x[pick, 0], x[pick, 1], x[pick, 2]
    y0, y1, y2 = y[pick, 0], y[pick, 1], y[pick, 2]
    w0, w1, w2 = w[pick, 0], w[pick, 1], w[pick, 2]
    m = np.stack([y1 * w2 - w1 * y2, x2 * w1 - w2 * x1, x1 * y2 - y1 * x2,
                  y2 * w0 - w2 * y0, x0 * w2 - w0 * x2, x2 * y0 - y2 * x0,
                  y0 * w1 - w0 * y1, x1 * w0 - w1 * x0, x0 * y1 - y0 * x1], 1).astype(np.float32)
    det = x0 * m[:, 0] + x1 * m[:, 3] + x2 * m[:, 6]
    m = np.where((det < 0)[:, None], -m, m)
    cx = ((np.arange(width, dtype=np.float64) + 0.5) / float(half_w) - 1.0).astype(np.float32)   # K.cpp:376-377
    cy = ((np.arange(height, dtype=np.float64) + 0.5) / float(half_h) - 1.0).astype(np.float32)
    a_s, bw_s, l_s, b_s = area[pick], bw[pick], left[pick], bottom[pick]
    n_inside = 0
    chunk_at, k = 0, len(pick)
    while chunk_at < k:
        stop = chunk_at + 1
        tests = int(a_s[chunk_at])
        while stop < k and tests + int(a_s[stop]) <= 4_000_000:
            tests += int(a_s[stop])
            stop += 1
        if tests:
            sel = np.arange(chunk_at, stop)
            reps = a_s[sel]
            tri = np.repeat(sel, reps)
            off = np.arange(tests) - np.repeat(np.cumsum(reps) - reps, reps)
            px = cx[np.minimum(l_s[tri] + off % np.maximum(bw_s[tri], 1), width - 1)]
            py = cy[np.minimum(b_s[tri] + off // np.maximum(bw_s[tri], 1), height - 1)]
            mm = m[tri]
            e0 = mm[:, 0] * px + mm[:, 1] * py + mm[:, 2]
            e1 = mm[:, 3] * px + mm[:, 4] * py + mm[:, 5]
            e2 = mm[:, 6] * px + mm[:, 7] * py + mm[:, 8]
            n_inside += int(((np.minimum(np.minimum(e0, e1), e2) >= 0) & (np.maximum(np.maximum(e0, e1), e2) > 0)).sum())
        chunk_at = stop
    return n_bbox * stride, n_inside * stride, stride


def parity_report(sc, g_host, mode, device, image=0):
    """The timed mode against the north-star tolerance on one image of the workload: the CUDA path's forward
    buffers must equal the CPU oracle's bit for bit; its gradients are reported as the fraction of entries outside
    1e-6 + 1e-5*|ref| against the reference's summation order and against the exactly (fp64) summed fp32 terms,
    next to the reference's own fraction against that yardstick (an fp32 sum in ANY order, the
    reference's included, leaves the tolerance on large-triangle configs).  The oracle is the checker here."""
    import numpy as np
    import torch
    import pytorch_mesh_renderer_b200 as pmr
    from oracle import oracle as cpu_oracle
    H, W = sc["height"], sc["width"]
    sl = slice(image, image + 1)
    cv_h, at_h = sc["clip_vertices"][sl], sc["attributes"][sl]
    g = np.ascontiguousarray(g_host[sl])
    t0 = time.perf_counter()
    ref = cpu_oracle.rasterize_clip_space(cv_h, at_h, sc["triangles"], W, H, sc["background"], grad_out=g,
                                          f64_yardstick=True)
    oracle_s = time.perf_counter() - t0
    to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    cv, at = to_dev(cv_h).requires_grad_(True), to_dev(at_h).requires_grad_(True)
    with pmr.backward_mode(mode):
        out, (ids, bary, z) = pmr.rasterize_clip_space(cv, at, to_dev(sc["triangles"]), W, H, to_dev(sc["background"]),
                                                        return_buffers=True)
        out.backward(to_dev(g))
    torch.cuda.synchronize()
    same = lambda a, b: bool(np.array_equal(a.detach().cpu().numpy(), b))
    return {"mode": mode, "image": image, "tolerance": "1e-6 + 1e-5*|ref| (north star), per entry",
            "ids_bit_exact": same(ids, ref["ids"]), "barycentrics_bit_exact": same(bary, ref["bary"]),
            "z_bit_exact": same(z, ref["z"]), "image_bit_exact": same(out, ref["out"]),
            "d_vertices": cpu_oracle.tolerance_report(cv.grad.cpu().numpy(), ref["d_vertices"], ref["d_vertices_f64"]),
            "d_attributes": cpu_oracle.tolerance_report(at.grad.cpu().numpy(), ref["d_attributes"], ref["d_attributes_f64"]),
            "oracle_seconds": oracle_s}


def algorithmic_bytes(P, B, V, T):
    """SURVEY 8d / BASELINE.md 4: forward writes ids 4 + bary 12 + z 4 + image 4A per pixel and reads each
    image's vertices, attributes and the index buffer once; backward reads grad 4A + ids 4 + bary 12 per pixel
    and writes each image's vertex / attribute gradients once."""
    fwd = P * (20 + 4 * A) + B * (V * (16 + 4 * A) + 12 * T)
    bwd = P * (16 + 4 * A) + B * V * (16 + 4 * A)
    per_kernel = {"scatter": B * (16 * V + 12 * T), "resolve": P * (20 + 4 * A) + B * V * 4 * A,
                  "raster": fwd, "backward": bwd}
    return fwd, bwd, per_kernel


KERNEL_NAMES = {"scatter": "scatter_small_kernel (small triangles -> packed depth keys)",
                "resolve": "resolve kernel (depth keys -> ids/bary/z + fused interpolation)",
                "raster": "raster_tile_kernel (tile path)",
                "backward": "backward kernel (fused interpolation backward, %s mode)"}


def measure_on_device(sc, device, mode, steps, warmup, rank=0, world=1, collective="peer", sample_clocks=True,
                      grad_seed=1, use_graph=True):
    """Device-timed steps of one workload on this rank's slice of it.  Returns a dict with ms_per_step (max
    over ranks), per-stage times from the library's own CUDA-event pairs, launch count, clocks and sizes."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import pytorch_mesh_renderer_b200 as pmr
    from pytorch_mesh_renderer_b200 import _lib
    from pytorch_mesh_renderer_b200 import distributed as D
    from pytorch_mesh_renderer_b200.camera_utils import transform_shared_mesh
    local = device.index
    shared_mesh = "world_vertices" in sc and sc["world_vertices"].ndim == 2
    B, V = sc["clip_vertices"].shape[:2]
    T = sc["triangles"].shape[0]
    H, W = sc["height"], sc["width"]
    to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    clip, attrs, tris, bg = (to_dev(sc[k]) for k in ("clip_vertices", "attributes", "triangles", "background"))
    gen = torch.Generator(device=device)
    gen.manual_seed(grad_seed + rank)
    grad = torch.randn((B, H, W, A), generator=gen, device=device, dtype=torch.float32)
    exchange = None
    if shared_mesh:
        mvp = to_dev(sc["camera_matrices"])
        world_vertices = to_dev(sc["world_vertices"])
        if world > 1 and collective == "peer":
            # partial sums go straight into the peers' memory from the vertex-stage backward kernel; None (all
            # ranks agree) when peer mapping is unavailable -> NCCL all-reduce
            exchange = D.SharedGradientExchange.create(V, device)

    def step():
        at = attrs.detach().requires_grad_(True)
        if shared_mesh:
            # multi-view fitting: the world-space mesh is the shared parameter; d(clip) flows back through the
            # view matrices and the broadcast, then the partials of the ranks are summed
            wv = world_vertices.detach().requires_grad_(True)
            cv = transform_shared_mesh(mvp, wv, exchange=exchange)   # backward sums over the local views
            out = pmr.rasterize_clip_space(cv, at, tris, W, H, bg)
            out.backward(grad)
            if exchange is None and world > 1:
                D.all_reduce_gradients([wv.grad])
            return wv.grad
        cv = clip.detach().requires_grad_(True)
        out = pmr.rasterize_clip_space(cv, at, tris, W, H, bg)
        out.backward(grad)
        return cv.grad

    exchange_check = None
    with pmr.backward_mode(mode):
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        if exchange is not None:
            # The peer-memory exchange against the NCCL all-reduce on this very workload (the multi-GPU test of
            # tests/test_gpu_peer_exchange.py, run where several GPUs exist): every rank must hold the SAME bits
            # (the partials are added in rank order everywhere), and they must equal the NCCL sum within rounding.
            at = attrs.detach().requires_grad_(True)
            wv = world_vertices.detach().requires_grad_(True)
            pmr.rasterize_clip_space(transform_shared_mesh(mvp, wv, exchange=exchange), at, tris, W, H, bg).backward(grad)
            via_peers = wv.grad.clone()
            wv2 = world_vertices.detach().requires_grad_(True)
            pmr.rasterize_clip_space(transform_shared_mesh(mvp, wv2, exchange=None), at.detach().requires_grad_(True),
                                     tris, W, H, bg).backward(grad)
            via_nccl = wv2.grad.clone()
            dist.all_reduce(via_nccl, op=dist.ReduceOp.SUM)
            ref_bits = via_peers.clone()
            dist.broadcast(ref_bits, src=0)
            same_bits = torch.tensor([int(torch.equal(ref_bits, via_peers))], device=device)
            dist.all_reduce(same_bits, op=dist.ReduceOp.MIN)
            scale = float(via_nccl.abs().max().item()) + 1e-30
            err = float((via_peers - via_nccl).abs().max().item())
            # (the atomic backward's own run-to-run rounding is in both numbers)
            exchange_check = {"all_ranks_bit_identical": bool(same_bits.item()), "max_abs_diff_vs_nccl": err,
                              "max_abs_value": scale, "finite": bool(torch.isfinite(via_peers).all().item())}
            if not exchange_check["all_ranks_bit_identical"] or not exchange_check["finite"] or err > 1e-3 * scale:
                raise RuntimeError("peer exchange check failed: %s" % (exchange_check,))
        # Per-stage times and the launch count come from a short eager pass (the library's own CUDA-event pairs);
        # the timed steps then replay ONE CUDA graph of the whole step -- the same kernels, without the host's
        # launch gaps between them (about a dozen launches of 0.02 - 0.6 ms each per step).  All ranks take the
        # same decision.
        _lib.enable_stage_timing(local, True)
        _lib.read_stage_timing(local, reset=True)
        launches0 = _lib.launch_count(local)
        stage_steps = max(3, min(steps, 10))
        for _ in range(stage_steps):
            step()
        torch.cuda.synchronize()
        launches_per_step = (_lib.launch_count(local) - launches0) / stage_steps
        stages = _lib.read_stage_timing(local, reset=True)
        _lib.enable_stage_timing(local, False)
        graph, graph_note = None, "eager launches"
        if use_graph:
            try:
                side = torch.cuda.Stream(device)
                side.wait_stream(torch.cuda.current_stream(device))
                with torch.cuda.stream(side):
                    for _ in range(2):
                        step()
                torch.cuda.current_stream(device).wait_stream(side)
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    captured = step()
                graph.replay()
                torch.cuda.synchronize()
                if not bool(torch.isfinite(captured).all().item()):
                    raise RuntimeError("graph replay produced non-finite gradients")
                graph_note = "one CUDA graph of the step, replayed"
            except Exception as e:                       # noqa: BLE001 -- fall back to eager launches, say so
                graph, graph_note = None, "eager launches (graph capture failed: %s: %s)" % (type(e).__name__, str(e)[:200])
                torch.cuda.synchronize()
        if world > 1:                                    # everybody or nobody
            ok = torch.tensor([1 if (graph is not None or not use_graph) else 0], device=device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if use_graph and int(ok.item()) == 0:
                graph, graph_note = None, "eager launches (a rank could not capture the step)"
        run = graph.replay if graph is not None else step
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        sampler = ClockSampler(local if (rank == 0 and sample_clocks) else None)   # one sampler per job
        sampler.start()                                              # returns once samples are flowing
        if world > 1:
            dist.barrier()
        sampler.t_begin = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            run()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms_total = e0.elapsed_time(e1)
        launches = int(round(launches_per_step * steps))
        clocks = sampler.stop()
        del graph
    if world > 1:
        t = torch.tensor([ms_total], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    timed_out = False
    if exchange is not None:
        timed_out = exchange.timed_out()
        exchange.close()
    if timed_out:
        raise RuntimeError("peer exchange: a rank stopped waiting for a peer's partial sums; the result is invalid")
    del grad
    return {"ms_per_step": ms_total / steps, "stages": stages, "steps": steps, "stage_steps": stage_steps,
            "launches": int(launches), "launch_mode": graph_note,
            "clocks": clocks, "B": B, "V": V, "T": T, "H": H, "W": W, "shared_mesh": shared_mesh,
            "exchange_check": exchange_check,
            "exchange": "peer" if (shared_mesh and world > 1 and collective == "peer" and not timed_out and exchange is not None)
                        else ("nccl" if (shared_mesh and world > 1) else "none"),
            "tensors": (clip, attrs, tris, bg)}


def roofline_of(m, mode, total_pixels_job=None):
    """The roofline block of the JSON line from one measure_on_device() result."""
    peak, peak_src = measured_peak_gbs()
    B, V, T, H, W = m["B"], m["V"], m["T"], m["H"], m["W"]
    P = B * H * W
    bytes_fwd, bytes_bwd, kernel_bytes = algorithmic_bytes(P, B, V, T)
    stages, steps, ms_step = m["stages"], m["stage_steps"], m["ms_per_step"]
    stage_ms = {k: (v[0] / max(v[1], 1)) for k, v in stages.items()}
    per_step = {k: v[0] / steps for k, v in stages.items()}
    dominant = max(kernel_bytes, key=lambda k: per_step[k])
    dom_bytes = kernel_bytes[dominant]
    achieved = dom_bytes / (stage_ms[dominant] * 1e-3) / 1e9 if stage_ms[dominant] > 0 else 0.0
    fwd_ms = per_step["bin"] + per_step["scatter"] + per_step["raster"] + per_step["resolve"]
    name = KERNEL_NAMES[dominant]
    return {
        "bound": "hbm", "kernel": name % mode if "%s" in name else name,
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
        "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes, "ms_per_launch": stage_ms[dominant],
        "stages_ms_per_step": per_step,
        "forward": {"algorithmic_bytes": bytes_fwd, "ms": fwd_ms,
                    "frac": bytes_fwd / (fwd_ms * 1e-3) / 1e9 / peak if fwd_ms > 0 else None},
        "backward": {"algorithmic_bytes": bytes_bwd, "ms": per_step["backward"],
                     "frac": bytes_bwd / (per_step["backward"] * 1e-3) / 1e9 / peak if per_step["backward"] > 0 else None},
        "step": {"algorithmic_bytes": bytes_fwd + bytes_bwd,
                 "achieved": (bytes_fwd + bytes_bwd) / (ms_step * 1e-3) / 1e9,
                 "frac": (bytes_fwd + bytes_bwd) / (ms_step * 1e-3) / 1e9 / peak},
    }


def fp32_roofline(sc, m, roofline, clocks, images=(0,)):
    """The other roofline: FP32 pipe, F = 12 N_bbox + 20 N_inside + (230 + 18 A) N_cov per step,
    against 148 SMs x 128 lanes x SM clock non-FMA instructions per second (the parity contract forbids FMA).
    N_bbox / N_inside come from an instrumented host pass over `images` (scaled by the batch), N_cov from the
    device buffers of one forward call.  Diagnostic: a failure here is reported, not raised."""
    import torch
    import pytorch_mesh_renderer_b200 as pmr
    try:
        clip, attrs, tris, bg = m["tensors"]
        B, H, W = m["B"], m["H"], m["W"]
        P = B * H * W
        n_bbox = n_inside = 0
        strides = []
        for i in images:
            nb, ni, stride = count_raster_work(sc["clip_vertices"][i], sc["triangles"], W, H,
                                               max_tests=60_000_000 // len(images))
            n_bbox += nb
            n_inside += ni
            strides.append(stride)
        n_bbox /= len(images)
        n_inside /= len(images)
        with torch.no_grad():
            _, (_, bary_dev, _) = pmr.rasterize_clip_space(clip, attrs, tris, W, H, bg, return_buffers=True)
            n_cov = int((bary_dev.sum(dim=3) > 0.5).sum().item())
            del bary_dev
        flops = 12.0 * n_bbox * B + 20.0 * n_inside * B + (230.0 + 18.0 * A) * n_cov
        fp32_peak = 148 * 128 * float(clocks.get("sm_max_mhz") or 1965.0) * 1e6
        fp32_floor_ms = flops / fp32_peak * 1e3
        bytes_total = roofline["step"]["algorithmic_bytes"]
        hbm_floor_ms = bytes_total / (roofline["peak"] * 1e9) * 1e3
        ms_step = m["ms_per_step"]
        roofline["fp32"] = {
            "flops_per_step": flops, "n_bbox_per_image": n_bbox, "n_inside_per_image": n_inside,
            "n_covered_pixels": n_cov, "depth_complexity": n_inside * B / float(P),
            "sample": "images %s, every %s triangle, scaled by the batch" % (list(images), strides),
            "peak": fp32_peak / 1e12, "unit": "T non-FMA fp32 instructions/s",
            "achieved": flops / (ms_step * 1e-3) / 1e12, "frac": fp32_floor_ms / ms_step, "floor_ms": fp32_floor_ms}
        roofline["hbm_floor_ms"] = hbm_floor_ms
        roofline["slower_roofline"] = "fp32" if fp32_floor_ms > hbm_floor_ms else "hbm"
        roofline["frac_of_slower_roofline"] = max(fp32_floor_ms, hbm_floor_ms) / ms_step
    except Exception as e:                       # noqa: BLE001
        roofline["fp32"] = {"error": "%s: %s" % (type(e).__name__, e)}


def host_entry_point_e2e(sc, device, mode, steps, world, placement):
    """e2e: the C-ABI host entry point, pinned host buffers, copies inside the timed region."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from pytorch_mesh_renderer_b200 import _lib
    L = _lib.load()
    local = device.index
    ctx = _lib.context(local)
    B, V = sc["clip_vertices"].shape[:2]
    T = sc["triangles"].shape[0]
    H, W = sc["height"], sc["width"]
    P = B * H * W
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_v, h_a, h_t, h_bg = pin(sc["clip_vertices"]), pin(sc["attributes"]), pin(sc["triangles"]), pin(sc["background"])
    h_g = torch.empty((B, H, W, A), dtype=torch.float32).pin_memory()
    h_g.normal_(generator=torch.Generator().manual_seed(7))
    h_out = torch.empty((B, H, W, A), dtype=torch.float32).pin_memory()
    h_dv = torch.empty((B, V, 4), dtype=torch.float32).pin_memory()
    h_da = torch.empty((B, V, A), dtype=torch.float32).pin_memory()
    p = lambda t_: ctypes.c_void_p(t_.data_ptr())
    mode_code = _lib.BACKWARD_ATOMIC if mode == "atomic" else _lib.BACKWARD_ORDERED
    stream = torch.cuda.current_stream(device)

    def host_step():
        rc = L.pmr_rasterize_clip_space_host(ctx, p(h_v), p(h_a), p(h_t), p(h_bg), p(h_g), B, V, T, A, W, H,
                                             p(h_out), p(h_dv), p(h_da), None, None, None, mode_code,
                                             ctypes.c_void_p(stream.cuda_stream))
        _lib.check(ctx, rc)

    e2e_steps = max(3, min(steps, 10))
    for _ in range(2):
        host_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_step()
    torch.cuda.synchronize()
    secs = time.perf_counter() - t0
    mine = secs
    if world > 1:
        t = torch.tensor([secs], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        secs = float(t.item())
    h2d = h_v.numel() * 4 + h_a.numel() * 4 + h_t.numel() * 4 + h_bg.numel() * 4 + h_g.numel() * 4
    d2h = h_out.numel() * 4 + h_dv.numel() * 4 + h_da.numel() * 4
    gbs = (h2d + d2h) * e2e_steps / mine / 1e9
    per_rank = [gbs]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, gbs)
        per_rank = gathered
    # The ceiling of the host link for this job: plain pinned copies, up and down at once, on all ranks at the
    # same time (what the call above cannot beat: it moves its bytes through the same link).
    n_probe = 256 << 20
    up_h = torch.empty(n_probe, dtype=torch.uint8).pin_memory()
    down_h = torch.empty(n_probe, dtype=torch.uint8).pin_memory()
    up_d = torch.empty(n_probe, dtype=torch.uint8, device=device)
    down_d = torch.zeros(n_probe, dtype=torch.uint8, device=device)
    s_up, s_down = torch.cuda.Stream(device), torch.cuda.Stream(device)

    def probe(reps):
        for _ in range(reps):
            with torch.cuda.stream(s_up):
                up_d.copy_(up_h, non_blocking=True)
            with torch.cuda.stream(s_down):
                down_h.copy_(down_d, non_blocking=True)
        s_up.synchronize()
        s_down.synchronize()

    probe(1)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    probe(6)
    ceiling = 2 * n_probe * 6 / (time.perf_counter() - t0) / 1e9
    ceilings = [ceiling]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, ceiling)
        ceilings = gathered
    del up_h, down_h, up_d, down_d
    total_pixels = P * e2e_steps
    if world > 1:
        t = torch.tensor([float(total_pixels)], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        total_pixels = float(t.item())
    return {"value": total_pixels / secs / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": 1e3 * secs / e2e_steps,
            "call": "pmr_rasterize_clip_space_host (C ABI, pinned host buffers in, host buffers out)",
            "host_placement": placement, "host_link_gbs_per_rank": [round(x, 1) for x in per_rank],
            "host_link_ceiling_gbs_per_rank": [round(x, 1) for x in ceilings],
            "host_link_note": "GB/s moved over the host link by each rank, both directions together: by the timed call, "
                              "and by plain pinned copies (256 MiB up + 256 MiB down at once, all ranks concurrently)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default=None, choices=sorted(WORKLOADS),
                    help="default: c2 on one GPU, c4 batch-sharded (strong scaling) on several")
    ap.add_argument("--batch", type=int, default=None, help="override the per-GPU batch (diagnostics only)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="atomic", choices=["atomic", "ordered"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the c1 / c3 / c4 / c5 lines of the default N = 1 run")
    ap.add_argument("--no-graph", action="store_true", help="launch every step from the host instead of replaying a CUDA graph")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"],
                    help="shared-mesh gradient over the ranks: fused push + reduce over peer memory, or kernel + NCCL all-reduce")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    args.out = claim_stdout()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    default_run = args.config is None and args.batch is None
    if args.config is None:
        args.config = "c2" if world == 1 else "c4"

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import pytorch_mesh_renderer_b200 as pmr           # noqa: F401
    from pytorch_mesh_renderer_b200 import distributed as D
    from pytorch_mesh_renderer_b200 import synthetic as S

    # CPU baseline first (rank 0, N = 1 only): worker processes are spawned before CUDA is touched.
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.cpu_reference import CpuReference
        ref = CpuReference(args.config)
        ref.warm()
        n_img = {"c1": 2 * ref.cores, "c2": 4 * ref.cores, "c3": ref.cores, "c4": 4 * ref.cores, "c5": ref.cores}[args.config]
        px, secs = ref.timed(n_img)
        ref.close()
        cpu_baseline = {"value": px / secs / 1e6, "unit": UNIT, "cores": ref.cores, "kind": ref.kind,
                        "sample": "%d images of the workload, one worker process per core, %.1f s wall" % (n_img, secs)}

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    placement = bind_near_gpu(local_rank)      # after the CPU baseline, which uses every core
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    # ---- the workload of this rank
    strong = world > 1 and args.config == "c4" and args.batch is None
    single_gpu = None
    if strong:
        # 256 views of one mesh, batch-sharded: view i belongs to rank i mod N
        full = make_workload("c4")
        if rank == 0:
            # the same job on one GPU, timed inside this run: the one-GPU reference of the strong-scaling line
            m1 = measure_on_device(full, device, args.mode, max(3, min(args.steps, 10)), 3, sample_clocks=False,
                                   use_graph=not args.no_graph)
            single_gpu = {"ms_per_step": m1["ms_per_step"], "views": m1["B"],
                          "value": m1["B"] * m1["H"] * m1["W"] / (m1["ms_per_step"] * 1e-3) / 1e6, "unit": UNIT,
                          "stages_ms_per_step": {k: v[0] / m1["stage_steps"] for k, v in m1["stages"].items()}}
            del m1
            torch.cuda.empty_cache()
        # views dealt round-robin: neighbouring views of the orbit cost alike, a step is as long as the slowest rank
        mine = D.shard_views(full["clip_vertices"].shape[0], rank, world, interleaved=True)
        sc = dict(full)
        for key in ("clip_vertices", "attributes", "camera_matrices"):
            sc[key] = np.ascontiguousarray(full[key][mine.start:mine.stop:mine.step])
        del full
        dist.barrier()
    else:
        sc = make_workload(args.config, args.batch)
        if world > 1 and "world_vertices" in sc and sc["world_vertices"].ndim == 2:
            # weak scaling: the job renders world * B views of ONE mesh; this rank owns a contiguous slice
            B_local = sc["clip_vertices"].shape[0]
            mine = D.shard_views(B_local * world, rank, world)
            sc["camera_matrices"] = S.orbit_cameras(B_local * world)[mine.start:mine.stop]
            sc["clip_vertices"] = S.transform(sc["camera_matrices"], sc["world_vertices"])

    m = measure_on_device(sc, device, args.mode, args.steps, args.warmup, rank, world, args.collective,
                          use_graph=not args.no_graph)
    B, V, T, H, W = m["B"], m["V"], m["T"], m["H"], m["W"]
    P = B * H * W
    ms_step = m["ms_per_step"]
    job_pixels = float(P)
    if world > 1:
        t = torch.tensor([job_pixels], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        job_pixels = float(t.item())
    value = job_pixels / (ms_step * 1e-3) / 1e6
    roofline = roofline_of(m, args.mode)
    if rank == 0:
        fp32_roofline(sc, m, roofline, m["clocks"], images=tuple(range(0, B, max(1, B // 4)))[:4])
        try:     # DRAM bytes per launch of the dominant kernel from this round's committed ncu capture
            with open(os.path.join(ROOT, "profiles", "r02", "dram_traffic.json")) as f:
                t_ = json.load(f).get(args.config, {}).get(roofline["kernel"].split(" ")[0])
            if t_ and args.batch is None and world == 1:
                roofline["traffic"] = int((t_["dram_read_mb"] + t_["dram_write_mb"]) * 1e6)
                roofline["traffic_source"] = t_.get("source")
        except Exception:
            pass

    e2e = None
    if not args.no_e2e:
        e2e = host_entry_point_e2e(sc, device, args.mode, args.steps, world, placement)

    parity = None
    if rank == 0 and not args.no_parity:
        try:
            g0 = np.random.default_rng(11).standard_normal((1, H, W, A), dtype=np.float32)
            parity = parity_report(sc, g0, args.mode, device)
        except Exception as e:                   # noqa: BLE001 -- reported, not raised: the timing stands on its own
            parity = {"error": "%s: %s" % (type(e).__name__, e)}
    m.pop("tensors", None)

    # ---- the other configs at full batch (default N = 1 run only): device-timed step, rooflines, parity
    configs = None
    if rank == 0 and world == 1 and default_run and not args.no_configs:
        configs = []
        torch.cuda.empty_cache()
        for name in ("c1", "c3", "c4", "c5"):
            try:
                sc_k = make_workload(name)
                m_k = measure_on_device(sc_k, device, args.mode, 10 if name != "c5" else 5, 3, sample_clocks=False,
                                        use_graph=not args.no_graph)
                r_k = roofline_of(m_k, args.mode)
                n_img = m_k["B"]
                fp32_roofline(sc_k, m_k, r_k, m["clocks"], images=tuple(range(0, n_img, max(1, n_img // 4)))[:4])
                g0 = np.random.default_rng(11).standard_normal((1, m_k["H"], m_k["W"], A), dtype=np.float32)
                par_k = parity_report(sc_k, g0, args.mode, device)
                P_k = m_k["B"] * m_k["H"] * m_k["W"]
                configs.append({
                    "workload": WORKLOADS[name], "batch": m_k["B"], "triangles": m_k["T"], "vertices": m_k["V"],
                    "image": [m_k["H"], m_k["W"]], "steps": m_k["steps"], "ms_per_step": m_k["ms_per_step"],
                    "launches": m_k["launch_mode"],
                    "value": P_k / (m_k["ms_per_step"] * 1e-3) / 1e6, "unit": UNIT,
                    "stages_ms_per_step": r_k["stages_ms_per_step"],
                    "hbm_frac_step": r_k["step"]["frac"], "hbm_frac_forward": r_k["forward"]["frac"],
                    "hbm_frac_backward": r_k["backward"]["frac"],
                    "fp32_frac_step": (r_k.get("fp32") or {}).get("frac"),
                    "slower_roofline": r_k.get("slower_roofline"),
                    "frac_of_slower_roofline": r_k.get("frac_of_slower_roofline"),
                    "depth_complexity": (r_k.get("fp32") or {}).get("depth_complexity"), "parity": par_k})
                del m_k, sc_k
                torch.cuda.empty_cache()
            except Exception as e:               # noqa: BLE001
                configs.append({"workload": WORKLOADS[name], "error": "%s: %s" % (type(e).__name__, e)})

    if rank == 0:
        exchange_text = {"none": "none",
                         "peer": "world-space vertex gradient [V,3] per step: pushed into peer memory by the vertex-stage backward kernel, summed in rank order (no NCCL call)",
                         "nccl": "NCCL all_reduce(sum) of the world-space vertex gradient [V,3] per step"}[m["exchange"]]
        bytes_total = roofline["step"]["algorithmic_bytes"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.config], "per_gpu_batch": B, "job_batch": int(round(job_pixels / (H * W))),
                       "triangles": T, "vertices": V,
                       "image": [H, W], "attributes": A, "backward_mode": args.mode,
                       "triangles_per_s": job_pixels / (H * W) * T / (ms_step * 1e-3),
                       "l2": "per-step working set %.2f GB exceeds the 126 MB L2; no explicit flush" % (bytes_total / 1e9),
                       "vertex_stage": "world->clip kernel + view-summed backward inside the step" if m["shared_mesh"] else "none",
                       "launches": m["launch_mode"] + "; per-stage times from %d eager steps before the timed ones" % m["stage_steps"],
                       "collective": exchange_text,
                       "sharding": ("views dealt round-robin (view i on rank i mod N): neighbouring views of the orbit "
                                    "cost alike, a step is as long as the slowest rank") if strong else "none"},
            "roofline": roofline, "parity": parity, "cpu_baseline": cpu_baseline, "e2e": e2e,
            "gpu_launches": m["launches"], "clocks": m["clocks"],
        }
        if single_gpu is not None:
            line["single_gpu"] = single_gpu
        if m.get("exchange_check") is not None:
            line["exchange_check"] = m["exchange_check"]
        if configs is not None:
            line["configs"] = configs
        args.out.write(json.dumps(line) + "\n")
        args.out.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
