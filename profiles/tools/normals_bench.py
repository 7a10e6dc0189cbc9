"""Times the vertex-normal kernels (csrc/mesh_normals.cu) on the c2 / c3 meshes and prints achieved GB/s against
the measured HBM peak, next to the reference's index_add_ formulation run with torch ops on the same GPU.
Algorithmic bytes: forward reads B*V*12 + T*12, writes B*V*12 (+ B*V*12 raw sums for the backward);
backward reads grad + raw + vertices (B*V*36) + T*12 and writes B*V*12.

    python profiles/tools/normals_bench.py"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pytorch_mesh_renderer_b200 import ops, synthetic  # noqa: E402


def torch_ops_normals(vertices, triangles):
    """meshes.py:19-34 of the reference, on the GPU with torch ops (explicit dim)."""
    tri = triangles.long()
    normals = torch.zeros_like(vertices)
    for b in range(vertices.shape[0]):
        vf = vertices[b, tri, :]
        for c in range(3):
            normals[b].index_add_(0, tri[:, c], torch.cross(vf[:, (c + 1) % 3] - vf[:, c],
                                                            vf[:, (c + 2) % 3] - vf[:, c], dim=-1))
    return torch.nn.functional.normalize(normals, eps=1e-6, p=2, dim=-1)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    for name, (n_lon, n_rings), B in (("c2 sphere", (159, 158), 64), ("c3 sphere", (708, 707), 16)):
        verts, tris = synthetic.uv_sphere(n_lon, n_rings)
        V, T = verts.shape[0], tris.shape[0]
        rng = np.random.default_rng(0)
        v = torch.from_numpy((verts[None] + 0.001 * rng.standard_normal((B, V, 3))).astype(np.float32)).cuda()
        t = torch.from_numpy(tris).cuda()
        g = torch.randn_like(v)
        table_ms = timed(lambda: ops.vertex_incidence(t, V), reps=5)
        offsets, incidence = ops.vertex_incidence(t, V)
        normals, raw = ops.vertex_normals_forward(v, t, offsets, incidence)
        fwd_ms = timed(lambda: ops.vertex_normals_forward(v, t, offsets, incidence))
        bwd_ms = timed(lambda: ops.vertex_normals_backward(g, raw, v, t, offsets, incidence))
        ref_fwd_ms = timed(lambda: torch_ops_normals(v, t), reps=3)
        vr = v.clone().requires_grad_(True)

        def ref_step():
            vr.grad = None
            torch_ops_normals(vr, t).backward(g)
        ref_step_ms = timed(ref_step, reps=3)
        fwd_bytes = B * V * 36 + T * 12 + 3 * T * 4
        bwd_bytes = B * V * 60 + T * 12 + 3 * T * 4
        print(json.dumps({"mesh": name, "B": B, "V": V, "T": T, "incidence_table_ms": table_ms,
                          "forward_ms": fwd_ms, "forward_GB/s": fwd_bytes / fwd_ms / 1e6,
                          "forward_frac_of_hbm_peak": fwd_bytes / fwd_ms / 1e6 / peak,
                          "backward_ms": bwd_ms, "backward_GB/s": bwd_bytes / bwd_ms / 1e6,
                          "backward_frac_of_hbm_peak": bwd_bytes / bwd_ms / 1e6 / peak,
                          "torch_ops_forward_ms": ref_fwd_ms, "torch_ops_forward_backward_ms": ref_step_ms,
                          "max_abs_diff_vs_torch_ops": float((normals - torch_ops_normals(v, t)).abs().max())}))


if __name__ == "__main__":
    main()
