"""render() without specular colours on the c2 geometry (50 244-triangle sphere, 64 views x 512^2, one light):
fwd+bwd time of (a) rasterize_clip_space + shade_diffuse (attribute image [B,H,W,9] and its gradient go through
HBM) and (b) the fused render path (lighting inside the resolve / backward kernels).
    python profiles/tools/render_bench.py"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import pytorch_mesh_renderer_b200 as pmr  # noqa: E402
from pytorch_mesh_renderer_b200 import synthetic as S  # noqa: E402
from pytorch_mesh_renderer_b200.render import render_diffuse_clip_space, shade_diffuse  # noqa: E402


def main():
    B, size = 64, 512
    sc = S.sphere_views(159, 158, B, size)
    world = sc["world_vertices"].astype(np.float32)
    normals = world / np.linalg.norm(world, axis=1, keepdims=True)
    diffuse = np.random.default_rng(0).random(world.shape, dtype=np.float32)
    attrs = np.concatenate([normals, world, diffuse], 1)[None].repeat(B, 0).astype(np.float32)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    clip, at, tris = dev(sc["clip_vertices"]), dev(attrs), dev(sc["triangles"])
    lp = torch.tensor([[[0.0, 0.0, 6.0]]]).repeat(B, 1, 1).cuda()
    li = torch.ones((B, 1, 3)).cuda()
    bg = torch.full((9,), -1.0, device="cuda")
    grad = torch.randn((B, size, size, 4), generator=torch.Generator().manual_seed(1)).cuda()

    def unfused():
        cv, a = clip.detach().requires_grad_(True), at.detach().requires_grad_(True)
        shade_diffuse(pmr.rasterize_clip_space(cv, a, tris, size, size, bg), lp, li).backward(grad)

    def fused():
        cv, a = clip.detach().requires_grad_(True), at.detach().requires_grad_(True)
        render_diffuse_clip_space(cv, a, tris, lp, li, size, size).backward(grad)

    for name, fn in (("rasterize_clip_space + shade_diffuse", unfused), ("fused render path", fused)):
        for _ in range(5):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(json.dumps({"path": name, "fwd+bwd ms": ms, "Mpixels/s": B * size * size / ms / 1e3}))


if __name__ == "__main__":
    main()
