"""Run-to-run check on one GPU: the shared-mesh step of bench.py (c4, a slice of its views) executed twice on the
same inputs; forward outputs must be bit-identical, the ATOMIC gradients equal within rounding.
    python profiles/tools/run_to_run.py [views] [config]"""
import sys

import numpy as np
import torch

import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
import pytorch_mesh_renderer_b200 as pmr
from pytorch_mesh_renderer_b200.camera_utils import transform_shared_mesh


def main():
    views = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    name = sys.argv[2] if len(sys.argv) > 2 else "c4"
    sc = bench.make_workload(name)
    dev = torch.device("cuda:0")
    to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    attrs, tris, bg = (to_dev(sc[k]) for k in ("attributes", "triangles", "background"))
    mvp = to_dev(sc["camera_matrices"])[:views]
    attrs = attrs[:views].contiguous()
    world = to_dev(sc["world_vertices"])
    H, W, A = sc["height"], sc["width"], attrs.shape[2]
    gen = torch.Generator(device=dev); gen.manual_seed(1234)
    grad = torch.randn((views, H, W, A), generator=gen, device=dev)
    outs = []
    for rep in range(3):
        wv = world.detach().requires_grad_(True)
        at = attrs.detach().requires_grad_(True)
        cv = transform_shared_mesh(mvp, wv)
        cv.retain_grad()
        out, (ids, bary, z) = pmr.rasterize_clip_space(cv, at, tris, W, H, bg, return_buffers=True)
        out.backward(grad)
        torch.cuda.synchronize()
        outs.append(dict(out=out.detach().clone(), ids=ids.clone(), bary=bary.detach().clone(), dclip=cv.grad.clone(),
                         dattr=at.grad.clone(), dworld=wv.grad.clone()))
    a = outs[0]
    for rep in (1, 2):
        b = outs[rep]
        for k in ("ids", "bary", "out"):
            same = torch.equal(a[k], b[k])
            print("run 0 vs %d  %-6s bit-identical: %s" % (rep, k, same), "" if same else "DIFFERENT entries: %d" % int((a[k] != b[k]).sum()))
        for k in ("dclip", "dattr", "dworld"):
            d = (a[k] - b[k]).abs()
            i = int(d.argmax())
            print("run 0 vs %d  %-6s max |diff| %.6g at flat index %d (values %.8g / %.8g), max |value| %.6g, entries differing by > 1e-3 of max: %d"
                  % (rep, k, float(d.max()), i, float(a[k].flatten()[i]), float(b[k].flatten()[i]), float(a[k].abs().max()),
                     int((d > 1e-3 * a[k].abs().max()).sum())))


if __name__ == "__main__":
    main()
