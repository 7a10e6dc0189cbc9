"""Where does the ATOMIC backward differ from the ORDERED one by more than rounding?  One GPU, c4 slice.
    python profiles/tools/atomic_vs_ordered.py [views]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
import pytorch_mesh_renderer_b200 as pmr


def main():
    views = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    sc = bench.make_workload("c4")
    dev = torch.device("cuda:0")
    to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    attrs, tris, bg = (to_dev(sc[k]) for k in ("attributes", "triangles", "background"))
    clip = to_dev(sc["clip_vertices"])[:views].contiguous()
    attrs = attrs[:views].contiguous()
    H, W, A = sc["height"], sc["width"], attrs.shape[2]
    gen = torch.Generator(device=dev); gen.manual_seed(1234)
    grad = torch.randn((views, H, W, A), generator=gen, device=dev)

    def run(mode):
        cv = clip.detach().requires_grad_(True)
        at = attrs.detach().requires_grad_(True)
        with pmr.backward_mode(mode):
            out, (ids, bary, z) = pmr.rasterize_clip_space(cv, at, tris, W, H, bg, return_buffers=True)
            out.backward(grad)
        torch.cuda.synchronize()
        return cv.grad.clone(), at.grad.clone(), ids.clone()

    dv_o, da_o, ids = run("ordered")
    for rep in range(2):
        dv_a, da_a, _ = run("atomic")
        err = (da_a - da_o).abs().amax(dim=2)              # [B,V]
        bad = torch.nonzero(err > 1e-3 * float(da_o.abs().max()))
        print("atomic run %d: %d (image, vertex) pairs off by > 1e-3 of max in d(attributes); d(vertices): %d"
              % (rep, bad.shape[0], int(((dv_a - dv_o).abs().amax(dim=2) > 1e-3 * float(dv_o.abs().max())).sum())))
        t_np = tris.cpu().numpy()
        for b, v in bad[:10].tolist():
            inc = np.nonzero((t_np == v).any(axis=1))[0]
            mask = torch.isin(ids[b], torch.from_numpy(inc).to(dev).int())
            ys, xs = torch.nonzero(mask, as_tuple=True)
            if ys.numel() == 0:
                print("  image %d vertex %d: no pixels?!" % (b, v)); continue
            x0, x1, y0, y1 = int(xs.min()), int(xs.max()), int(ys.min()), int(ys.max())
            print("  image %3d vertex %6d  pixels x %d..%d y %d..%d  block columns %d..%d rows %d..%d  n_px %d\n      d(attr) atomic-ordered %s\n      d(vert) atomic-ordered %s of %s"
                  % (b, v, x0, x1, y0, y1, x0 // 8, x1 // 8, y0 // 4, y1 // 4, int(mask.sum()),
                     np.round((da_a[b, v] - da_o[b, v]).cpu().numpy(), 4), np.round((dv_a[b, v] - dv_o[b, v]).cpu().numpy(), 3),
                     np.round(dv_o[b, v].cpu().numpy(), 2)))


if __name__ == "__main__":
    main()
