"""Times the shade_diffuse kernels on a c2-sized attribute image (64 x 512 x 512 x 9) and prints achieved
GB/s against the measured HBM peak.  Algorithmic bytes per pixel: forward reads 36 + writes 16,
backward reads 16 + 36 and writes 36.   python profiles/tools/shade_bench.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pytorch_mesh_renderer_b200 import _lib, ops  # noqa: E402


def main():
    B, H, W, A, L = 64, 512, 512, 9, 1
    g = torch.Generator().manual_seed(0)
    px = torch.randn((B, H, W, A), generator=g).cuda()
    px[..., 6:9] = px[..., 6:9].abs()
    lp = torch.randn((B, L, 3), generator=g).cuda() * 3
    li = torch.rand((B, L, 3), generator=g).cuda()
    grad = torch.randn((B, H, W, 4), generator=g).cuda()
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    P = B * H * W
    px13 = torch.cat([px, torch.rand((B, H, W, 3), generator=g).cuda(), 1.0 + 9.0 * torch.rand((B, H, W, 1), generator=g).cuda()], 3)
    cam = torch.randn((B, 3), generator=g).cuda() * 4
    _, norm2 = ops.shade_phong_forward(px13, lp, li, None, cam, None)
    # specular: two passes each way; bytes = what the two kernels of a call must move (13 channels = 52 B/px)
    for name, fn, nbytes in (("shade_diffuse_forward", lambda: ops.shade_diffuse_forward(px, lp, li, None), P * 52),
                             ("shade_diffuse_backward", lambda: ops.shade_diffuse_backward(grad, px, lp, li, None), P * 88),
                             ("shade_phong_forward (norm pass + shade)", lambda: ops.shade_phong_forward(px13, lp, li, None, cam, None), P * (52 + 52 + 16)),
                             ("shade_phong_backward (sums pass + gradient)", lambda: ops.shade_phong_backward(grad, px13, lp, li, None, cam, None, norm2), P * (68 + 68 + 52))):
        for _ in range(5):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        _lib.enable_stage_timing(0, True)
        _lib.read_stage_timing(0, reset=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        kernel_ms, launches = _lib.read_stage_timing(0, reset=True)["shade"]
        _lib.enable_stage_timing(0, False)
        loop_ms = e0.elapsed_time(e1) / 20          # includes allocating the output tensor in every call
        ms = kernel_ms / max(launches, 1)           # the kernel alone (event pair around the launch)
        gbs = nbytes / ms / 1e6
        print(json.dumps({"kernel": name, "ms": ms, "loop_ms_per_call": loop_ms, "algorithmic_bytes": nbytes, "GB/s": gbs, "frac_of_measured_hbm_peak": gbs / peak}))


if __name__ == "__main__":
    main()
