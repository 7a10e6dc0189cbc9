"""Step time of each rank's share of c4 (256 views) on ONE GPU: contiguous slices (shard_views) against strided
ones (view i belongs to rank i mod N).  python profiles/tools/slice_times.py [N]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    full = bench.make_workload("c4")
    B = full["clip_vertices"].shape[0]
    dev = torch.device("cuda:0")
    for label, pick in (("contiguous", lambda r: np.arange(r * (B // n), (r + 1) * (B // n))),
                        ("strided", lambda r: np.arange(r, B, n))):
        times = []
        for r in range(n):
            sc = dict(full)
            idx = pick(r)
            for key in ("clip_vertices", "attributes", "camera_matrices"):
                sc[key] = np.ascontiguousarray(full[key][idx])
            m = bench.measure_on_device(sc, dev, "atomic", 10, 3, 0, 1, "nccl", sample_clocks=False)
            times.append(m["ms_per_step"])
        print(label, "ms per step of each rank's share:", " ".join("%.4f" % t for t in times),
              "| max %.4f mean %.4f" % (max(times), sum(times) / n))


if __name__ == "__main__":
    main()
