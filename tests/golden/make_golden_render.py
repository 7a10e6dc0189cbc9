"""Golden vectors for the `render` caller, produced by RUNNING THE UNMODIFIED REFERENCE
(src/mesh_renderer/render.py with the C++ rasterizer switched on) in the build container:

    python tests/golden/make_golden_render.py

Cases: the cube scene of mesh_renderer_test.py:30-70 (diffuse only) and of :72-149 (specular, two lights
per image, ambient, per-image fov, per-vertex and scalar shininess) at 96x72, with the gradient of
sum(render * g) with respect to vertices, normals and diffuse colours from the reference's autograd.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import reference_harness as rh  # noqa: E402

torch.set_num_threads(1)


def main():
    assert rh.available()
    rh.rasterize_module()
    import importlib
    render_mod = importlib.import_module("src.mesh_renderer.render")
    cam = rh.camera_utils()
    cube_v = torch.tensor([[-1, -1, 1], [-1, -1, -1], [-1, 1, -1], [-1, 1, 1], [1, -1, 1],
                           [1, -1, -1], [1, 1, -1], [1, 1, 1]], dtype=torch.float32)
    cube_n = torch.nn.functional.normalize(cube_v, dim=1, p=2)
    tris = torch.tensor([[0, 1, 2], [2, 3, 0], [3, 2, 6], [6, 7, 3], [7, 6, 5], [5, 4, 7],
                         [4, 5, 1], [1, 0, 4], [5, 6, 2], [2, 1, 5], [7, 4, 0], [0, 3, 7]], dtype=torch.int32)
    rot = cam.euler_matrices(torch.tensor([[-20.0, 0.0, 60.0], [45.0, 60.0, 0.0]]))[:, :3, :3]
    W, H = 96, 72

    def run(name, kwargs_extra, eye, center, up, lights_p, lights_i, diffuse, seed):
        v = torch.matmul(torch.stack([cube_v, cube_v]), rot.transpose(1, 2)).clone().requires_grad_(True)
        n = torch.matmul(torch.stack([cube_n, cube_n]), rot.transpose(1, 2)).clone().requires_grad_(True)
        d = diffuse.clone().requires_grad_(True)
        out = render_mod.render(v, tris, n, d, eye, center, up, lights_p, lights_i, W, H, **kwargs_extra)
        g = torch.from_numpy(np.random.default_rng(seed).standard_normal(tuple(out.shape)).astype(np.float32))
        out.backward(g)
        arrays = dict(vertices=v.detach().numpy(), normals=n.detach().numpy(), diffuse=d.detach().numpy(),
                      triangles=tris.numpy(), eye=eye.numpy(), center=center.numpy(), up=up.numpy(),
                      light_positions=lights_p.numpy(), light_intensities=lights_i.numpy(), width=W, height=H,
                      image=out.detach().numpy(), grad_out=g.numpy(), d_vertices=v.grad.numpy(),
                      d_normals=n.grad.numpy(), d_diffuse=d.grad.numpy())
        for k, val in kwargs_extra.items():
            arrays["arg_" + k] = val.numpy() if isinstance(val, torch.Tensor) else np.float32(val)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
        print(name, out.shape, float(out.abs().max()))

    eye = torch.tensor(2 * [[0.0, 0.0, 6.0]]); ctr = torch.zeros(2, 3); up = torch.tensor(2 * [[0.0, 1.0, 0.0]])
    run("render_cube_96x72", {}, eye, ctr, up, torch.tensor([[[0.0, 0.0, 6.0]], [[0.0, 0.0, 6.0]]]),
        torch.ones(2, 1, 3), torch.ones(2, 8, 3), 50)

    eye2 = torch.tensor([[0.0, 0.0, 6.0], [0., 0.2, 18.0]]); ctr2 = torch.tensor([[0.0, 0.0, 0.0], [0.1, -0.1, 0.1]])
    up2 = torch.tensor([[0.0, 1.0, 0.0], [0.1, 1.0, 0.15]])
    lp = torch.tensor([[[0.0, 0.0, 6.0], [1.0, 2.0, 6.0]], [[0.0, -2.0, 4.0], [1.0, 3.0, 4.0]]])
    li = torch.tensor([[[1.0, 1.0, 1.0], [1.0, 1.0, 1.0]], [[2.0, 0.0, 1.0], [0.0, 2.0, 1.0]]])
    diffuse = torch.tensor(2 * [[[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [1.0, 1.0, 1.0], [1.0, 1.0, 0.0],
                                 [1.0, 0.0, 1.0], [0.0, 1.0, 1.0], [0.5, 0.5, 0.5]]])
    spec = torch.tensor(2 * [[[0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [1.0, 1.0, 1.0], [1.0, 1.0, 0.0], [1.0, 0.0, 1.0],
                              [0.0, 1.0, 1.0], [0.5, 0.5, 0.5], [1.0, 0.0, 0.0]]])
    common = dict(specular_colors=spec, ambient_color=torch.tensor([[0., 0., 0.], [0.1, 0.1, 0.2]]),
                  fov_y=torch.tensor([40., 13.3]), near_clip=torch.tensor(0.1), far_clip=torch.tensor(25.0))
    run("render_complex_vertex_shininess_96x72", dict(shininess_coefficients=6.0 * torch.ones(2, 8), **common),
        eye2, ctr2, up2, lp, li, diffuse, 51)
    run("render_complex_scalar_shininess_96x72", dict(shininess_coefficients=torch.tensor(6.0), **common),
        eye2, ctr2, up2, lp, li, diffuse, 52)


if __name__ == "__main__":
    main()
