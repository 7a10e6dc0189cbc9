"""Golden vectors for the mesh utilities around the renderer, produced by RUNNING
THE UNMODIFIED REFERENCE (src/common/meshes.py, shapes.py, obj_utils.py, src/mesh_renderer/render.py with the
C++ rasterizer) in the build container:

    python tests/golden/make_golden_mesh.py

Writes
  mesh_normals_<case>.npz    vertices, triangles, the reference's compute_vertex_normals and its autograd gradient
                             of sum(normals * g) with respect to the vertices
  mesh_shapes.json           sha256 digests of shapes.sphere / shapes.cube outputs
  mesh_obj_normals.obj/.npz  a file written by the reference's save_obj (with normals) and what its load_obj
  mesh_obj_plain.obj/.npz    returns for it (normalize True / False); the same without normals
  mesh_obj_quads.obj/.npz    hand-written file: quads, v/vt/vn references, a vertex without normal, comments
  mesh_example1_160x120.npz  example1.py's scene (camera, light, white diffuse) on the sphere, 160x120
"""
import hashlib
import json
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import reference_harness as rh  # noqa: E402

torch.set_num_threads(1)
warnings.filterwarnings("ignore", message="Using torch.cross without specifying the dim")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    assert rh.available()
    rh.rasterize_module()
    import importlib
    meshes = importlib.import_module("src.common.meshes")
    shapes = importlib.import_module("src.common.shapes")
    obj_utils = importlib.import_module("src.common.obj_utils")
    render_mod = importlib.import_module("src.mesh_renderer.render")
    rng = np.random.default_rng(7)

    def normals_case(name, vertices, triangles, seed):
        v = torch.tensor(vertices, requires_grad=True)
        n = meshes.compute_vertex_normals(v, torch.from_numpy(triangles))
        g = np.random.default_rng(seed).standard_normal(tuple(n.shape)).astype(np.float32)
        n.backward(torch.from_numpy(g))
        np.savez_compressed(os.path.join(HERE, "mesh_normals_%s.npz" % name), vertices=vertices, triangles=triangles,
                            normals=n.detach().numpy(), grad_normals=g, d_vertices=v.grad.numpy())
        print(name, vertices.shape, triangles.shape)

    sv, st, _ = shapes.sphere(1.0, 12)
    jitter = (0.05 * rng.standard_normal((3,) + tuple(sv.shape))).astype(np.float32)
    normals_case("sphere12", sv.numpy()[None] + jitter, st.numpy(), 11)
    cv, ct, _ = shapes.cube(2.0)
    normals_case("cube", cv.numpy()[None].copy(), ct.numpy(), 12)
    # random indices: repeated vertices inside a triangle, unreferenced vertices (zero normal -> eps branch)
    rv = rng.standard_normal((2, 60, 3)).astype(np.float32)
    rt = rng.integers(0, 50, (150, 3)).astype(np.int32)
    rt[:10, 1] = rt[:10, 0]
    normals_case("random_indices", rv, rt, 13)
    # one vertex shared by many triangles (fan), like the pole of a UV sphere
    fan_v = np.concatenate([np.zeros((1, 3)), np.stack([np.cos(np.linspace(0, 6.2, 97)), np.sin(np.linspace(0, 6.2, 97)),
                                                        0.3 * rng.standard_normal(97)], 1)]).astype(np.float32)
    fan_t = np.stack([np.zeros(96, np.int32), np.arange(1, 97, dtype=np.int32), np.arange(2, 98, dtype=np.int32)], 1)
    normals_case("fan", fan_v[None].copy(), fan_t, 14)

    digests = {}
    for K in (3, 5, 20, 25):
        v, t, n = shapes.sphere(1.5, K)
        digests["sphere_1.5_%d" % K] = dict(vertices=sha(v.numpy()), triangles=sha(t.numpy()), normals=sha(n.numpy()),
                                            counts=[int(v.shape[0]), int(t.shape[0])])
    v, t, n = shapes.cube(2.0)
    digests["cube_2"] = dict(vertices=sha(v.numpy()), triangles=sha(t.numpy()), normals=sha(n.numpy()),
                             counts=[8, 12])
    json.dump(digests, open(os.path.join(HERE, "mesh_shapes.json"), "w"), indent=1, sort_keys=True)

    v, t, n = shapes.sphere(0.8, 6)
    v = v + torch.tensor([0.3, -0.2, 0.5])
    def loaded(path, stem):
        out = {}
        for flag in (True, False):
            lv, lf, ln = obj_utils.load_obj(path, normalize=flag)
            out.update({"vertices_%d" % flag: lv.numpy(), "faces_%d" % flag: lf.numpy(), "normals_%d" % flag: ln.numpy()})
        np.savez_compressed(os.path.join(HERE, stem + ".npz"), **out)
    obj_utils.save_obj(os.path.join(HERE, "mesh_obj_normals.obj"), v, t, n)
    loaded(os.path.join(HERE, "mesh_obj_normals.obj"), "mesh_obj_normals")
    obj_utils.save_obj(os.path.join(HERE, "mesh_obj_plain.obj"), v, t)
    loaded(os.path.join(HERE, "mesh_obj_plain.obj"), "mesh_obj_plain")
    with open(os.path.join(HERE, "mesh_obj_quads.obj"), "w") as f:
        f.write("# hand-written: quads, v/vt/vn references, vertex 6 has no normal\n\n"
                "v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 0.5 0.5 1 1.0\nv 2 2 2\n"
                "vt 0 0\nvt 1 1\n"
                "vn 0 0 1\nvn 0 1 0\nvn 1 0 0\nvn 0.5 0.5 0.70710678\n"
                "f 1/1/1 2/2/1 3/1/1 4/2/1\n"
                "f 1/1/2 2/1/3 5/2/4\n"
                "f 2/1/1 3/1/2 5/2/3\n"
                "f 3/1/4 4/1/4 5/2/4\n"
                "g ignored\n"
                "f 4/1/2 1/1/2 5/2/2\n")
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        loaded(os.path.join(HERE, "mesh_obj_quads.obj"), "mesh_obj_quads")

    # example1.py's scene on the reference's sphere (its faces flipped to the CW winding the renderer's own tests use)
    v, t, n = shapes.sphere(1.0, 25)
    t = torch.flip(t, [1])
    W, H = 160, 120
    image = render_mod.render(v[None], t, n[None], torch.ones_like(v[None]), torch.tensor([[0.0, 0.0, 3.0]]),
                              torch.zeros(1, 3), torch.tensor([[0.0, 1.0, 0.0]]), torch.tensor([[[0.0, 3.0, 0.0]]]),
                              torch.ones(1, 1, 3), W, H)
    np.savez_compressed(os.path.join(HERE, "mesh_example1_160x120.npz"), image=image.reshape(H, W, 4).numpy(),
                        width=W, height=H)
    print("example1", image.shape, float(image.max()))


if __name__ == "__main__":
    main()
