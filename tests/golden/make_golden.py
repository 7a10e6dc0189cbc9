"""Generates tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE in the build container.

    python tests/golden/make_golden.py

The reference kernel (oracle/_ref, compiled from
/root/reference/src/mesh_renderer/kernels/rasterize_triangles.cpp) and the reference Python
layers (imported from /root/reference: src/mesh_renderer/rasterize.py with
USE_CPP_RASTERIZER=True, src/common/camera_utils.py) produce every output stored here; nothing
from this repository's oracle or CUDA path is involved.  torch runs with one thread so that
the reference's index_put_ accumulation is deterministic.

Stored per case: the inputs and the reference outputs (small cases in full; the 640x480
cases of the reference's own tests as sha256 digests of the raw little-endian arrays plus a
strided sample, to keep the repository small).  tests/golden/reference_png/ holds the PNG
fixtures of the reference's tests (rasterize_triangles_test.py:72-117, mesh_renderer_test.py
:30-70) copied from /root/reference/src/mesh_renderer/test_data/.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import reference_harness as rh  # noqa: E402

torch.set_num_threads(1)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ref_kernel_case(verts, tris, W, H, seed):
    """rasterize_triangles_cpp.forward + backward on one image."""
    K = rh.kernel()
    v, t = torch.from_numpy(verts), torch.from_numpy(tris)
    ids, bary, z = K.forward(v, t, W, H)
    g = np.random.default_rng(seed).standard_normal((H, W, 3)).astype(np.float32)
    dv, = K.backward(torch.from_numpy(g), v, t, ids, bary.detach())
    return dict(vertices=verts, triangles=tris, width=W, height=H, ids=ids.numpy(),
                bary=bary.detach().numpy(), z=z.numpy(), df_dbary=g, df_dvertices=dv.numpy())


def ref_full_case(clip, attrs, tris, W, H, bg, seed):
    """Reference rasterize_clip_space forward + autograd backward (rast.py:66-152)."""
    R = rh.rasterize_module()
    tv = torch.tensor(clip, requires_grad=True)
    ta = torch.tensor(attrs, requires_grad=True)
    out = R.rasterize_clip_space(tv, ta, torch.from_numpy(tris), W, H, torch.from_numpy(bg))
    g = np.random.default_rng(seed).standard_normal(tuple(out.shape)).astype(np.float32)
    out.backward(torch.from_numpy(g))
    K = rh.kernel()
    ids, bary, z = [], [], []
    for b in range(clip.shape[0]):
        i, ba, zz = K.forward(torch.from_numpy(clip[b]), torch.from_numpy(tris), W, H)
        ids.append(i.numpy()); bary.append(ba.detach().numpy()); z.append(zz.numpy())
    return dict(clip_vertices=clip, attributes=attrs, triangles=tris, width=W, height=H,
                background=bg, out=out.detach().numpy(), grad_out=g,
                d_clip_vertices=tv.grad.numpy(), d_attributes=ta.grad.numpy(),
                ids=np.stack(ids), bary=np.stack(bary), z=np.stack(z))


def soup(rng, T, spread=0.08, mixed_w=0.0, behind=0.0):
    """T independent triangles (V = 3T), perspective-scaled like rasterize_triangles_test.py:52-53."""
    c = rng.uniform(-0.9, 0.9, (T, 1, 2))
    xy = c + spread * rng.standard_normal((T, 3, 2))
    z = rng.uniform(-0.9, 0.9, (T, 3, 1))
    w = rng.uniform(0.5, 2.0, (T, 3, 1))
    if mixed_w > 0:
        w = np.where(rng.random((T, 3, 1)) < mixed_w, -w, w)
    if behind > 0:
        w = np.where(rng.random((T, 1, 1)) < behind, -np.abs(w), w)
    v = np.concatenate([xy, z, np.ones_like(z)], 2) * w
    return v.reshape(3 * T, 4).astype(np.float32), np.arange(3 * T, dtype=np.int32).reshape(T, 3)


def grid_mesh(rng, n, jitter=0.3):
    """(n+1)^2 shared vertices, 2n^2 triangles with shared edges (watertightness / crack test)."""
    u = np.linspace(-0.8, 0.8, n + 1)
    X, Y = np.meshgrid(u, u)
    xy = np.stack([X, Y], -1).reshape(-1, 2) + jitter * (1.6 / n) * rng.uniform(-1, 1, ((n + 1) ** 2, 2))
    z = rng.uniform(-0.5, 0.5, ((n + 1) ** 2, 1))
    w = rng.uniform(0.7, 1.5, ((n + 1) ** 2, 1))
    v = (np.concatenate([xy, z, np.ones_like(z)], 1) * w).astype(np.float32)
    tris = []
    for j in range(n):
        for i in range(n):
            a = j * (n + 1) + i
            tris += [[a, a + 1, a + n + 2], [a + n + 2, a + n + 1, a]]
    t = np.array(tris, np.int32)
    flip = rng.random(len(t)) < 0.5
    t[flip] = t[flip][:, ::-1]
    return v, t


CUBE_V = np.array([[-1, -1, 1], [-1, -1, -1], [-1, 1, -1], [-1, 1, 1], [1, -1, 1],
                   [1, -1, -1], [1, 1, -1], [1, 1, 1]], np.float32)
CUBE_T = np.array([[0, 1, 2], [2, 3, 0], [3, 2, 6], [6, 7, 3], [7, 6, 5], [5, 4, 7],
                   [4, 5, 1], [1, 0, 4], [5, 6, 2], [2, 1, 5], [7, 4, 0], [0, 3, 7]], np.int32)


def main():
    assert rh.available(), "needs /root/reference and oracle/_ref (python oracle/build_ref.py)"
    cam = rh.camera_utils()
    rng = np.random.default_rng(20261018)
    cases = {}
    digests = {}

    # --- rasterize_triangles_test.py:37-77: the two single-triangle kernel tests (640x480)
    tri = np.array([[-0.5, -0.5, 0.8, 1.0], [0.0, 0.5, 0.3, 1.0], [0.5, -0.5, 0.3, 1.0]], np.float32)
    for name, wv in (("simple_triangle", (1.0, 1.0, 1.0)), ("perspective_triangle", (0.2, 0.5, 2.0))):
        v = tri * np.array(wv, np.float32).reshape(3, 1)
        full = ref_kernel_case(v, np.array([[0, 1, 2]], np.int32), 640, 480, 1)
        digests[name] = {k: sha(full[k]) for k in ("ids", "bary", "z", "df_dvertices")}
        small = dict(full)
        for k in ("ids", "bary", "z"):
            small[k + "_sample"] = full[k][::7, ::5].copy()
            del small[k]
        del small["df_dbary"]          # regenerated from seed 1 by the test
        small["df_dbary_seed"] = 1
        cases[name + "_640x480"] = small
        cases[name + "_80x60"] = ref_kernel_case(v, np.array([[0, 1, 2]], np.int32), 80, 60, 2)

    # --- rasterize_triangles_test.py:79-117: two cubes through rasterize() (A=4)
    persp = cam.perspective(640 / 480, torch.tensor([40.0]), torch.tensor([0.01]), torch.tensor([10.0]))
    center = torch.tensor([[0.0, 0, 0]]); up = torch.tensor([[0.0, 1, 0]])
    proj = torch.cat([torch.matmul(persp, cam.look_at(torch.tensor([[2.0, 3, 6]]), center, up)),
                      torch.matmul(persp, cam.look_at(torch.tensor([[-3.0, 1, 6]]), center, up))], 0)
    world = torch.stack([torch.from_numpy(CUBE_V)] * 2)
    clip = cam.transform_homogeneous(proj, world).numpy()
    rgba = np.concatenate([CUBE_V * 0.5 + 0.5, np.ones((8, 1), np.float32)], 1)
    attrs = np.stack([rgba, rgba]).astype(np.float32)
    bg0 = np.zeros(4, np.float32)
    full = ref_full_case(clip, attrs, CUBE_T, 640, 480, bg0, 3)
    digests["two_cubes"] = {k: sha(full[k]) for k in ("ids", "bary", "z", "out", "d_clip_vertices", "d_attributes")}
    cases["two_cubes_640x480"] = dict(
        clip_vertices=clip, attributes=attrs, triangles=CUBE_T, width=640, height=480,
        background=bg0, camera_matrices=proj.numpy(), world_vertices=world.numpy(),
        grad_out_seed=3, d_clip_vertices=full["d_clip_vertices"], d_attributes=full["d_attributes"],
        out_sample=full["out"][:, ::7, ::5].copy(), ids_sample=full["ids"][:, ::7, ::5].copy())
    cases["two_cubes_96x72"] = ref_full_case(clip, attrs, CUBE_T, 96, 72, bg0, 4)
    # A = 9 (render.py:172-181 packs normals, positions, diffuse) with background -1 (render.py:197)
    a9 = rng.uniform(-1, 1, (2, 8, 9)).astype(np.float32)
    cases["two_cubes_a9_96x72"] = ref_full_case(clip, a9, CUBE_T, 96, 72, -np.ones(9, np.float32), 5)

    # --- rasterize_triangles_test.py:160-199: precomputed cube clip coordinates, 28x21
    jac = np.array([[-0.43889722, -0.53184521, 0.85293502, 1.0], [-0.37635487, 0.22206162, 0.90555805, 1.0],
                    [-0.22849123, 0.76811147, 0.80993629, 1.0], [-0.2805393, -0.14092168, 0.71602166, 1.0],
                    [0.18631913, -0.62634289, 0.88603103, 1.0], [0.16183566, 0.08129397, 0.93020856, 1.0],
                    [0.44147962, 0.53497446, 0.85076219, 1.0], [0.53008741, -0.31276882, 0.77620775, 1.0]], np.float32)
    cases["jacobian_cube_28x21"] = ref_kernel_case(jac, CUBE_T, 28, 21, 6)

    # --- mesh_renderer_test.py:36-57: BASELINE config c1 geometry (euler cube, eye z=6), A=9
    rot = cam.euler_matrices(torch.tensor([[-20.0, 0.0, 60.0], [45.0, 60.0, 0.0]]))[:, :3, :3]
    wv = torch.matmul(torch.stack([torch.from_numpy(CUBE_V)] * 2), rot.transpose(1, 2))
    eye = torch.tensor(2 * [[0.0, 0.0, 6.0]]); ctr = torch.zeros(2, 3); upv = torch.tensor(2 * [[0.0, 1.0, 0.0]])
    p1 = cam.perspective(640 / 480, torch.tensor([40.0, 40.0]), torch.tensor([0.01, 0.01]), torch.tensor([10.0, 10.0]))
    mvp = torch.matmul(p1, cam.look_at(eye, ctr, upv))
    clip_c1 = cam.transform_homogeneous(mvp, wv).numpy()
    nrm = torch.nn.functional.normalize(torch.from_numpy(CUBE_V), dim=1)
    nw = torch.matmul(torch.stack([nrm] * 2), rot.transpose(1, 2))
    attrs_c1 = torch.cat([nw, wv, torch.ones_like(wv)], 2).numpy().astype(np.float32)
    full = ref_full_case(clip_c1, attrs_c1, CUBE_T, 640, 480, -np.ones(9, np.float32), 7)
    digests["c1_cube"] = {k: sha(full[k]) for k in ("ids", "bary", "z", "out", "d_clip_vertices", "d_attributes")}
    cases["c1_cube_640x480"] = dict(
        clip_vertices=clip_c1, attributes=attrs_c1, triangles=CUBE_T, width=640, height=480,
        background=-np.ones(9, np.float32), grad_out_seed=7, camera_matrices=mvp.numpy(),
        world_vertices=wv.numpy(), d_clip_vertices=full["d_clip_vertices"],
        d_attributes=full["d_attributes"], out_sample=full["out"][:, ::7, ::5].copy(),
        ids_sample=full["ids"][:, ::7, ::5].copy())
    cases["c1_cube_64x48"] = ref_full_case(clip_c1, attrs_c1, CUBE_T, 64, 48, -np.ones(9, np.float32), 8)

    # --- edge cases the reference kernel handles (SURVEY.md F1, F4, F7-adjacent, K.cpp:339, :356-360)
    v, t = soup(rng, 400, spread=0.10)
    cases["soup400_97x61"] = ref_kernel_case(v, t, 97, 61, 9)                       # odd sizes
    v, t = soup(rng, 300, spread=0.15, mixed_w=0.06, behind=0.05)
    cases["soup_mixed_w_64x64"] = ref_kernel_case(v, t, 64, 64, 10)                # w<0 handling
    v, t = soup(rng, 40, spread=0.5)
    t = np.concatenate([t, t[::-1], t], 0)                                         # exact duplicates
    cases["soup_ties_72x40"] = ref_kernel_case(v, t, 72, 40, 11)                   # tie -> max id
    v, t = grid_mesh(rng, 12)
    cases["grid_shared_edges_90x70"] = ref_kernel_case(v, t, 90, 70, 12)           # no cracks, both windings
    v, t = soup(rng, 30, spread=0.3)
    v[t[3, 1]] = v[t[3, 0]]                                                        # zero-area triangle
    v[t[5]] = v[t[5, 0]]                                                           # point triangle
    t[7] = [t[7, 0], t[7, 0], t[7, 2]]                                             # repeated index
    v[t[9], 2] *= 3.0                                                              # outside z range
    cases["soup_degenerate_48x48"] = ref_kernel_case(v, t, 48, 48, 13)
    cases["empty_mesh_16x12"] = ref_kernel_case(np.zeros((1, 4), np.float32), np.zeros((0, 3), np.int32), 16, 12, 14)
    v, t = soup(rng, 5, spread=0.02)
    cases["one_pixel_image"] = ref_kernel_case(v, t, 1, 1, 15)
    v, t = soup(rng, 60, spread=1.5)
    cases["soup_offscreen_50x34"] = ref_kernel_case(v, t, 50, 34, 16)              # bbox clamping

    # --- full path, several attribute counts (torch's inner-sum order changes with A)
    for A in (1, 3, 4, 5, 9, 12, 13):
        v, t = soup(rng, 120, spread=0.12)
        clip = np.stack([v, v * np.float32(1.0)])
        clip[1, :, 0] += np.float32(0.07) * clip[1, :, 3]
        at = rng.uniform(-1, 1, (2, v.shape[0], A)).astype(np.float32)
        bgv = rng.uniform(-1, 1, A).astype(np.float32)
        cases["full_soup_A%d_56x40" % A] = ref_full_case(clip, at, t, 56, 40, bgv, 20 + A)
    v, t = grid_mesh(rng, 10)
    clip = np.stack([v, v, v]); clip[1, :, 1] *= -1; clip[2, :, :2] *= np.float32(0.5)
    at = rng.uniform(0, 1, (3, v.shape[0], 9)).astype(np.float32)
    cases["full_grid_A9_64x64"] = ref_full_case(clip, at, t, 64, 64, -np.ones(9, np.float32), 40)

    total = 0
    for name, c in cases.items():
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **c)
        total += os.path.getsize(path)
    with open(os.path.join(HERE, "digests.json"), "w") as f:
        json.dump(digests, f, indent=1, sort_keys=True)
    print("wrote %d cases, %.2f MB" % (len(cases), total / 1e6))


if __name__ == "__main__":
    main()
