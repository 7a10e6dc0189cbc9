"""Parity of the CUDA path (through the C ABI) against the reference golden vectors and the CPU
oracle.  Bars (BASELINE.json north_star / BASELINE.md section 5):
  * triangle ids, coverage: bit-exact;
  * barycentrics, z, interpolated image: bit-exact is asserted (the kernels reproduce the
    reference's rounding points; the stated tolerance 1e-6 + 1e-5*|ref| is the fallback bar and is
    checked by assert_close where summation order legitimately differs);
  * gradients in ORDERED mode: bit-exact (reference summation order);
  * gradients in ATOMIC mode: compared with the fp64-accumulated evaluation of the same fp32 terms,
    next to the reference's own distance from it.
"""
import ctypes
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, assert_bits, golden_names, grad_from_seed, load_golden

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

KERNEL_CASES = [n for n in golden_names() if "df_dbary" in np.load(os.path.join(GOLDEN_DIR, n + ".npz")).files
                and not n.endswith("640x480") and not n.startswith("render_")]
FULL_CASES = [n for n in golden_names() if "clip_vertices" in np.load(os.path.join(GOLDEN_DIR, n + ".npz")).files
              and not n.endswith("640x480") and not n.startswith("render_")]
DIGESTS = json.load(open(os.path.join(GOLDEN_DIR, "digests.json")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def pmr():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pytorch_mesh_renderer_b200 as m
    return m


@pytest.fixture(params=[64, 0], ids=["small-mesh-path", "binned-path"])
def threshold(request, pmr):
    """Run every case through both forward paths: whole-mesh tiles and binned tile lists."""
    from pytorch_mesh_renderer_b200 import _lib
    ctx = _lib.context(torch.cuda.current_device())
    _lib.load().pmr_set_small_mesh_threshold(ctx, request.param)
    yield request.param
    _lib.load().pmr_set_small_mesh_threshold(ctx, 64)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def atomic_close(actual, ref32, ref64, what):
    """ATOMIC mode sums the same fp32 terms in another order: its distance from the exactly summed
    value must be of the order of the reference's own distance from it (oracle.atomic_mode_bound)."""
    from oracle import oracle as o
    actual = np.asarray(actual, np.float64)
    finite = np.isfinite(ref64) & np.isfinite(np.asarray(ref32, np.float64))
    assert (np.isfinite(actual) == finite).all(), "%s: non-finite entries differ from the reference's" % what
    err = np.abs(actual - ref64)[finite]
    bound = o.atomic_mode_bound(ref32, ref64)[finite]
    assert (err <= bound).all(), "%s: max err %g vs bound %g (reference's own max err %g)" % (
        what, err.max(), bound.min(), np.abs(np.asarray(ref32, np.float64) - ref64)[finite].max())


@pytest.mark.parametrize("name", KERNEL_CASES)
def test_kernel_golden(pmr, oracle, threshold, name):
    from pytorch_mesh_renderer_b200 import ops
    c = load_golden(name)
    W, H = int(c["width"]), int(c["height"])
    v, t = dev(c["vertices"])[None], dev(c["triangles"])
    ids, bary, z = ops.rasterize_forward(v, t, W, H)
    assert_bits(ids[0].cpu().numpy(), c["ids"], "ids")
    assert_bits(bary[0].cpu().numpy(), c["bary"], "bary")
    assert_bits(z[0].cpu().numpy(), c["z"], "z")
    g = dev(c["df_dbary"])[None]
    dv = ops.rasterize_backward(g, v, t, dev(c["ids"])[None], dev(c["bary"])[None], "ordered")
    assert_bits(dv[0].cpu().numpy(), c["df_dvertices"], "df_dvertices (ordered)")
    dva = ops.rasterize_backward(g, v, t, dev(c["ids"])[None], dev(c["bary"])[None], "atomic")
    ref64 = oracle.backward_f64acc(c["df_dbary"], c["vertices"], c["triangles"], c["ids"], c["bary"])
    atomic_close(dva[0].cpu().numpy(), c["df_dvertices"], ref64, "df_dvertices (atomic)")
    assert not dva[0, :, 2].any().item()


@pytest.mark.parametrize("name", FULL_CASES)
def test_full_path_golden(pmr, oracle, threshold, name):
    c = load_golden(name)
    W, H = int(c["width"]), int(c["height"])
    cv = dev(c["clip_vertices"]).requires_grad_(True)
    at = dev(c["attributes"]).requires_grad_(True)
    with pmr.backward_mode("ordered"):
        out, (ids, bary, z) = pmr.rasterize_clip_space(cv, at, dev(c["triangles"]), W, H, dev(c["background"]),
                                                        return_buffers=True)
        out.backward(dev(c["grad_out"]))
    assert_bits(ids.cpu().numpy(), c["ids"], "ids")
    assert_bits(bary.detach().cpu().numpy(), c["bary"], "bary")
    assert_bits(z.detach().cpu().numpy(), c["z"], "z")
    assert_bits(out.detach().cpu().numpy(), c["out"], "out")
    assert_bits(at.grad.cpu().numpy(), c["d_attributes"], "d_attributes (ordered)")
    assert_bits(cv.grad.cpu().numpy(), c["d_clip_vertices"], "d_clip_vertices (ordered)")
    # throughput mode
    cv2 = dev(c["clip_vertices"]).requires_grad_(True)
    at2 = dev(c["attributes"]).requires_grad_(True)
    with pmr.backward_mode("atomic"):
        out2 = pmr.rasterize_clip_space(cv2, at2, dev(c["triangles"]), W, H, dev(c["background"]))
        out2.backward(dev(c["grad_out"]))
    assert_bits(out2.detach().cpu().numpy(), c["out"], "out")
    y = oracle.rasterize_clip_space(c["clip_vertices"], c["attributes"], c["triangles"], W, H, c["background"],
                                    grad_out=c["grad_out"], f64_yardstick=True)
    atomic_close(at2.grad.cpu().numpy(), c["d_attributes"], y["d_attributes_f64"], "d_attributes (atomic)")
    atomic_close(cv2.grad.cpu().numpy(), c["d_clip_vertices"], y["d_vertices_f64"], "d_clip_vertices (atomic)")


@pytest.mark.parametrize("name", ["simple_triangle", "perspective_triangle"])
def test_reference_triangle_tests_640x480(pmr, threshold, name):
    """rasterize_triangles_test.py:72-77 at the test's own resolution, digest-pinned."""
    from pytorch_mesh_renderer_b200 import ops
    c = load_golden(name + "_640x480")
    v, t = dev(c["vertices"])[None], dev(c["triangles"])
    ids, bary, z = ops.rasterize_forward(v, t, 640, 480)
    d = DIGESTS[name]
    assert sha(ids[0].cpu().numpy()) == d["ids"]
    assert sha(bary[0].cpu().numpy()) == d["bary"]
    assert sha(z[0].cpu().numpy()) == d["z"]
    g = dev(grad_from_seed(c["df_dbary_seed"], (480, 640, 3)))[None]
    dv = ops.rasterize_backward(g, v, t, ids, bary, "ordered")
    assert_bits(dv[0].cpu().numpy(), c["df_dvertices"], "df_dvertices")


@pytest.mark.parametrize("name", ["two_cubes", "c1_cube"])
def test_reference_cube_tests_640x480(pmr, threshold, name):
    """rasterize_triangles_test.py:79-117 (A=4) and BASELINE config c1 (A=9) at 2 x 640x480."""
    c = load_golden(name + "_640x480")
    A = c["attributes"].shape[2]
    g = dev(grad_from_seed(c["grad_out_seed"], (2, 480, 640, A)))
    cv = dev(c["clip_vertices"]).requires_grad_(True)
    at = dev(c["attributes"]).requires_grad_(True)
    with pmr.backward_mode("ordered"):
        out, (ids, bary, z) = pmr.rasterize_clip_space(cv, at, dev(c["triangles"]), 640, 480, dev(c["background"]),
                                                        return_buffers=True)
        out.backward(g)
    d = DIGESTS[name]
    assert sha(ids.cpu().numpy()) == d["ids"]
    assert sha(bary.detach().cpu().numpy()) == d["bary"]
    assert sha(z.detach().cpu().numpy()) == d["z"]
    assert sha(out.detach().cpu().numpy()) == d["out"]
    assert_bits(at.grad.cpu().numpy(), c["d_attributes"], "d_attributes")
    assert_bits(cv.grad.cpu().numpy(), c["d_clip_vertices"], "d_clip_vertices")


def test_reference_png_fixtures(pmr):
    """The PNG fixtures of rasterize_triangles_test.py:72-117 under its comparison rule
    (test_utils.py:105-160): at most 0.1 % of pixels off by more than 0.01."""
    from PIL import Image

    def near(image, name):
        png = np.asarray(Image.open(os.path.join(GOLDEN_DIR, "reference_png", name))).astype(np.float64) / 255.0
        diff = np.abs(png - np.clip(image, 0.0, 1.0))
        return np.any(diff > 0.01, axis=2).mean() <= 0.001

    for case, png in (("simple_triangle", "Simple_Triangle.png"),
                      ("perspective_triangle", "Perspective_Corrected_Triangle.png")):
        c = load_golden(case + "_640x480")
        # CPU tensors in, CPU tensors out: the reference test's own calling convention.
        _, bary, _ = pmr.rasterize_barycentric(torch.from_numpy(c["vertices"]), torch.from_numpy(c["triangles"]), 640, 480)
        assert bary.device.type == "cpu" and bary.shape == (480, 640, 3)
        assert near(np.concatenate([bary.numpy(), np.ones((480, 640, 1), np.float32)], 2), png)
    c = load_golden("two_cubes_640x480")
    out = pmr.rasterize(torch.from_numpy(c["world_vertices"]), torch.from_numpy(c["attributes"]),
                        torch.from_numpy(c["triangles"]), torch.from_numpy(c["camera_matrices"]), 640, 480,
                        torch.from_numpy(c["background"]))
    for i in (0, 1):
        assert near(out[i].numpy(), "Unlit_Cube_%d.png" % i)


def test_barycentric_rasterizer_signature(pmr):
    """ext.py:8,43,63: (vertices[V,4], triangles, W, H) -> (ids, bary, z); backward returns
    (df_dvertices, zeros_like(triangles), None, None)."""
    c = load_golden("jacobian_cube_28x21")
    v = dev(c["vertices"]).requires_grad_(True)
    t = dev(c["triangles"])
    with pmr.backward_mode("ordered"):
        ids, bary, z = pmr.BarycentricRasterizer.apply(v, t, 28, 21)
        assert ids.shape == (21, 28) and ids.dtype == torch.int32 and bary.shape == (21, 28, 3) and z.shape == (21, 28)
        bary.backward(dev(c["df_dbary"]))
    assert_bits(v.grad.cpu().numpy(), c["df_dvertices"], "df_dvertices")


def test_error_behaviour(pmr):
    v = torch.zeros(1, 3, 4, device="cuda"); a = torch.zeros(1, 3, 2, device="cuda")
    t = torch.zeros(1, 3, dtype=torch.int32, device="cuda"); bg = torch.zeros(2, device="cuda")
    with pytest.raises(ValueError, match="Image width must be > 0"):
        pmr.rasterize_clip_space(v, a, t, 0, 4, bg)
    with pytest.raises(ValueError, match="Image height must be > 0"):
        pmr.rasterize_clip_space(v, a, t, 4, -1, bg)
    with pytest.raises(ValueError, match="must be 3D"):
        pmr.rasterize_clip_space(v[0], a, t, 4, 4, bg)
    with pytest.raises(RuntimeError, match="expected scalar type Int"):
        pmr.rasterize_clip_space(v, a, t.long(), 4, 4, bg)
    with pytest.raises(RuntimeError, match="expected scalar type Float"):
        pmr.rasterize_barycentric(v[0].double(), t, 4, 4)


def test_sphere_views_vs_oracle(pmr, oracle):
    """A reduced c2: 3 120-triangle UV sphere, 3 views, 160x160, A=9 -- binned path, shared edges,
    both windings, pole fans."""
    from pytorch_mesh_renderer_b200 import synthetic as S
    sc = S.sphere_views(40, 39, 3, 160)
    g = S.upstream_gradient((3, 160, 160, 9))
    ref = oracle.rasterize_clip_space(sc["clip_vertices"], sc["attributes"], sc["triangles"], 160, 160,
                                      sc["background"], grad_out=g, f64_yardstick=True)
    cv = dev(sc["clip_vertices"]).requires_grad_(True)
    at = dev(sc["attributes"]).requires_grad_(True)
    with pmr.backward_mode("ordered"):
        out, (ids, bary, z) = pmr.rasterize_clip_space(cv, at, dev(sc["triangles"]), 160, 160, dev(sc["background"]),
                                                        return_buffers=True)
        out.backward(dev(g))
    assert_bits(ids.cpu().numpy(), ref["ids"], "ids")
    assert_bits(bary.detach().cpu().numpy(), ref["bary"], "bary")
    assert_bits(z.detach().cpu().numpy(), ref["z"], "z")
    assert_bits(out.detach().cpu().numpy(), ref["out"], "out")
    assert_bits(at.grad.cpu().numpy(), ref["d_attributes"], "d_attributes")
    assert_bits(cv.grad.cpu().numpy(), ref["d_vertices"], "d_vertices")
    cv2 = dev(sc["clip_vertices"]).requires_grad_(True)
    at2 = dev(sc["attributes"]).requires_grad_(True)
    with pmr.backward_mode("atomic"):
        pmr.rasterize_clip_space(cv2, at2, dev(sc["triangles"]), 160, 160, dev(sc["background"])).backward(dev(g))
    atomic_close(at2.grad.cpu().numpy(), ref["d_attributes"], ref["d_attributes_f64"], "d_attributes (atomic)")
    atomic_close(cv2.grad.cpu().numpy(), ref["d_vertices"], ref["d_vertices_f64"], "d_vertices (atomic)")


def test_occlusion_soup_vs_oracle(pmr, oracle):
    """A reduced c5: 300 large overlapping triangles (depth complexity ~10), 2 x 192x192."""
    from pytorch_mesh_renderer_b200 import synthetic as S
    sc = S.occlusion_soup(2, 192, n_triangles=300, scale=0.4)
    ref = oracle.rasterize_clip_space(sc["clip_vertices"], sc["attributes"], sc["triangles"], 192, 192, sc["background"])
    out, (ids, bary, z) = pmr.rasterize_clip_space(dev(sc["clip_vertices"]), dev(sc["attributes"]), dev(sc["triangles"]),
                                                   192, 192, dev(sc["background"]), return_buffers=True)
    assert_bits(ids.cpu().numpy(), ref["ids"], "ids")
    assert_bits(bary.cpu().numpy(), ref["bary"], "bary")
    assert_bits(z.cpu().numpy(), ref["z"], "z")
    assert_bits(out.cpu().numpy(), ref["out"], "out")


def test_deep_occlusion_vs_oracle(pmr, oracle):
    """Depth complexity ~45 (1 500 large triangles, 2 x 256x256): exercises the hierarchical-z and
    conservative edge rejection of the large-triangle path; still bit-exact, gradients included."""
    from pytorch_mesh_renderer_b200 import synthetic as S
    sc = S.occlusion_soup(2, 256, n_triangles=1500, scale=0.4)
    g = S.upstream_gradient((2, 256, 256, 9), seed=3)
    ref = oracle.rasterize_clip_space(sc["clip_vertices"], sc["attributes"], sc["triangles"], 256, 256, sc["background"],
                                      grad_out=g)
    cv = dev(sc["clip_vertices"]).requires_grad_(True)
    at = dev(sc["attributes"]).requires_grad_(True)
    with pmr.backward_mode("ordered"):
        out, (ids, bary, z) = pmr.rasterize_clip_space(cv, at, dev(sc["triangles"]), 256, 256, dev(sc["background"]),
                                                        return_buffers=True)
        out.backward(dev(g))
    assert_bits(ids.cpu().numpy(), ref["ids"], "ids")
    assert_bits(bary.detach().cpu().numpy(), ref["bary"], "bary")
    assert_bits(z.detach().cpu().numpy(), ref["z"], "z")
    assert_bits(out.detach().cpu().numpy(), ref["out"], "out")
    assert_bits(cv.grad.cpu().numpy(), ref["d_vertices"], "d_vertices")
    assert_bits(at.grad.cpu().numpy(), ref["d_attributes"], "d_attributes")


def test_c_abi_host_entry_point(pmr, oracle):
    """pmr_rasterize_clip_space_host with plain numpy host buffers (what bench.py times as e2e)."""
    from pytorch_mesh_renderer_b200 import _lib
    c = load_golden("full_grid_A9_64x64")
    L = _lib.load()
    ctx = _lib.context(torch.cuda.current_device())
    B, V, A = c["attributes"].shape
    T = c["triangles"].shape[0]
    out = np.empty_like(c["out"]); dv = np.empty_like(c["d_clip_vertices"]); da = np.empty_like(c["d_attributes"])
    ids = np.empty_like(c["ids"]); bary = np.empty_like(c["bary"]); z = np.empty_like(c["z"])
    p = lambda a: ctypes.c_void_p(a.ctypes.data)
    arrs = [np.ascontiguousarray(c[k]) for k in ("clip_vertices", "attributes", "triangles", "background", "grad_out")]
    rc = L.pmr_rasterize_clip_space_host(ctx, *[p(a) for a in arrs], B, V, T, A, 64, 64, p(out), p(dv), p(da),
                                         p(ids), p(bary), p(z), _lib.BACKWARD_ORDERED, ctypes.c_void_p(0))
    assert rc == 0, L.pmr_last_error(ctx)
    assert_bits(ids, c["ids"], "ids"); assert_bits(out, c["out"], "out")
    assert_bits(dv, c["d_clip_vertices"], "d_clip_vertices"); assert_bits(da, c["d_attributes"], "d_attributes")


def test_vertex_stage_transform(pmr):
    """transform_homogeneous / transform_shared_mesh (camera_utils.py:142-170) against torch fp32 matmul."""
    from pytorch_mesh_renderer_b200 import camera_utils as cu
    from pytorch_mesh_renderer_b200 import synthetic as S
    rng = np.random.default_rng(3)
    B, V = 5, 1000
    m = dev(S.orbit_cameras(B)).requires_grad_(True)
    w = dev(rng.standard_normal((V, 3)).astype(np.float32)).requires_grad_(True)
    g = dev(rng.standard_normal((B, V, 4)).astype(np.float32))
    hom = torch.cat([w, torch.ones_like(w[:, :1])], 1)
    ref = torch.matmul(hom.unsqueeze(0).expand(B, -1, -1), m.transpose(1, 2))
    ref.backward(g)
    ref_dw, ref_dm = w.grad.clone(), m.grad.clone()
    w.grad = None; m.grad = None
    out = cu.transform_shared_mesh(m, w)
    out.backward(g)
    assert_close_t(out, ref, 1e-5); assert_close_t(w.grad, ref_dw, 1e-4); assert_close_t(m.grad, ref_dm, 1e-4)
    w3 = w.detach().unsqueeze(0).expand(B, -1, -1).contiguous().requires_grad_(True)
    out3 = cu.transform_homogeneous(m.detach(), w3)
    out3.backward(g)
    assert_close_t(out3, ref, 1e-5)
    assert_close_t(w3.grad.sum(0), ref_dw, 1e-4)
    out_cpu = cu.transform_homogeneous(m.detach().cpu(), w3.detach().cpu())
    assert out_cpu.device.type == "cpu" and out_cpu.shape == (B, V, 4)


def assert_close_t(a, b, tol):
    a, b = a.detach().float().cpu().numpy(), b.detach().float().cpu().numpy()
    scale = np.abs(b).max() + 1e-30
    assert np.abs(a - b).max() <= tol * scale, (np.abs(a - b).max(), scale)


def _large_triangle_soup(rng, n, spread=1.2):
    """n triangles with vertices spread over (and beyond) the screen: pixel boxes far above 16x16."""
    xy = rng.uniform(-spread, spread, (n, 3, 2))
    z = rng.uniform(-0.9, 0.9, (n, 3, 1))
    w = rng.uniform(0.5, 2.0, (n, 3, 1))
    v = np.concatenate([xy, z, np.ones_like(z)], 2) * w
    return v.reshape(-1, 4).astype(np.float32), np.arange(3 * n, dtype=np.int32).reshape(n, 3)


def test_tile_kernel_many_candidates_per_macro_tile(pmr, oracle):
    """3 000 screen-sized triangles on one 96x80 image: every 64x64 macro tile sees more than the 1 024
    candidates its shared-memory list holds, so the tile kernel works in several passes that merge through
    the global depth keys (raster_tile_kernel, kMacroCap)."""
    from pytorch_mesh_renderer_b200 import ops
    v, t = _large_triangle_soup(np.random.default_rng(11), 3000)
    ids, bary, z = ops.rasterize_forward(dev(v)[None], dev(t), 96, 80)
    ref_ids, ref_bary, ref_z = oracle.forward(v, t, 96, 80)
    assert_bits(ids[0].cpu().numpy(), ref_ids, "ids")
    assert_bits(bary[0].cpu().numpy(), ref_bary, "bary")
    assert_bits(z[0].cpu().numpy(), ref_z, "z")


def test_tile_kernel_more_images_than_one_slice(pmr, oracle):
    """300 small images (the tile kernel lists the images with large triangles in slices of 256), a third of
    them without any triangle that reaches the tile path."""
    from pytorch_mesh_renderer_b200 import ops
    rng = np.random.default_rng(12)
    B, W, H, n = 300, 40, 24, 70
    vs = []
    for b in range(B):
        v, t = _large_triangle_soup(rng, n)
        if b % 3 == 1:
            v[:, :2] *= 0.05                      # only small triangles in this image
        vs.append(v)
    v = np.stack(vs)
    ids, bary, z = ops.rasterize_forward(dev(v), dev(t), W, H)
    ids, bary, z = ids.cpu().numpy(), bary.cpu().numpy(), z.cpu().numpy()
    for b in range(0, B, 7):
        ref_ids, ref_bary, ref_z = oracle.forward(v[b], t, W, H)
        assert_bits(ids[b], ref_ids, "ids of image %d" % b)
        assert_bits(bary[b], ref_bary, "bary of image %d" % b)
        assert_bits(z[b], ref_z, "z of image %d" % b)


def test_forward_and_backward_capture_into_a_cuda_graph(pmr, oracle):
    """No device-pointer entry point waits for the device, so a whole step (fused forward + backward) records
    into a CUDA graph once the context's workspace has its size; replays on new inputs match fresh calls."""
    from pytorch_mesh_renderer_b200 import ops
    from pytorch_mesh_renderer_b200 import synthetic as S
    sc = S.sphere_views(40, 39, 3, 128)                      # 3 040 triangles: the scatter / tile / resolve pipeline
    v, a, t, bg = (dev(sc[k]) for k in ("clip_vertices", "attributes", "triangles", "background"))
    grad = dev(S.upstream_gradient((3, 128, 128, 9)))
    eager = ops.rasterize_interpolate_forward(v, a, t, bg, 128, 128)            # also grows the workspace
    eager_grads = ops.rasterize_interpolate_backward(grad, v, a, t, eager[1], eager[2], "ordered")
    torch.cuda.synchronize()

    static_v = v.clone()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        image, ids, bary, z = ops.rasterize_interpolate_forward(static_v, a, t, bg, 128, 128)
        dv, da = ops.rasterize_interpolate_backward(grad, static_v, a, t, ids, bary, "ordered")
    graph.replay()
    torch.cuda.synchronize()
    for mine, ref, what in ((image, eager[0], "image"), (ids, eager[1], "ids"), (dv, eager_grads[0], "d_vertices"),
                            (da, eager_grads[1], "d_attributes")):
        assert_bits(mine.cpu().numpy(), ref.cpu().numpy(), what)
    # new input through the same graph: a slightly rotated mesh
    c, s_ = np.cos(0.1), np.sin(0.1)
    rot = torch.tensor([[c, -s_, 0, 0], [s_, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=torch.float32, device="cuda")
    v2 = (v @ rot.T).contiguous()
    static_v.copy_(v2)
    graph.replay()
    torch.cuda.synchronize()
    fresh = ops.rasterize_interpolate_forward(v2, a, t, bg, 128, 128)
    fresh_grads = ops.rasterize_interpolate_backward(grad, v2, a, t, fresh[1], fresh[2], "ordered")
    assert_bits(image.cpu().numpy(), fresh[0].cpu().numpy(), "image (replay)")
    assert_bits(dv.cpu().numpy(), fresh_grads[0].cpu().numpy(), "d_vertices (replay)")
    assert_bits(da.cpu().numpy(), fresh_grads[1].cpu().numpy(), "d_attributes (replay)")


@pytest.mark.parametrize("seed", range(24))
def test_random_scenes_vs_oracle(pmr, oracle, threshold, seed):
    """Randomised sweep: image size (odd sizes, single rows / columns), triangle count, triangle scale from
    sub-pixel to screen-filling, shared vertices, mixed-sign and tiny w, degenerate and duplicated triangles,
    attribute counts with and without a specialised kernel -- everything bit-exact against the oracle
    (forward buffers, image, ORDERED gradients), ATOMIC gradients within the summation-order bound."""
    _check_random_scene(pmr, oracle, seed)


# Sizes that the tensor-map paths take with partly filled boxes: W a multiple of 4 but not of 8 (half a block at the
# right edge, clipped by the copy engine), strips that end inside the image, H a multiple of 4 or not (tensor stores
# of the resolve kernel or per-row copies), a single block row.
@pytest.mark.parametrize("size", [(36, 20), (100, 52), (12, 8), (44, 64), (68, 36), (20, 4), (132, 16), (260, 30)],
                         ids=lambda s: "%dx%d" % s)
def test_random_scenes_partial_boxes(pmr, oracle, size):
    _check_random_scene(pmr, oracle, 500 + size[0], size)


def _check_random_scene(pmr, oracle, seed, size=None):
    rng = np.random.default_rng(1000 + seed)
    W, H = int(rng.choice([1, 7, 16, 33, 64, 97, 130])), int(rng.choice([1, 5, 16, 31, 64, 75, 128]))
    if size is not None:
        W, H = size
    B = int(rng.integers(1, 4))
    A = int(rng.choice([1, 3, 4, 7, 9, 12, 13]))
    kind = seed % 3
    if kind == 0:        # soup of independent triangles, from sub-pixel to screen-filling
        T = int(rng.integers(1, 900))
        V = 3 * T
        scale = float(rng.choice([0.01, 0.03, 0.1, 0.5, 2.0]))
        centre = rng.uniform(-1.1, 1.1, (B, T, 1, 2))
        xy = (centre + scale * rng.standard_normal((B, T, 3, 2))).reshape(B, V, 2)
        tris = np.arange(V).reshape(T, 3)
    elif kind == 1:      # jittered grid mesh: small triangles that share vertices and edges
        n = int(rng.integers(3, 22))
        V = n * n
        gy, gx = np.meshgrid(np.linspace(-1.05, 1.05, n), np.linspace(-1.05, 1.05, n), indexing="ij")
        xy = np.stack([gx, gy], 2).reshape(1, V, 2) + (0.6 / n) * rng.standard_normal((B, V, 2))
        i, j = np.meshgrid(np.arange(n - 1), np.arange(n - 1), indexing="ij")
        q = (i * n + j).reshape(-1)
        tris = np.concatenate([np.stack([q, q + 1, q + n], 1), np.stack([q + 1, q + n + 1, q + n], 1)])
        rng.shuffle(tris)
        T = tris.shape[0]
    else:                # random index triples over random vertices: large overlapping triangles
        V = int(rng.integers(3, 400))
        T = int(rng.integers(1, 600))
        xy = rng.uniform(-1.3, 1.3, (B, V, 2))
        tris = rng.integers(0, V, (T, 3))
    z = rng.uniform(-1.2, 1.2, (B, V, 1))                       # some depths outside [-1, 1]
    w = rng.uniform(0.3, 2.5, (B, V, 1))
    if seed % 4 == 0:
        w[rng.random((B, V, 1)) < 0.1] *= -1.0                  # vertices behind the eye
    if seed % 4 == 1:
        w[rng.random((B, V, 1)) < 0.05] = 1e-3                  # tiny w: huge projected coordinates
    verts = (np.concatenate([xy, z, np.ones_like(z)], 2) * w).astype(np.float32)
    tris = np.array(tris)
    degenerate = rng.random(T) < 0.03
    tris[degenerate, 2] = tris[degenerate, 0]                   # zero-area triangles
    if T > 4:
        tris[T - 1] = tris[0]                                   # an exact duplicate: equal depth, larger id wins
    tris = tris.astype(np.int32)
    attrs = rng.standard_normal((B, V, A)).astype(np.float32)
    bg = rng.standard_normal(A).astype(np.float32)
    g = rng.standard_normal((B, H, W, A)).astype(np.float32)

    ref = oracle.rasterize_clip_space(verts, attrs, tris, W, H, bg, grad_out=g, f64_yardstick=True)
    cv, at = dev(verts).requires_grad_(True), dev(attrs).requires_grad_(True)
    with pmr.backward_mode("ordered"):
        out, (ids, bary, zb) = pmr.rasterize_clip_space(cv, at, dev(tris), W, H, dev(bg), return_buffers=True)
        out.backward(dev(g))
    assert_bits(ids.cpu().numpy(), ref["ids"], "ids")
    assert_bits(bary.detach().cpu().numpy(), ref["bary"], "bary")
    assert_bits(zb.detach().cpu().numpy(), ref["z"], "z")
    assert_bits(out.detach().cpu().numpy(), ref["out"], "out")
    assert_bits(cv.grad.cpu().numpy(), ref["d_vertices"], "d_vertices (ordered)")
    assert_bits(at.grad.cpu().numpy(), ref["d_attributes"], "d_attributes (ordered)")
    cv2, at2 = dev(verts).requires_grad_(True), dev(attrs).requires_grad_(True)
    with pmr.backward_mode("atomic"):
        pmr.rasterize_clip_space(cv2, at2, dev(tris), W, H, dev(bg)).backward(dev(g))
    atomic_close(at2.grad.cpu().numpy(), ref["d_attributes"], ref["d_attributes_f64"], "d_attributes (atomic)")
    atomic_close(cv2.grad.cpu().numpy(), ref["d_vertices"], ref["d_vertices_f64"], "d_vertices (atomic)")
