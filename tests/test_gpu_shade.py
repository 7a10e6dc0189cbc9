"""shade_diffuse (csrc/shade.cu): the fused diffuse + ambient Phong kernels against

* the torch-op `phong_shader` mirror of render.py:231-386 (itself checked against outputs of the unmodified
  reference in test_gpu_render.py) on random attribute images, forward and autograd gradients, and
* outputs of the unmodified reference's `render` (tests/golden/render_cube_96x72.npz), through `render`.

Lighting is float arithmetic through different libraries, so values are compared within the north-star
tolerance 1e-6 absolute + 1e-5 relative; gradients additionally relative to their largest magnitude
(sums of many terms)."""
import numpy as np
import pytest

from conftest import load_golden

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pmr():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pytorch_mesh_renderer_b200 as m
    return m


def _random_pixels(B, H, W, A, seed, background_fraction=0.3):
    g = torch.Generator().manual_seed(seed)
    px = torch.randn((B, H, W, A), generator=g)
    px[..., 6:9] = torch.rand((B, H, W, 3), generator=g)                  # diffuse colours in [0, 1)
    bg = torch.rand((B, H, W), generator=g) < background_fraction
    px[bg] = -1.0                                                         # what the rasterizer writes there
    px[0, 0, 0, 0:3] = 0.0                                                # a zero normal (normalize eps path)
    return px


def _mirror(pmr, px, lp, li, ambient):
    import torch.nn.functional as F
    from pytorch_mesh_renderer_b200.render import phong_shader
    normals = F.normalize(px[..., 0:3], p=2, dim=3)
    mask = (px[..., 6:9] >= 0.0).any(dim=3).to(torch.float32)
    return phong_shader(normals, mask, px[..., 3:6], lp, li, px[..., 6:9], ambient_color=ambient)


@pytest.mark.parametrize("B,H,W,A,L,use_ambient", [(2, 37, 53, 9, 1, False), (3, 16, 40, 9, 3, True),
                                                  (1, 64, 64, 13, 2, True), (2, 8, 8, 9, 16, False)])
def test_shade_diffuse_matches_torch_mirror(pmr, B, H, W, A, L, use_ambient):
    from pytorch_mesh_renderer_b200.render import shade_diffuse
    g = torch.Generator().manual_seed(100 + L)
    px = _random_pixels(B, H, W, A, seed=7 + A).cuda()
    lp = (3.0 * torch.randn((B, L, 3), generator=g)).cuda()
    li = torch.rand((B, L, 3), generator=g).cuda()
    ambient = torch.rand((B, 3), generator=g).cuda() if use_ambient else None
    grad = torch.randn((B, H, W, 4), generator=g).cuda()

    a = px.clone().requires_grad_(True)
    out = shade_diffuse(a, lp, li, ambient)
    out.backward(grad)
    b = px.clone().requires_grad_(True)
    ref = _mirror(pmr, b, lp, li, ambient)
    ref.backward(grad)

    o, r = out.detach().cpu().numpy(), ref.detach().cpu().numpy()
    assert np.array_equal(o[..., 3], r[..., 3])                            # mask identical
    assert (np.abs(o - r) <= 1e-6 + 1e-5 * np.abs(r)).all(), np.abs(o - r).max()
    go, gr = a.grad.cpu().numpy(), b.grad.cpu().numpy()
    assert (go[..., 9:] == 0).all()
    # the zero-normal pixel: torch's normalize backward divides 0/0 there; compare everything else
    go[0, 0, 0, 0:3] = gr[0, 0, 0, 0:3] = 0.0
    err = np.abs(go - gr)
    assert (err <= 1e-6 + 1e-5 * np.abs(gr) + 2e-6 * np.abs(gr).max()).all(), (err.max(), np.abs(gr).max())


def test_render_uses_the_fused_shader_and_matches_reference(pmr):
    """render() without specular goes through shade_diffuse; same goldens as test_gpu_render.py."""
    from pytorch_mesh_renderer_b200 import _lib
    c = load_golden("render_cube_96x72")
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    v = dev(c["vertices"]).requires_grad_(True)
    n = dev(c["normals"]).requires_grad_(True)
    d = dev(c["diffuse"]).requires_grad_(True)
    index = torch.cuda.current_device()
    _lib.enable_stage_timing(index, True)
    _lib.read_stage_timing(index, reset=True)
    with pmr.backward_mode("ordered"):
        out = pmr.render(v, dev(c["triangles"]), n, d, dev(c["eye"]), dev(c["center"]), dev(c["up"]),
                         dev(c["light_positions"]), dev(c["light_intensities"]), int(c["width"]), int(c["height"]))
        out.backward(dev(c["grad_out"]))
    stages = _lib.read_stage_timing(index, reset=True)
    _lib.enable_stage_timing(index, False)
    assert stages["shade"][1] == 2                                         # one forward, one backward launch
    img = out.detach().cpu().numpy()
    assert np.array_equal(img[..., 3], c["image"][..., 3])
    assert (np.abs(img - c["image"]) <= 1e-5 + 1e-4 * np.abs(c["image"])).all()
    for mine, key in ((v.grad, "d_vertices"), (n.grad, "d_normals"), (d.grad, "d_diffuse")):
        ref = c[key]
        assert np.abs(mine.cpu().numpy() - ref).max() <= 2e-4 * (np.abs(ref).max() + 1e-12), key


def _mirror_specular(pmr, px, lp, li, ambient, camera, shininess):
    import torch.nn.functional as F
    from pytorch_mesh_renderer_b200.render import phong_shader
    normals = F.normalize(px[..., 0:3], p=2, dim=3)
    mask = (px[..., 6:9] >= 0.0).any(dim=3).to(torch.float32)
    shin = px[..., 12] if px.shape[3] > 12 else shininess.reshape(-1, 1, 1)
    return phong_shader(normals, mask, px[..., 3:6], lp, li, px[..., 6:9], camera, px[..., 9:12], shin, ambient)


@pytest.mark.parametrize("B,H,W,A,L,use_ambient", [(2, 37, 53, 12, 1, False), (3, 16, 40, 13, 3, True),
                                                  (1, 64, 64, 13, 2, False), (2, 24, 24, 12, 4, True)])
def test_shade_phong_matches_torch_mirror(pmr, B, H, W, A, L, use_ambient):
    """Specular branch: two passes each way; the per-(image, light) sums are float atomics, so values are
    compared by tolerance (1e-6 + 1e-5 relative; gradients also relative to their largest magnitude)."""
    from pytorch_mesh_renderer_b200.render import shade_phong
    g = torch.Generator().manual_seed(200 + L)
    px = _random_pixels(B, H, W, A, seed=17 + A)
    px[..., 9:12] = torch.rand((B, H, W, 3), generator=g)                 # specular colours
    if A > 12:
        px[..., 12] = 1.0 + 9.0 * torch.rand((B, H, W), generator=g)      # per-pixel shininess in [1, 10)
    bg = px[..., 6] < 0
    px[bg] = -1.0
    px = px.cuda()
    lp = (3.0 * torch.randn((B, L, 3), generator=g)).cuda()
    li = torch.rand((B, L, 3), generator=g).cuda()
    cam = (4.0 * torch.randn((B, 3), generator=g)).cuda()
    ambient = torch.rand((B, 3), generator=g).cuda() if use_ambient else None
    shininess = (2.0 + 6.0 * torch.rand((B,), generator=g)).cuda() if A == 12 else None
    grad = torch.randn((B, H, W, 4), generator=g).cuda()

    a = px.clone().requires_grad_(True)
    out = shade_phong(a, lp, li, cam, ambient, shininess)
    out.backward(grad)
    b = px.clone().requires_grad_(True)
    ref = _mirror_specular(pmr, b, lp, li, ambient, cam, shininess)
    ref.backward(grad)

    o, r = out.detach().cpu().numpy(), ref.detach().cpu().numpy()
    assert np.array_equal(o[..., 3], r[..., 3])
    assert (np.abs(o - r) <= 1e-6 + 1e-5 * np.abs(r)).all(), np.abs(o - r).max()
    go, gr = a.grad.cpu().numpy(), b.grad.cpu().numpy()
    go[0, 0, 0, 0:3] = gr[0, 0, 0, 0:3] = 0.0                             # the zero-normal pixel (0/0 in torch)
    finite = np.isfinite(gr)
    assert (np.isfinite(go) == finite).all()
    err = np.abs(go - gr)[finite]
    scale = np.abs(gr[finite]).max()
    assert (err <= 1e-6 + 1e-5 * np.abs(gr[finite]) + 5e-6 * scale).all(), (err.max(), scale)


@pytest.mark.parametrize("name", ["render_complex_vertex_shininess_96x72", "render_complex_scalar_shininess_96x72"])
def test_render_specular_uses_the_fused_shader_and_matches_reference(pmr, name):
    from pytorch_mesh_renderer_b200 import _lib
    c = load_golden(name)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    v = dev(c["vertices"]).requires_grad_(True)
    n = dev(c["normals"]).requires_grad_(True)
    d = dev(c["diffuse"]).requires_grad_(True)
    extra = {}
    for k in c:
        if k.startswith("arg_"):
            val = c[k]
            extra[k[4:]] = dev(val) if val.ndim > 0 else torch.tensor(float(val))
    index = torch.cuda.current_device()
    _lib.enable_stage_timing(index, True)
    _lib.read_stage_timing(index, reset=True)
    with pmr.backward_mode("ordered"):
        out = pmr.render(v, dev(c["triangles"]), n, d, dev(c["eye"]), dev(c["center"]), dev(c["up"]),
                         dev(c["light_positions"]), dev(c["light_intensities"]), int(c["width"]), int(c["height"]), **extra)
        out.backward(dev(c["grad_out"]))
    stages = _lib.read_stage_timing(index, reset=True)
    _lib.enable_stage_timing(index, False)
    assert stages["shade"][1] == 2                                         # the fused specular path ran (fwd + bwd)
    img = out.detach().cpu().numpy()
    assert np.array_equal(img[..., 3], c["image"][..., 3])
    assert (np.abs(img - c["image"]) <= 1e-5 + 1e-4 * np.abs(c["image"])).all()
    for mine, key in ((v.grad, "d_vertices"), (n.grad, "d_normals"), (d.grad, "d_diffuse")):
        ref = c[key]
        assert np.abs(mine.cpu().numpy() - ref).max() <= 2e-4 * (np.abs(ref).max() + 1e-12), key


def _scene_with_normals(B, size, n_lon=48, n_rings=47):
    from pytorch_mesh_renderer_b200 import synthetic as S
    sc = S.sphere_views(n_lon, n_rings, B, size)
    world = sc["world_vertices"].astype(np.float32)
    normals = world / np.linalg.norm(world, axis=1, keepdims=True)
    rng = np.random.default_rng(5)
    diffuse = rng.random(world.shape, dtype=np.float32)
    attrs = np.concatenate([normals, world, diffuse], 1)[None].repeat(B, 0).astype(np.float32)
    return sc, attrs


@pytest.mark.parametrize("size,use_ambient,L", [(96, False, 1), (131, True, 3)])
def test_fused_render_path_equals_rasterize_then_shade(pmr, size, use_ambient, L):
    """pmr_render_diffuse_*: lighting inside the resolve / backward kernels.  Forward: RGBA, ids, barycentrics
    bit-identical to rasterize_clip_space followed by shade_diffuse.  Backward (atomic accumulation): against
    the unfused ORDERED gradients, relative to their largest magnitude."""
    from pytorch_mesh_renderer_b200 import ops
    from pytorch_mesh_renderer_b200.render import render_diffuse_clip_space, shade_diffuse
    B = 3
    sc, attrs = _scene_with_normals(B, size)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    g = torch.Generator().manual_seed(9)
    lp = (3.0 * torch.randn((B, L, 3), generator=g)).cuda()
    li = torch.rand((B, L, 3), generator=g).cuda()
    ambient = torch.rand((B, 3), generator=g).cuda() if use_ambient else None
    grad = torch.randn((B, size, size, 4), generator=g).cuda()
    tris = dev(sc["triangles"])
    bg = torch.full((9,), -1.0, device="cuda")

    cv = dev(sc["clip_vertices"]).requires_grad_(True)
    at = dev(attrs).requires_grad_(True)
    fused = render_diffuse_clip_space(cv, at, tris, lp, li, size, size, ambient)
    fused.backward(grad)

    cv2 = dev(sc["clip_vertices"]).requires_grad_(True)
    at2 = dev(attrs).requires_grad_(True)
    with pmr.backward_mode("ordered"):
        pixels, (ids, bary, z) = pmr.rasterize_clip_space(cv2, at2, tris, size, size, bg, return_buffers=True)
        unfused = shade_diffuse(pixels, lp, li, ambient)
        unfused.backward(grad)

    assert torch.equal(fused.detach(), unfused.detach())                     # bit-identical image
    f_rgba, f_ids, f_bary, f_z = ops.render_diffuse_forward(cv.detach(), at.detach(), tris, bg, lp, li, ambient, size, size)
    assert torch.equal(f_ids, ids) and torch.equal(f_bary, bary.detach()) and torch.equal(f_z, z.detach())
    for mine, ref, what in ((cv.grad, cv2.grad, "d_clip_vertices"), (at.grad, at2.grad, "d_attributes")):
        mine, ref = mine.cpu().numpy(), ref.cpu().numpy()
        assert np.abs(mine - ref).max() <= 2e-5 * (np.abs(ref).max() + 1e-12), (what, np.abs(mine - ref).max(), np.abs(ref).max())


def test_render_default_mode_takes_the_fused_path_and_matches_reference(pmr):
    """render() in the default (atomic) mode: scatter / resolve+shade forward, one backward kernel; against the
    unmodified reference's outputs (tests/golden/render_cube_96x72.npz)."""
    from pytorch_mesh_renderer_b200 import _lib
    c = load_golden("render_cube_96x72")
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    v = dev(c["vertices"]).requires_grad_(True)
    n = dev(c["normals"]).requires_grad_(True)
    d = dev(c["diffuse"]).requires_grad_(True)
    index = torch.cuda.current_device()
    _lib.enable_stage_timing(index, True)
    _lib.read_stage_timing(index, reset=True)
    out = pmr.render(v, dev(c["triangles"]), n, d, dev(c["eye"]), dev(c["center"]), dev(c["up"]),
                     dev(c["light_positions"]), dev(c["light_intensities"]), int(c["width"]), int(c["height"]))
    out.backward(dev(c["grad_out"]))
    stages = _lib.read_stage_timing(index, reset=True)
    _lib.enable_stage_timing(index, False)
    assert stages["shade"][1] == 0 and stages["resolve"][1] == 1 and stages["backward"][1] == 1
    img = out.detach().cpu().numpy()
    assert np.array_equal(img[..., 3], c["image"][..., 3])
    assert (np.abs(img - c["image"]) <= 1e-5 + 1e-4 * np.abs(c["image"])).all()
    for mine, key in ((v.grad, "d_vertices"), (n.grad, "d_normals"), (d.grad, "d_diffuse")):
        ref = c[key]
        assert np.abs(mine.cpu().numpy() - ref).max() <= 2e-4 * (np.abs(ref).max() + 1e-12), key
