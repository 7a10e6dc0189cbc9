"""The reference's own Python layers (rasterize_triangles_ext.py, rasterize.py: UNMODIFIED, imported from
/root/reference or its verbatim staging oracle/_ref/pysrc) running on this library through the
`rasterize_triangles_cpp` drop-in module (SURVEY 8 a7 / b): ext.py:3 `import rasterize_triangles_cpp` resolves to
pytorch_mesh_renderer_b200.rasterize_triangles_cpp, ext.py:39 / :56 call its forward / backward."""
import numpy as np
import pytest
import torch

from conftest import assert_bits, golden_names, grad_from_seed, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def reference_on_b200():
    from oracle import reference_harness as rh
    import pytorch_mesh_renderer_b200.rasterize_triangles_cpp as shim
    if not rh.available():
        pytest.skip("reference Python layers not staged (run oracle/build_ref.py in the build container)")
    rast, ext = rh.import_reference(shim)
    yield rast, ext
    rh.import_reference(rh.kernel())        # leave the harness as the oracle tests expect it


def _cases(*keys):
    names = []
    for name in golden_names(""):
        g = load_golden(name)
        if all(k in g for k in keys):
            names.append(name)
    return names


@pytest.mark.parametrize("name", _cases("vertices", "ids", "bary", "z", "df_dbary", "df_dvertices"))
def test_reference_autograd_function_on_the_dropin(reference_on_b200, name):
    """BarycentricRasterizer.apply of the reference (ext.py:6-63) with CUDA tensors: ids, barycentrics, z and
    d(vertices) bit-equal to what the reference's own kernel produced (tests/golden/make_golden.py)."""
    _, ext = reference_on_b200
    g = load_golden(name)
    W, H = int(g["width"]), int(g["height"])
    v = torch.from_numpy(g["vertices"]).cuda().requires_grad_(True)
    t = torch.from_numpy(g["triangles"]).cuda()
    ids, bary, z = ext.BarycentricRasterizer.apply(v, t, W, H)
    assert ids.is_cuda and bary.is_cuda and z.is_cuda
    assert_bits(ids.detach().cpu().numpy(), g["ids"], name + " ids")
    assert_bits(bary.detach().cpu().numpy(), g["bary"], name + " bary")
    assert_bits(z.detach().cpu().numpy(), g["z"], name + " z")
    bary.backward(torch.from_numpy(g["df_dbary"]).cuda())
    assert_bits(v.grad.cpu().numpy(), g["df_dvertices"], name + " df_dvertices (reference order)")


@pytest.mark.parametrize("name", _cases("clip_vertices", "attributes", "out", "grad_out", "d_clip_vertices", "d_attributes"))
def test_reference_rasterize_clip_space_on_the_dropin(reference_on_b200, name):
    """The reference's rasterize_clip_space (rast.py:66-152: its Python loop over images, index_select, gather,
    mul, sum, clamp, blend) fed CUDA tensors, the kernel calls landing in libpmr_b200.so: image within the
    north-star tolerance of the reference's own run on its CPU kernel (the goldens; torch's CUDA ops around the
    kernel need not round like its CPU ops), gradients within 2e-5 of the largest entry (d(attributes) comes from
    torch's index_put_ on the GPU here: atomics)."""
    from conftest import assert_close
    rast, _ = reference_on_b200
    g = load_golden(name)
    W, H = int(g["width"]), int(g["height"])
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    cv = dev(g["clip_vertices"]).requires_grad_(True)
    at = dev(g["attributes"]).requires_grad_(True)
    out = rast.rasterize_clip_space(cv, at, dev(g["triangles"]), W, H, dev(g["background"]))
    assert out.is_cuda
    out.backward(dev(g["grad_out"]))
    # torch's CUDA mul / sum kernels need not round like its CPU kernels (one attribute: a different order of
    # the three corner products); ids and barycentrics underneath are bit-exact (test above)
    assert_close(out.detach().cpu().numpy(), g["out"], name + ": image through the reference's torch ops")
    # d(bary) is reduced over the attribute axis by torch's CUDA sum here and by its CPU sum in the golden run:
    # same terms, different order
    scale = np.abs(g["d_clip_vertices"]).max() + 1e-30
    err = np.abs(cv.grad.cpu().numpy() - g["d_clip_vertices"]).max()
    assert err <= 2e-5 * scale, (name, err, scale)
    assert np.array_equal(cv.grad.cpu().numpy()[..., 2], np.zeros_like(g["d_clip_vertices"][..., 2]))
    err = np.abs(at.grad.cpu().numpy() - g["d_attributes"]).max()
    assert err <= 2e-5 * (np.abs(g["d_attributes"]).max() + 1e-30), (name, err)


def test_dropin_module_surface():
    """K.cpp:421-424 exports exactly forward and backward; wrong scalar types raise RuntimeError like accessor<>."""
    import pytorch_mesh_renderer_b200.rasterize_triangles_cpp as shim
    assert sorted(shim.__all__) == ["backward", "forward"]
    v = torch.zeros((3, 4), dtype=torch.float64, device="cuda")
    t = torch.zeros((1, 3), dtype=torch.int32, device="cuda")
    with pytest.raises(RuntimeError, match="expected scalar type Float but found"):
        shim.forward(v, t, 8, 8)
    with pytest.raises(RuntimeError, match="expected scalar type Int but found"):
        shim.forward(v.float(), t.long(), 8, 8)
    ids, bary, z = shim.forward(torch.tensor([[-.5, -.5, .8, 1], [0, .5, .3, 1], [.5, -.5, .3, 1]]), t.cpu(), 16, 12)
    assert not ids.is_cuda and tuple(ids.shape) == (12, 16) and tuple(bary.shape) == (12, 16, 3) and tuple(z.shape) == (12, 16)
