"""The `render` caller (reference render.py:16-228) on the CUDA path: against outputs of the unmodified
reference (tests/golden/render_*.npz from make_golden_render.py), the reference's PNG fixtures
Gray_Cube_{0,1}.png (mesh_renderer_test.py:30-70) and the cube-rotation fit of :204-271.

Lighting is float arithmetic through different libraries (torch CUDA here, torch CPU in the reference),
so images are compared within 1e-5 absolute / 1e-4 relative; gradients pass through sums over thousands
of pixels and are compared relative to their largest magnitude."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, load_golden

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pmr():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pytorch_mesh_renderer_b200 as m
    return m


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _near_png(image, name, max_outlier_fraction=0.001, threshold=0.01):
    from PIL import Image
    png = np.asarray(Image.open(os.path.join(GOLDEN_DIR, "reference_png", name))).astype(np.float64) / 255.0
    assert image.shape == png.shape
    diff = np.abs(png - np.clip(image, 0.0, 1.0))
    return np.any(diff > threshold, axis=2).mean() <= max_outlier_fraction


@pytest.mark.parametrize("name", ["render_cube_96x72", "render_complex_vertex_shininess_96x72",
                                  "render_complex_scalar_shininess_96x72"])
def test_render_matches_reference(pmr, name):
    c = load_golden(name)
    v = dev(c["vertices"]).requires_grad_(True)
    n = dev(c["normals"]).requires_grad_(True)
    d = dev(c["diffuse"]).requires_grad_(True)
    extra = {}
    for k in c:
        if k.startswith("arg_"):
            val = c[k]
            extra[k[4:]] = dev(val) if val.ndim > 0 else torch.tensor(float(val))
    with pmr.backward_mode("ordered"):
        out = pmr.render(v, dev(c["triangles"]), n, d, dev(c["eye"]), dev(c["center"]), dev(c["up"]),
                         dev(c["light_positions"]), dev(c["light_intensities"]), int(c["width"]), int(c["height"]), **extra)
        out.backward(dev(c["grad_out"]))
    img = out.detach().cpu().numpy()
    assert img.shape == c["image"].shape
    assert np.array_equal(img[..., 3], c["image"][..., 3])                 # coverage mask identical
    err = np.abs(img - c["image"])
    assert (err <= 1e-5 + 1e-4 * np.abs(c["image"])).all(), err.max()
    for mine, key in ((v.grad, "d_vertices"), (n.grad, "d_normals"), (d.grad, "d_diffuse")):
        ref = c[key]
        assert np.abs(mine.cpu().numpy() - ref).max() <= 2e-4 * (np.abs(ref).max() + 1e-12), key


def test_gray_cube_png_fixtures(pmr):
    """mesh_renderer_test.py:30-70 testRendersSimpleCube, CPU tensors in and out like the reference test."""
    from pytorch_mesh_renderer_b200 import camera_utils as cu
    cube = torch.tensor([[-1, -1, 1], [-1, -1, -1], [-1, 1, -1], [-1, 1, 1], [1, -1, 1],
                         [1, -1, -1], [1, 1, -1], [1, 1, 1]], dtype=torch.float32)
    normals = torch.nn.functional.normalize(cube, dim=1, p=2)
    tris = torch.tensor([[0, 1, 2], [2, 3, 0], [3, 2, 6], [6, 7, 3], [7, 6, 5], [5, 4, 7],
                         [4, 5, 1], [1, 0, 4], [5, 6, 2], [2, 1, 5], [7, 4, 0], [0, 3, 7]], dtype=torch.int32)
    rot = cu.euler_matrices(torch.tensor([[-20.0, 0.0, 60.0], [45.0, 60.0, 0.0]]))[:, :3, :3]
    v = torch.matmul(torch.stack([cube, cube]), rot.transpose(1, 2))
    n = torch.matmul(torch.stack([normals, normals]), rot.transpose(1, 2))
    eye = torch.tensor(2 * [[0.0, 0.0, 6.0]]); center = torch.zeros(2, 3); up = torch.tensor(2 * [[0.0, 1.0, 0.0]])
    images = pmr.render(v, tris, n, torch.ones_like(v), eye, center, up,
                        torch.tensor([[[0.0, 0.0, 6.0]], [[0.0, 0.0, 6.0]]]), torch.ones(2, 1, 3), 640, 480)
    assert images.device.type == "cpu" and images.shape == (2, 480, 640, 4)
    for i in (0, 1):
        assert _near_png(images[i].numpy(), "Gray_Cube_%d.png" % i)


def test_cube_rotation_fit(pmr):
    """mesh_renderer_test.py:204-271 testThatCubeRotates: 35 SGD steps on the Euler angles through
    render() recover the target view (<= 1 % of pixels off by more than 0.04 against Gray_Cube_0.png)."""
    from pytorch_mesh_renderer_b200 import camera_utils as cu
    device = torch.device("cuda")
    cube = torch.tensor([[-1, -1, 1], [-1, -1, -1], [-1, 1, -1], [-1, 1, 1], [1, -1, 1],
                         [1, -1, -1], [1, 1, -1], [1, 1, 1]], dtype=torch.float32, device=device)
    normals = torch.nn.functional.normalize(cube, dim=1, p=2)
    tris = torch.tensor([[0, 1, 2], [2, 3, 0], [3, 2, 6], [6, 7, 3], [7, 6, 5], [5, 4, 7],
                         [4, 5, 1], [1, 0, 4], [5, 6, 2], [2, 1, 5], [7, 4, 0], [0, 3, 7]], dtype=torch.int32, device=device)
    eye = torch.tensor([[0.0, 0.0, 6.0]], device=device)

    def render_with(angles):
        rot = cu.euler_matrices(angles)[0, :3, :3]
        v = torch.matmul(cube, rot.T).reshape(1, 8, 3)
        n = torch.matmul(normals, rot.T).reshape(1, 8, 3)
        out = pmr.render(v, tris, n, torch.ones_like(v), eye, torch.zeros(1, 3, device=device),
                         torch.tensor([[0.0, 1.0, 0.0]], device=device), eye.reshape(1, 1, 3),
                         torch.ones(1, 1, 3, device=device), 640, 480)
        return out.reshape(480, 640, 4)

    target = render_with(torch.tensor([[-20.0, 0.0, 60.0]], device=device)).detach()
    assert _near_png(target.cpu().numpy(), "Gray_Cube_0.png")
    angles = torch.zeros(1, 3, device=device, requires_grad=True)
    opt = torch.optim.SGD([angles], 0.7, 0.1)

    def closure():
        opt.zero_grad()
        loss = torch.mean(torch.abs(render_with(angles) - target))
        loss.backward()
        torch.nn.utils.clip_grad_norm_([angles], 1.0)
        return loss

    for _ in range(35):
        opt.step(closure)
    final = render_with(angles).detach().cpu().numpy()
    assert _near_png(final, "Gray_Cube_0.png", max_outlier_fraction=0.01, threshold=0.04)
