"""Vertex normals on the CUDA path (csrc/mesh_normals.cu behind pmr_vertex_*; reference src/common/meshes.py:3-35):
the topology table, the forward pass bit for bit against the oracle and the reference's golden outputs, the
backward pass against the reference's autograd gradients and a float64 torch mirror, and load_obj on a file
without normals.  Tolerances: forward bit-exact; gradients 1e-6 + 1e-5 |ref| relative to the largest gradient of
the case (they are sums of products whose order differs from autograd's)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, assert_bits, golden_names, load_golden

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

NORMAL_CASES = golden_names("mesh_normals_")


@pytest.fixture(scope="module")
def meshes():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pytorch_mesh_renderer_b200 import meshes as m
    return m


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def mirror_normals(vertices, triangles):
    """meshes.py:19-34 with torch ops (explicit dim), any dtype/device; the checker for gradients."""
    tri = triangles.long()
    normals = torch.zeros_like(vertices)
    for b in range(vertices.shape[0]):
        vf = vertices[b, tri, :]
        for c in range(3):
            normals[b].index_add_(0, tri[:, c], torch.cross(vf[:, (c + 1) % 3] - vf[:, c],
                                                            vf[:, (c + 2) % 3] - vf[:, c], dim=-1))
    return torch.nn.functional.normalize(normals, eps=1e-6, p=2, dim=-1)


def expected_incidence(triangles, V):
    t = np.asarray(triangles)
    T = t.shape[0]
    flat = t.T.reshape(-1)                                   # corner-major: the order of the reference's passes
    codes = (np.repeat(np.arange(3), T).astype(np.int64) << 30) | np.tile(np.arange(T), 3)
    keep = (flat >= 0) & (flat < V)
    order = np.argsort(flat[keep], kind="stable")
    offsets = np.concatenate([[0], np.cumsum(np.bincount(flat[keep], minlength=V))])
    return offsets.astype(np.int32), codes[keep][order].astype(np.int32)


@pytest.mark.parametrize("name", NORMAL_CASES)
def test_incidence_table(meshes, name):
    from pytorch_mesh_renderer_b200 import ops
    g = load_golden(name)
    V = g["vertices"].shape[1]
    offsets, incidence = ops.vertex_incidence(dev(g["triangles"]), V)
    want_offsets, want_incidence = expected_incidence(g["triangles"], V)
    assert_bits(offsets.cpu().numpy(), want_offsets, name)
    assert_bits(incidence.cpu().numpy(), want_incidence, name)


def test_incidence_table_large_and_out_of_range(meshes):
    """More vertices than one scan chunk, high-valence poles, ids outside [0, V) skipped, empty topology."""
    from pytorch_mesh_renderer_b200 import ops, shapes
    _, tris, _ = shapes.sphere(1.0, 70)                       # V = 4902, poles with 70 incident corners
    V = 4902
    t = tris.numpy().copy()
    t[5, 1] = V + 3
    t[9, 0] = -1
    offsets, incidence = ops.vertex_incidence(dev(t), V)
    want_offsets, want_incidence = expected_incidence(t, V)
    assert_bits(offsets.cpu().numpy(), want_offsets)
    assert_bits(incidence.cpu().numpy()[:want_offsets[-1]], want_incidence)
    offsets, incidence = ops.vertex_incidence(torch.zeros((0, 3), dtype=torch.int32, device="cuda"), 7)
    assert offsets.cpu().tolist() == [0] * 8 and incidence.numel() == 0


@pytest.mark.parametrize("name", NORMAL_CASES)
def test_forward_bits_and_backward_against_reference(meshes, oracle, name):
    g = load_golden(name)
    v = dev(g["vertices"]).requires_grad_(True)
    normals = meshes.compute_vertex_normals(v, dev(g["triangles"]))
    assert_bits(normals.detach().cpu().numpy(), g["normals"], name + " vs reference")
    assert_bits(normals.detach().cpu().numpy(), oracle.vertex_normals(g["vertices"], g["triangles"]), name + " vs oracle")
    normals.backward(dev(g["grad_normals"]))
    ref = g["d_vertices"]
    err = np.abs(v.grad.cpu().numpy() - ref)
    assert (err <= 1e-6 + 1e-5 * np.abs(ref) + 1e-5 * np.abs(ref).max()).all(), (name, err.max(), np.abs(ref).max())


def test_gradient_against_float64_mirror(meshes):
    """Random closed mesh: CUDA fp32 gradient vs the torch-op mirror evaluated in float64."""
    from pytorch_mesh_renderer_b200 import shapes
    sv, st, _ = shapes.sphere(1.0, 16)
    rng = np.random.default_rng(3)
    verts = (sv.numpy()[None] + 0.03 * rng.standard_normal((4,) + tuple(sv.shape))).astype(np.float32)
    grad = rng.standard_normal(verts.shape).astype(np.float32)
    v = dev(verts).requires_grad_(True)
    meshes.compute_vertex_normals(v, st.cuda()).backward(dev(grad))
    v64 = torch.from_numpy(verts).double().requires_grad_(True)
    mirror_normals(v64, st).backward(torch.from_numpy(grad).double())
    ref = v64.grad.numpy()
    err = np.abs(v.grad.cpu().numpy() - ref)
    assert err.max() <= 2e-5 * np.abs(ref).max(), (err.max(), np.abs(ref).max())


def test_full_size_sphere_matches_oracle(meshes, oracle):
    """The c2 mesh (25 124 vertices, 50 244 triangles), 8 perturbed copies: bit-exact against the oracle, unit
    length, deterministic, gradient orthogonal to rigid translation."""
    from pytorch_mesh_renderer_b200 import synthetic
    verts, tris = synthetic.uv_sphere(159, 158)
    verts = np.asarray(verts, np.float32)
    tris = np.asarray(tris, np.int32)
    rng = np.random.default_rng(5)
    batch = (verts[None] + 0.002 * rng.standard_normal((8,) + verts.shape)).astype(np.float32)
    v = dev(batch).requires_grad_(True)
    t = dev(tris)
    normals = meshes.compute_vertex_normals(v, t)
    assert_bits(normals.detach().cpu().numpy(), oracle.vertex_normals(batch, tris))
    assert_bits(meshes.compute_vertex_normals(v.detach(), t).cpu().numpy(), normals.detach().cpu().numpy())
    lengths = normals.detach().norm(dim=-1)
    assert float((lengths - 1.0).abs().max()) < 1e-5
    normals.backward(dev(rng.standard_normal(batch.shape).astype(np.float32)))
    # normals do not change under translation of the whole mesh: the gradient sums to zero over the vertices
    total = v.grad.sum(dim=1).abs().max()
    assert float(total) <= 1e-3 * float(v.grad.abs().max()) * np.sqrt(verts.shape[0])


def test_cpu_tensors_round_trip_and_cache(meshes):
    g = load_golden("mesh_normals_cube")
    out = meshes.compute_vertex_normals(torch.from_numpy(g["vertices"]), torch.from_numpy(g["triangles"]))
    assert out.device.type == "cpu"
    assert_bits(out.numpy(), g["normals"])
    t = dev(g["triangles"])
    a = meshes.compute_vertex_normals(dev(g["vertices"]), t)
    t[0] = t[0].flip(0)                                       # in-place edit: the cached table must not be reused
    b = meshes.compute_vertex_normals(dev(g["vertices"]), t)
    tri2 = g["triangles"].copy()
    tri2[0] = tri2[0][::-1]
    want = mirror_normals(torch.from_numpy(g["vertices"]), torch.from_numpy(tri2)).numpy()
    assert np.abs(b.cpu().numpy() - want).max() < 1e-6 and not torch.equal(a, b)


def test_load_obj_without_normals_matches_reference(meshes):
    from pytorch_mesh_renderer_b200 import obj_utils
    want = load_golden("mesh_obj_plain")
    for flag in (True, False):
        v, f, n = obj_utils.load_obj(os.path.join(GOLDEN_DIR, "mesh_obj_plain.obj"), normalize=flag)
        assert_bits(v.numpy(), want["vertices_%d" % flag])
        assert_bits(f.numpy(), want["faces_%d" % flag])
        assert_bits(n.numpy(), want["normals_%d" % flag])
