"""Live pinning: oracle/raster_oracle.c against the reference's own compiled kernel
(oracle/_ref/rasterize_triangles_cpp.so, built from rasterize_triangles.cpp by
oracle/build_ref.py) and, where /root/reference is mounted, against the reference's
rasterize_clip_space -- on fresh random inputs, bit-exact."""
import numpy as np
import pytest

from conftest import assert_bits

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ref_kernel():
    from oracle import reference_harness as rh
    k = rh.kernel()
    if k is None:
        pytest.skip("oracle/_ref not built (python oracle/build_ref.py)")
    return k


def _soup(rng, T, spread, mixed_w):
    c = rng.uniform(-0.9, 0.9, (T, 1, 2))
    xy = c + spread * rng.standard_normal((T, 3, 2))
    z = rng.uniform(-0.9, 0.9, (T, 3, 1))
    w = rng.uniform(0.5, 2.0, (T, 3, 1))
    w = np.where(rng.random((T, 3, 1)) < mixed_w, -w, w)
    v = np.concatenate([xy, z, np.ones_like(z)], 2) * w
    return v.reshape(3 * T, 4).astype(np.float32), np.arange(3 * T, dtype=np.int32).reshape(T, 3)


@pytest.mark.parametrize("seed,T,W,H,spread,mixed", [
    (1, 500, 128, 96, 0.05, 0.0), (2, 2000, 333, 217, 0.08, 0.03), (3, 20, 640, 480, 0.6, 0.0),
    (4, 64, 31, 257, 0.3, 0.1), (5, 1, 8, 8, 0.5, 0.0)])
def test_oracle_equals_reference_kernel(oracle, ref_kernel, seed, T, W, H, spread, mixed):
    rng = np.random.default_rng(seed)
    v, t = _soup(rng, T, spread, mixed)
    ids, bary, z = oracle.forward(v, t, W, H)
    rid, rb, rz = ref_kernel.forward(torch.from_numpy(v), torch.from_numpy(t), W, H)
    assert_bits(ids, rid.numpy(), "ids")
    assert_bits(bary, rb.detach().numpy(), "bary")
    assert_bits(z, rz.numpy(), "z")
    g = rng.standard_normal((H, W, 3)).astype(np.float32)
    rdv, = ref_kernel.backward(torch.from_numpy(g), torch.from_numpy(v), torch.from_numpy(t), rid, rb.detach())
    assert_bits(oracle.backward(g, v, t, ids, bary), rdv.numpy(), "df_dvertices")


@pytest.mark.parametrize("A", [4, 9])
def test_oracle_equals_reference_rasterize_clip_space(oracle, A):
    from oracle import reference_harness as rh
    if not rh.available():
        pytest.skip("/root/reference not mounted")
    torch.set_num_threads(1)           # index_put_ order
    R = rh.rasterize_module()
    rng = np.random.default_rng(100 + A)
    v, t = _soup(rng, 150, 0.1, 0.0)
    clip = np.stack([v, v]); clip[1, :, 1] += np.float32(0.05) * clip[1, :, 3]
    attrs = rng.uniform(-1, 1, (2, v.shape[0], A)).astype(np.float32)
    bg = rng.uniform(-1, 1, A).astype(np.float32)
    g = rng.standard_normal((2, 45, 60, A)).astype(np.float32)
    tv = torch.tensor(clip, requires_grad=True); ta = torch.tensor(attrs, requires_grad=True)
    out = R.rasterize_clip_space(tv, ta, torch.from_numpy(t), 60, 45, torch.from_numpy(bg))
    out.backward(torch.from_numpy(g))
    r = oracle.rasterize_clip_space(clip, attrs, t, 60, 45, bg, grad_out=g)
    assert_bits(r["out"], out.detach().numpy(), "out")
    assert_bits(r["d_attributes"], ta.grad.numpy(), "d_attributes")
    assert_bits(r["d_vertices"], tv.grad.numpy(), "d_clip_vertices")
