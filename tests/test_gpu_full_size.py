"""BASELINE.json configs c2..c5 at their FULL sizes on the GPU.

The whole batch runs through the CUDA path; parity is then established two ways:
  * sampled images of the batch are recomputed by the CPU oracle at full resolution and must match
    BIT-EXACTLY (ids, barycentrics, z, image; ORDERED-mode gradients of those images);
  * size-independent properties over the whole batch: id range, barycentric sums, clear values,
    fused image == standalone interpolation of the produced buffers, run-to-run determinism, exact
    linearity of the ORDERED backward under scaling by 2, ATOMIC gradients within the summation-order
    bound of the oracle's double-accumulated yardstick, and the full-batch ATOMIC backward (twice) against
    the full-batch ORDERED one.
"""
import numpy as np
import pytest

from conftest import assert_bits

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pmr():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pytorch_mesh_renderer_b200 as m
    return m


def _scene(name):
    from pytorch_mesh_renderer_b200 import synthetic as S
    if name == "c2":
        return S.sphere_views(159, 158, 64, 512), (0, 37)
    if name == "c3":
        return S.sphere_views(708, 707, 16, 1024), (5,)
    if name == "c4":
        return S.sphere_views(224, 223, 256, 512), (100, 255)
    if name == "c5":
        return S.occlusion_soup(32, 2048), (3,)
    raise ValueError(name)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("name", ["c2", "c3", "c4", "c5"])
def test_full_size_config(pmr, oracle, name):
    from pytorch_mesh_renderer_b200 import ops
    sc, sample = _scene(name)
    W, H = sc["width"], sc["height"]
    B, V, A = sc["attributes"].shape
    T = sc["triangles"].shape[0]
    tris, bg = dev(sc["triangles"]), dev(sc["background"])
    clip, attrs = dev(sc["clip_vertices"]), dev(sc["attributes"])

    out, (ids, bary, z) = pmr.rasterize_clip_space(clip, attrs, tris, W, H, bg, return_buffers=True)

    # ---- properties over the whole batch
    assert int(ids.min()) >= 0 and int(ids.max()) < T
    bsum = bary.sum(-1)
    covered = bsum > 0.5
    assert torch.all((bsum[covered] - 1.0).abs() < 1e-5) and torch.all(bary[~covered] == 0)
    assert torch.all(z[~covered] == 1.0) and torch.all(ids[~covered] == 0)
    assert torch.all((z >= -1.0) & (z <= 1.0))
    assert torch.all(out[~covered] == bg)
    again, (ids2, bary2, z2) = pmr.rasterize_clip_space(clip, attrs, tris, W, H, bg, return_buffers=True)
    assert torch.equal(ids, ids2) and torch.equal(bary, bary2) and torch.equal(z, z2) and torch.equal(out, again)
    del again, ids2, bary2, z2
    assert torch.equal(ops.interpolate_forward(attrs, tris, ids, bary, bg), out)      # fused == standalone
    coverage = float(covered.float().mean())
    assert coverage > 0.3, coverage

    # ---- sampled images against the oracle at full resolution, bit-exact
    for b in sample:
        g = np.random.default_rng(1000 + b).standard_normal((1, H, W, A), dtype=np.float32)
        ref = oracle.rasterize_clip_space(sc["clip_vertices"][b:b + 1], sc["attributes"][b:b + 1], sc["triangles"],
                                          W, H, sc["background"], grad_out=g, f64_yardstick=True)
        assert_bits(ids[b].cpu().numpy(), ref["ids"][0], "ids[%d]" % b)
        assert_bits(bary[b].cpu().numpy(), ref["bary"][0], "bary[%d]" % b)
        assert_bits(z[b].cpu().numpy(), ref["z"][0], "z[%d]" % b)
        assert_bits(out[b].cpu().numpy(), ref["out"][0], "image[%d]" % b)
        gd = dev(g)
        dv, da = ops.rasterize_interpolate_backward(gd, clip[b:b + 1], attrs[b:b + 1], tris, ids[b:b + 1],
                                                    bary[b:b + 1], "ordered")
        assert_bits(dv.cpu().numpy(), ref["d_vertices"], "d_vertices[%d] (ordered)" % b)
        assert_bits(da.cpu().numpy(), ref["d_attributes"], "d_attributes[%d] (ordered)" % b)
        dv2, da2 = ops.rasterize_interpolate_backward(2.0 * gd, clip[b:b + 1], attrs[b:b + 1], tris, ids[b:b + 1],
                                                      bary[b:b + 1], "ordered")
        assert torch.equal(dv2, 2.0 * dv) and torch.equal(da2, 2.0 * da)            # exact linearity
        dva, daa = ops.rasterize_interpolate_backward(gd, clip[b:b + 1], attrs[b:b + 1], tris, ids[b:b + 1],
                                                      bary[b:b + 1], "atomic")
        for mine, key in ((dva, "d_vertices"), (daa, "d_attributes")):
            err = np.abs(mine.cpu().numpy().astype(np.float64) - ref[key + "_f64"])
            assert (err <= oracle.atomic_mode_bound(ref[key], ref[key + "_f64"])).all(), key

    # ---- the full-batch backward in throughput mode runs and is finite
    g_all = torch.randn((B, H, W, A), device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    dv, da = ops.rasterize_interpolate_backward(g_all, clip, attrs, tris, ids, bary, "atomic")
    assert torch.isfinite(dv).all() and torch.isfinite(da).all() and not dv[..., 2].any()
    # ... and gives the ORDERED (bit-exact) mode's sums up to the order of the additions, over the WHOLE batch and
    # twice: under full load a backward kernel that refilled its tensor-copy boxes too early once produced a
    # fraction of 1e-5 of the entries wrong by up to the size of the entry (never on a single image).
    dvo, dao = ops.rasterize_interpolate_backward(g_all, clip, attrs, tris, ids, bary, "ordered")
    for rep in range(2):
        if rep:
            dv, da = ops.rasterize_interpolate_backward(g_all, clip, attrs, tris, ids, bary, "atomic")
        for mine, ref, what in ((dv, dvo, "d_vertices"), (da, dao, "d_attributes")):
            off = (mine - ref).abs() > 1e-4 * float(ref.abs().max())
            assert not bool(off.any()), "%s: %d entries of the atomic mode differ from the ordered mode (run %d)" % (
                what, int(off.sum()), rep)
