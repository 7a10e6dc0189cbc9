"""The CPU oracle (oracle/raster_oracle.c) against vectors produced by the unmodified reference.

Golden vectors come from tests/golden/make_golden.py (reference kernel
rasterize_triangles.cpp:131-273, 302-419 and reference rasterize_clip_space rasterize.py:66-152).
Everything is compared BIT-EXACT: the oracle restates the reference's rounding points and
summation orders, not just its mathematics.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, assert_bits, golden_names, grad_from_seed, load_golden

KERNEL_CASES = [n for n in golden_names() if "df_dbary" in np.load(os.path.join(GOLDEN_DIR, n + ".npz")).files
                and not n.endswith("640x480") and not n.startswith("render_")]
FULL_CASES = [n for n in golden_names() if "clip_vertices" in np.load(os.path.join(GOLDEN_DIR, n + ".npz")).files
              and not n.endswith("640x480") and not n.startswith("render_")]
DIGESTS = json.load(open(os.path.join(GOLDEN_DIR, "digests.json")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_case_inventory():
    assert len(KERNEL_CASES) >= 10 and len(FULL_CASES) >= 10


@pytest.mark.parametrize("name", KERNEL_CASES)
def test_kernel_forward_backward_bit_exact(oracle, name):
    c = load_golden(name)
    W, H = int(c["width"]), int(c["height"])
    ids, bary, z = oracle.forward(c["vertices"], c["triangles"], W, H)
    assert_bits(ids, c["ids"], "ids")
    assert_bits(bary, c["bary"], "bary")
    assert_bits(z, c["z"], "z")
    dv = oracle.backward(c["df_dbary"], c["vertices"], c["triangles"], c["ids"], c["bary"])
    assert_bits(dv, c["df_dvertices"], "df_dvertices")
    assert not dv[:, 2].any()          # z column never receives gradient (K.cpp:232-269)


@pytest.mark.parametrize("name", FULL_CASES)
def test_full_path_bit_exact(oracle, name):
    c = load_golden(name)
    r = oracle.rasterize_clip_space(c["clip_vertices"], c["attributes"], c["triangles"],
                                    int(c["width"]), int(c["height"]), c["background"],
                                    grad_out=c["grad_out"])
    for k_mine, k_ref in (("ids", "ids"), ("bary", "bary"), ("z", "z"), ("out", "out"),
                          ("d_attributes", "d_attributes"), ("d_vertices", "d_clip_vertices")):
        assert_bits(r[k_mine], c[k_ref], k_ref)


@pytest.mark.parametrize("name", ["simple_triangle", "perspective_triangle"])
def test_reference_triangle_tests_640x480(oracle, name):
    """rasterize_triangles_test.py:72-77 inputs at the test's own resolution (digest-pinned)."""
    c = load_golden(name + "_640x480")
    ids, bary, z = oracle.forward(c["vertices"], c["triangles"], 640, 480)
    d = DIGESTS[name]
    assert sha(ids) == d["ids"] and sha(bary) == d["bary"] and sha(z) == d["z"]
    assert_bits(bary[::7, ::5], c["bary_sample"], "bary sample")
    g = grad_from_seed(c["df_dbary_seed"], (480, 640, 3))
    dv = oracle.backward(g, c["vertices"], c["triangles"], ids, bary)
    assert sha(dv) == d["df_dvertices"]
    assert_bits(dv, c["df_dvertices"], "df_dvertices")


@pytest.mark.parametrize("name", ["two_cubes", "c1_cube"])
def test_reference_cube_tests_640x480(oracle, name):
    """rasterize_triangles_test.py:79-117 (A=4) and the BASELINE c1 geometry (A=9), full size."""
    c = load_golden(name + "_640x480")
    g = grad_from_seed(c["grad_out_seed"], (2, 480, 640, c["attributes"].shape[2]))
    r = oracle.rasterize_clip_space(c["clip_vertices"], c["attributes"], c["triangles"], 640, 480,
                                    c["background"], grad_out=g)
    d = DIGESTS[name]
    for k_mine, k_ref in (("ids", "ids"), ("bary", "bary"), ("z", "z"), ("out", "out"),
                          ("d_attributes", "d_attributes"), ("d_vertices", "d_clip_vertices")):
        assert sha(r[k_mine]) == d[k_ref], k_ref
    assert_bits(r["d_vertices"], c["d_clip_vertices"], "d_clip_vertices")


def _read_png(name):
    from PIL import Image
    return np.asarray(Image.open(os.path.join(GOLDEN_DIR, "reference_png", name))).astype(np.float64) / 255.0


def _near_png(image, png, max_outlier_fraction=0.001, threshold=0.01):
    """test_utils.py:105-160 comparison rule."""
    assert image.shape == png.shape
    diff = np.abs(png - np.clip(image, 0.0, 1.0))
    return np.any(diff > threshold, axis=2).mean() <= max_outlier_fraction


@pytest.mark.parametrize("name,png", [("simple_triangle", "Simple_Triangle.png"),
                                      ("perspective_triangle", "Perspective_Corrected_Triangle.png")])
def test_reference_png_triangles(oracle, name, png):
    c = load_golden(name + "_640x480")
    _, bary, _ = oracle.forward(c["vertices"], c["triangles"], 640, 480)
    image = np.concatenate([bary, np.ones((480, 640, 1), np.float32)], 2)
    assert _near_png(image, _read_png(png))


def test_reference_png_unlit_cubes(oracle):
    c = load_golden("two_cubes_640x480")
    r = oracle.rasterize_clip_space(c["clip_vertices"], c["attributes"], c["triangles"], 640, 480,
                                    c["background"])
    for i in (0, 1):
        assert _near_png(r["out"][i], _read_png("Unlit_Cube_%d.png" % i))


def test_tie_goes_to_highest_id(oracle):
    """SURVEY.md F1 / K.cpp:401: equal depth overwrites, so the last duplicate wins."""
    v = np.array([[-0.5, -0.5, 0.2, 1], [0, 0.5, 0.2, 1], [0.5, -0.5, 0.2, 1]], np.float32)
    t = np.array([[0, 1, 2]] * 3, np.int32)
    ids, bary, _ = oracle.forward(v, t, 32, 32)
    covered = bary.sum(-1) > 0.5
    assert covered.any() and (ids[covered] == 2).all() and (ids[~covered] == 0).all()


def test_no_backface_culling(oracle):
    """SURVEY.md F4 / K.cpp:79-84."""
    v = np.array([[-0.5, -0.5, 0.2, 1], [0, 0.5, 0.2, 1], [0.5, -0.5, 0.2, 1]], np.float32)
    a = oracle.forward(v, np.array([[0, 1, 2]], np.int32), 40, 30)[1]
    b = oracle.forward(v, np.array([[0, 2, 1]], np.int32), 40, 30)[1]
    assert ((a.sum(-1) > 0.5) == (b.sum(-1) > 0.5)).all() and (a.sum(-1) > 0.5).sum() > 50
