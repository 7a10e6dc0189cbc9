"""CPU-side checks of the mesh utilities around the renderer: the vertex-normal
oracle, the shape generators and OBJ IO against golden vectors produced by the unmodified reference
(tests/golden/make_golden_mesh.py), plus the host-side image IO of the examples."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, assert_bits, golden_names, load_golden

NORMAL_CASES = golden_names("mesh_normals_")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_golden_cases_exist():
    assert len(NORMAL_CASES) >= 4


@pytest.mark.parametrize("name", NORMAL_CASES)
def test_oracle_vertex_normals_match_reference_bits(oracle, name):
    """oracle/raster_oracle.c pmr_oracle_vertex_normals == the reference's compute_vertex_normals, bit for bit."""
    g = load_golden(name)
    assert_bits(oracle.vertex_normals(g["vertices"], g["triangles"]), g["normals"], name)


def test_shape_generators_match_reference_digests():
    from pytorch_mesh_renderer_b200 import shapes
    digests = json.load(open(os.path.join(GOLDEN_DIR, "mesh_shapes.json")))
    for key, want in digests.items():
        kind, size, *rest = key.split("_")
        v, t, n = shapes.sphere(float(size), int(rest[0])) if kind == "sphere" else shapes.cube(float(size))
        assert (v.dtype, t.dtype, n.dtype) == (torch.float32, torch.int32, torch.float32)
        assert [v.shape[0], t.shape[0]] == want["counts"], key
        assert sha(v.numpy()) == want["vertices"], key
        assert sha(t.numpy()) == want["triangles"], key
        assert sha(n.numpy()) == want["normals"], key


@pytest.mark.parametrize("stem", ["mesh_obj_normals", "mesh_obj_quads"])
def test_load_obj_matches_reference(stem):
    """Files that carry normals need no device: parse, per-vertex averaging and the unit-cube normalisation."""
    import warnings
    from pytorch_mesh_renderer_b200 import obj_utils
    want = load_golden(stem)
    for flag in (True, False):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            v, f, n = obj_utils.load_obj(os.path.join(GOLDEN_DIR, stem + ".obj"), normalize=flag)
        assert_bits(v.numpy(), want["vertices_%d" % flag], stem)
        assert_bits(f.numpy(), want["faces_%d" % flag], stem)
        assert_bits(n.numpy(), want["normals_%d" % flag], stem)


def test_load_obj_warns_about_polygons():
    from pytorch_mesh_renderer_b200 import obj_utils
    with pytest.warns(UserWarning, match="more than 3 vertices"):
        obj_utils.load_obj(os.path.join(GOLDEN_DIR, "mesh_obj_quads.obj"))


def test_save_obj_writes_the_reference_bytes(tmp_path):
    """The golden .obj files were written by the reference's save_obj from exactly these tensors."""
    from pytorch_mesh_renderer_b200 import obj_utils, shapes
    v, t, n = shapes.sphere(0.8, 6)
    v = v + torch.tensor([0.3, -0.2, 0.5])
    obj_utils.save_obj(str(tmp_path / "a.obj"), v, t, n)
    assert (tmp_path / "a.obj").read_text() == open(os.path.join(GOLDEN_DIR, "mesh_obj_normals.obj")).read()
    obj_utils.save_obj(str(tmp_path / "b.obj"), v.requires_grad_(True), t)
    assert (tmp_path / "b.obj").read_text() == open(os.path.join(GOLDEN_DIR, "mesh_obj_plain.obj")).read()


def test_save_obj_argument_errors():
    from pytorch_mesh_renderer_b200 import obj_utils
    v, t = torch.zeros(4, 3), torch.zeros(2, 3, dtype=torch.int32)
    with pytest.raises(ValueError, match="vertices must have shape"):
        obj_utils.save_obj(os.devnull, v[None], t)
    with pytest.raises(ValueError, match="faces must have shape"):
        obj_utils.save_obj(os.devnull, v, t[:, :2])
    with pytest.raises(ValueError, match="normals must have shape"):
        obj_utils.save_obj(os.devnull, v, t, v[:, :2])


def test_vertex_normals_need_a_device():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pytorch_mesh_renderer_b200 import meshes, obj_utils
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        meshes.compute_vertex_normals(torch.zeros(1, 3, 3), torch.zeros(1, 3, dtype=torch.int32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        obj_utils.load_obj(os.path.join(GOLDEN_DIR, "mesh_obj_plain.obj"))
    with pytest.raises(ValueError, match="vertices must have shape"):
        meshes.compute_vertex_normals(torch.zeros(3, 3), torch.zeros(1, 3, dtype=torch.int32))


def test_example_image_io_round_trip(tmp_path):
    from pytorch_mesh_renderer_b200.examples import image_io
    rng = np.random.default_rng(0)
    render = rng.random((12, 16, 4)).astype(np.float32)
    image_io.imsave(str(tmp_path / "a.png"), image_io.to_uint8(render))
    back = image_io.imread(str(tmp_path / "a.png"))
    assert back.shape == (12, 16, 4) and np.array_equal(back, (render * 255.0).astype(np.uint8))
    frame = image_io.frame_on_black(render)
    assert frame.shape == (12, 16, 4) and (frame[:, :, 3] == 255).all()
    writer = image_io.FrameWriter(str(tmp_path / "anim.gif"), fps=20)
    for _ in range(3):
        writer.append_data(frame)
    writer.close()
    assert (tmp_path / "anim.gif").stat().st_size > 0
    stills = image_io.FrameWriter(str(tmp_path / "still.png"))
    stills.append_data(frame)
    stills.close()
    assert (tmp_path / "still_0000.png").exists()
    image_io.FrameWriter(None).append_data(frame)


def test_examples_reach_the_device_boundary():
    """Without a GPU every example runs its host-side setup and then fails loudly at the first kernel call
    (no CPU path) -- which also proves that the modules import and bind the right callables."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pytorch_mesh_renderer_b200 import shapes
    from pytorch_mesh_renderer_b200.examples import example1, example5, example6
    v, t, n = shapes.sphere(1.0, 6)
    cube = shapes.cube(2.0)
    cpu = torch.device("cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        example1.render_obj(v, t, n, 32, 24, device=cpu)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        example5.render_cube_with_rotation(torch.zeros(1, 3), cube, cpu)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        example6.render_with_rotation(torch.zeros(1, 3), (v[None], t, n[None]), cpu)
