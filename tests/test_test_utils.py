"""pytorch_mesh_renderer_b200/test_utils.py and debug_utils.py (counterparts of the reference's
src/mesh_renderer/test_utils.py and src/common/debug_utils.py) on the CPU: the helpers are plain torch / numpy."""
import unittest

import numpy as np
import pytest
import torch

from pytorch_mesh_renderer_b200 import debug_utils, test_utils


def test_jacobian_helpers_agree_on_a_smooth_function():
    w = torch.randn(4, 3, generator=torch.Generator().manual_seed(0), dtype=torch.float64)
    fn = lambda x: torch.sin(x) @ w
    x = torch.randn(5, 4, generator=torch.Generator().manual_seed(1), dtype=torch.float64, requires_grad=True)
    analytical = test_utils.get_analytical_jacobian(x, fn(x))
    numerical = test_utils.get_numerical_jacobian(fn, x, eps=1e-4)
    assert analytical.shape == numerical.shape == (20, 15)
    # zero entries of the true Jacobian make the relative error 0/0 = nan, which counts as "not an outlier"
    ok, message = test_utils.check_jacobians_are_nearly_equal(analytical.numpy(), numerical.numpy(), 1e-3, 0.0)
    assert ok, message
    broken = analytical * 1.5          # (the rule divides by the signed numerical value: only positive entries count)
    ok, message = test_utils.check_jacobians_are_nearly_equal(broken.numpy(), numerical.numpy(), 1e-3, 0.01, True)
    assert not ok and "Numerical Jacobian" in message


def test_image_comparison(tmp_path):
    from PIL import Image
    rng = np.random.default_rng(0)
    image = rng.integers(0, 256, (24, 32, 4), dtype=np.uint8)
    path = str(tmp_path / "baseline.png")
    Image.fromarray(image).save(path)
    render = torch.from_numpy(image.astype(np.float32) / 255.0)
    case = unittest.TestCase()
    test_utils.expect_image_file_and_render_are_near(case, path, render * 1.5 - 0.0, max_outlier_fraction=1.0)
    test_utils.expect_image_file_and_render_are_near(case, path, render, outputs_dir=str(tmp_path))
    off = render.clone()
    off[:5] = 1.0 - off[:5]                                     # 5 of 24 rows wrong
    match, fraction, _ = test_utils.images_are_near(image, off.numpy())
    assert not match and 0.15 < fraction <= 5 / 24
    with pytest.raises(AssertionError, match="does not match"):
        test_utils.expect_image_file_and_render_are_near(case, path, off, outputs_dir=str(tmp_path))
    assert (tmp_path / "baseline_result.png").exists() and (tmp_path / "baseline_diff.png").exists()
    with pytest.raises(AssertionError, match="do not match"):
        test_utils.expect_image_file_and_render_are_near(case, path, render[:, :16])


def test_debug_utils(capsys):
    debug_utils.check_isnan_isinf(torch.zeros(3), "fine")
    with pytest.raises(ValueError, match="bad normals"):
        debug_utils.check_isnan_isinf(torch.tensor([0.0, float("nan")]), "bad normals")
    with pytest.raises(ValueError):
        debug_utils.check_isnan_isinf(torch.tensor([float("inf")]))
    debug_utils.debug_tensor(torch.arange(3), "three")
    assert "[debug tensor] three" in capsys.readouterr().out
