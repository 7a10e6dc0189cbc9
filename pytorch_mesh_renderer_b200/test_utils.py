"""Counterpart of the reference's src/mesh_renderer/test_utils.py: the helpers its tests are written with, so that
tests of user code port over unchanged -- Jacobians by one-hot backward passes and by central differences, the
"few relative outliers" comparison, and the soft comparison of a render with an image file.  PIL instead of skimage;
tensors may live on any device (results come back on the CPU)."""
import os
from itertools import product

import numpy as np
import torch
from PIL import Image


def check_jacobians_are_nearly_equal(theoretical, numerical, outlier_relative_error_threshold, max_outlier_fraction,
                                     include_jacobians_in_error_message=False):
    """-> (success, message): success iff at most `max_outlier_fraction` of the entries of `theoretical` are further
    than `outlier_relative_error_threshold` (relative to `numerical`) from the finite-difference value
    (test_utils.py:12-51).  Mind that the reference's own tests pass the TUPLE to assertTrue, which never fails."""
    theoretical, numerical = np.asarray(theoretical), np.asarray(numerical)
    with np.errstate(divide="ignore", invalid="ignore"):
        outliers = np.abs(numerical - theoretical) / numerical > outlier_relative_error_threshold
    outlier_fraction = np.count_nonzero(outliers) / np.prod(numerical.shape[:2])
    message = (" %f of theoretical gradients are relative outliers, but the maximum allowable fraction is %f "
               % (outlier_fraction, max_outlier_fraction))
    if include_jacobians_in_error_message:
        message += "\nNumerical Jacobian:\n%r\nTheoretical Jacobian:\n%r" % (numerical.T, theoretical.T)
    return bool(outlier_fraction <= max_outlier_fraction), message


def get_analytical_jacobian(input, output):
    """[input.numel(), output.numel()]: column i is the gradient of output element i (test_utils.py:54-77)."""
    jacobian = torch.zeros(input.numel(), output.numel())
    grad_output = torch.zeros_like(output)
    flat = grad_output.view(-1)
    for i in range(flat.numel()):
        flat.zero_()
        flat[i] = 1
        d_x = torch.autograd.grad(output, [input], grad_output, retain_graph=True, allow_unused=True)[0]
        if d_x is not None:
            jacobian[:, i] = d_x.detach().contiguous().view(-1).cpu()
    return jacobian


def get_numerical_jacobian(fn, input, eps=1e-3):
    """Central differences, one input element at a time (test_utils.py:80-102)."""
    jacobian = torch.zeros(input.numel(), fn(input).numel())
    x = input.data
    for d_idx, x_idx in enumerate(product(*[range(m) for m in x.size()])):
        orig = x[x_idx].item()
        x[x_idx] = orig - eps
        outa = fn(input).clone()
        x[x_idx] = orig + eps
        outb = fn(input).clone()
        x[x_idx] = orig
        jacobian[d_idx] = ((outb - outa) / (2 * eps)).detach().reshape(-1).cpu()
    return jacobian


def images_are_near(baseline_image, result_image, max_outlier_fraction=0.001, pixel_error_threshold=0.01):
    """-> (match, outlier_fraction, diff): `baseline_image` uint8 [H,W,C], `result_image` float [H,W,C] clipped to
    [0,1]; a pixel is an outlier if any channel differs by more than the threshold (test_utils.py:130-138)."""
    result = np.clip(np.asarray(result_image, dtype=np.float64), 0.0, 1.0)
    diff = np.abs(np.asarray(baseline_image).astype(np.float64) / 255.0 - result)
    outlier_fraction = np.count_nonzero(np.any(diff > pixel_error_threshold, axis=2)) / np.prod(diff.shape[:2])
    return bool(outlier_fraction <= max_outlier_fraction), float(outlier_fraction), diff


def expect_image_file_and_render_are_near(test_instance, baseline_path, result_image, max_outlier_fraction=0.001,
                                          pixel_error_threshold=0.01, outputs_dir="/tmp"):
    """unittest-style soft comparison of a render (tensor, any device) with an image on disk (test_utils.py:105-160);
    on a mismatch the result and the difference image are written to `outputs_dir`."""
    baseline_image = np.asarray(Image.open(baseline_path))
    result = result_image.detach().cpu().numpy() if isinstance(result_image, torch.Tensor) else np.asarray(result_image)
    test_instance.assertEqual(tuple(baseline_image.shape), tuple(result.shape),
                              "Images shapes {} and {} do not match.".format(baseline_image.shape, result.shape))
    match, outlier_fraction, diff = images_are_near(baseline_image, result, max_outlier_fraction, pixel_error_threshold)
    prefix = os.path.splitext(os.path.basename(baseline_path))[0]
    result_path = os.path.join(outputs_dir, prefix + "_result.png")
    diff_path = os.path.join(outputs_dir, prefix + "_diff.png")
    if not match:
        Image.fromarray((np.clip(result, 0.0, 1.0) * 255.0).astype(np.uint8)).save(result_path)
        if diff.shape[2] == 4:
            diff[:, :, 3] = 1.0
        Image.fromarray((diff * 255.0).astype(np.uint8)).save(diff_path)
    test_instance.assertTrue(match, msg="{} does not match. ({} of pixels are outliers, {} is allowed.). Result image "
                             "written to {}, Diff written to {}".format(baseline_path, outlier_fraction,
                                                                        max_outlier_fraction, result_path, diff_path))
