"""Library switches that are read once, in pmr_create: the cases run once more in a process of their own.

  * PMR_NO_TMA=1: the paths that do without tensor maps (per-lane loads in the backward kernel, per-row bulk copies in
    the resolve epilogue) must give the same bits;
  * PMR_STRIP_BLOCKS=4: on test-sized images a warp's strip is one block long, so the strip loops of the backward and
    resolve kernels (the next block's boxes in flight, stores draining behind the next block, strips that end inside the
    image) are forced here -- the full-size tests are the only other place they run."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_golden_cases_without_tensor_maps():
    env = dict(os.environ, PMR_NO_TMA="1")
    run = subprocess.run(
        [sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider",
         os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-k", "golden"],
        cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert run.returncode == 0, run.stdout[-3000:] + run.stderr[-2000:]
    assert " passed" in run.stdout


@pytest.mark.gpu
def test_small_cases_with_strips_of_four_blocks():
    env = dict(os.environ, PMR_STRIP_BLOCKS="4")
    run = subprocess.run(
        [sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider",
         os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-k", "golden or partial_boxes or random_scenes"],
        cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    assert run.returncode == 0, run.stdout[-3000:] + run.stderr[-2000:]
    assert " passed" in run.stdout
