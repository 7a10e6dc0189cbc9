"""The paths that do without tensor maps (PMR_NO_TMA=1: per-lane loads in the backward kernel, per-row bulk
copies in the resolve epilogue) must give the same bits: the golden kernel and full-path cases once more, in a
process whose library context is created with the switch set (it is read in pmr_create)."""
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
