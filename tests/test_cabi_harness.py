"""The torch-free C++ consumer of the C-ABI (tests/cabi_device_harness.cu): real device work through include/yolopost_b200.h with
plain cudaMalloc'd buffers - fused path == two-call path bit for bit in every scan form, TorchNMS against a scalar restatement,
the result-row kernels and the one-sided ring looped back onto one GPU.  No Python / torch on the compute path at all."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _env(**extra):
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = "/usr/local/cuda/lib64" + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    env.update(extra)
    return env


def _harness():
    from ultralytics_pro_b200 import build

    return build.build_harness()


def test_harness_builds_and_lists_its_cases():
    out = subprocess.run([_harness(), "--list"], capture_output=True, text=True, env=_env(), timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ABI version" in out.stdout


@pytest.mark.gpu
def test_harness_full_sizes(cuda_device):
    out = subprocess.run([_harness()], capture_output=True, text=True, env=_env(), timeout=300)
    assert out.returncode == 0, out.stdout[-4000:] + out.stderr[-2000:]
    assert "FAIL" not in out.stdout


@pytest.mark.gpu
def test_harness_split_decode_forms(cuda_device):
    """YPB_FUSE_DECODE=0 forces the throughput form (separate survivor-decode kernel, persistent TMA scan where asked for) on the
    small geometries too."""
    out = subprocess.run([_harness(), "--small"], capture_output=True, text=True, env=_env(YPB_FUSE_DECODE="0"), timeout=300)
    assert out.returncode == 0, out.stdout[-4000:] + out.stderr[-2000:]
    assert "FAIL" not in out.stdout
