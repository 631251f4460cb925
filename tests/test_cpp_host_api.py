"""The C++ host API (include/loam/*.h, same signatures as the reference's loam/*.h) over the C-ABI library.

CPU run: the test program compiles, links against libloamgpu.so, exercises the value types / error paths and confirms
that a call without a CUDA device fails loudly (there is no CPU fallback).  GPU run: the reference's unit-test scenes.
"""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CPP = os.path.join(HERE, "cpp")


def build():
    from loam_b200 import build as b
    b.build()
    subprocess.check_call(["make", "-C", CPP, "build"], stdout=subprocess.DEVNULL)
    return os.path.join(CPP, "test_host_api")


def test_cpp_host_api_builds_and_fails_loudly_without_gpu():
    exe = build()
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present: the no-device path cannot be exercised here")
    r = subprocess.run([exe, "--no-gpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failed" in r.stdout


@pytest.mark.gpu
def test_cpp_host_api_reference_scenarios_on_gpu():
    exe = build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failed" in r.stdout
