"""A pure-C program drives the drop-in boundary (tests/cabi/cabi_smoke.c against include/raycore_cuda.h + libraycore_cuda.so):
what a cgo / ccall / JNI binding would see, with no Python or torch in the process."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "raycore.jl_b200")


def _build(tmp_path, name="cabi_smoke"):
    exe = str(tmp_path / name)
    subprocess.check_call(["gcc", "-std=gnu99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cabi", name + ".c"),
                           "-o", exe, "-L", LIBDIR, "-lraycore_cuda", "-lm", f"-Wl,-rpath,{LIBDIR}"])
    return exe


def test_c_program_links_and_fails_loudly_without_a_gpu(tmp_path):
    import torch

    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu-marked test")
    p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert p.returncode == 2 and "rc_create" in p.stderr  # no CPU fallback: context creation reports the CUDA error


@pytest.mark.gpu
def test_c_program_runs_on_the_gpu(tmp_path):
    p = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "cabi smoke ok" in p.stdout, p.stderr


def test_multi_gpu_c_program_links(tmp_path):
    import torch

    exe = _build(tmp_path, "cabi_multi")
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu-marked test")
    p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert p.returncode == 2 and "rc_multi_create" in p.stderr and "no CPU fallback" in p.stderr


@pytest.mark.gpu
def test_multi_gpu_c_program_runs(tmp_path):
    """one process, every visible GPU, pure C: the sharded trace is byte-identical to the single-device one (tests/cabi/cabi_multi.c)"""
    p = subprocess.run([_build(tmp_path, "cabi_multi")], capture_output=True, text=True, timeout=600)
    print(p.stdout)
    assert p.returncode == 0 and "cabi multi ok" in p.stdout, p.stderr + p.stdout
