"""The C-ABI library loads and exports every symbol include/raycore_cuda.h declares (no compute calls: no GPU here)."""
import os
import re

import pytest

import raycore_b200 as rc
from raycore_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "raycore_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rc_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    for name in _declared():
        assert hasattr(L, name), name
    assert L.rc_abi_version() == 1


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(rc.RaycoreError) as e:
        rc.TLAS()
    assert e.value.code == _lib.RC_ERR_CUDA and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "raycore.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".jl")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle.h" not in txt and "liboracle" not in txt, f


def test_header_is_plain_c_and_cxx(tmp_path):
    """include/raycore_cuda.h is the drop-in boundary: it must compile as C99 and as C++11 on its own (plain pointers and sizes)."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "hdr.c"
    src.write_text('#include "raycore_cuda.h"\nint main(void) { rc_ray r; rc_hit h; (void)r; (void)h; return RC_ABI_VERSION ? 0 : 1; }\n')
    inc = os.path.join(root, "include")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", inc, str(src)])
    subprocess.check_call(["g++", "-std=c++11", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, "-x", "c++", str(src)])
