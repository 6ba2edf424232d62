"""The drop-in boundary on a CPU box: libslamgpu.so loads, exports exactly the symbols include/slamgpu.h
declares, and refuses to compute without a GPU (no CPU fallback).  No compute calls here."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "slamgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(slamgpu_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(sg):
    decl = declared_symbols()
    assert len(decl) >= 40
    out = subprocess.check_output(["nm", "-D", "--defined-only", sg.library_path()], text=True)
    exported = sorted(set(re.findall(r" T (slamgpu_[a-z0-9_]+)", out)))
    assert exported == decl, (set(decl) ^ set(exported))
    from slam_constructor_b200 import capi
    assert sorted(capi.SYMBOLS) == decl  # the ctypes layer binds all of them
    L = sg.lib()
    for name in decl:
        assert getattr(L, name) is not None


def test_abi_version_and_strides(sg):
    L = sg.lib()
    assert L.slamgpu_abi_version() == 1
    assert [L.slamgpu_model_stride(m) for m in range(7)] == [3, 2, 2, 6, 6, 5, 6]
    assert L.slamgpu_model_stride(99) == 0


def test_no_cpu_fallback(sg):
    L = sg.lib()
    if L.slamgpu_device_count() > 0:
        pytest.skip("a B200 is visible: the refusal path is not reachable")
    with pytest.raises(sg.SlamGpuError) as e:
        sg.Context(0)
    assert e.value.code == -6 and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's baseline legs may touch oracle/"""
    pkg = os.path.join(ROOT, "slam_constructor_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.replace("the CPU oracle", ""), os.path.join(dirpath, f)
                assert "/root/reference" not in txt, os.path.join(dirpath, f)
