"""CPU tests: the C-ABI library loads, exports every symbol include/qtb.h declares, and refuses to compute without a
GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "qtb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qtb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(engine):
    lib = engine.load_library()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/qtb.h but not exported by libqtb.so"
    assert sorted(engine.EXPORTED_SYMBOLS) == names, "python binding and header disagree"


def test_no_cpu_fallback(engine):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(engine.NoDeviceError):
        engine.Context(0)
    lib = engine.load_library()
    assert b"no CPU fallback" in lib.qtb_last_error() or b"CUDA" in lib.qtb_last_error()


def test_product_never_imports_oracle():
    """the oracle is test infrastructure: nothing under quantit_b200/ may import, link or execute it"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "quantit_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "qtb_oracle" not in txt and "ref_harness" not in txt and "oracle/" not in txt, f


def test_cpp_binding_fails_loudly_without_a_gpu():
    """the reference-side C++ binding (oracle/_ref/adaptor_check) goes through the same C ABI: without a CUDA device the
    engine refuses to run (QTB_ERR_NO_DEVICE -> a C++ exception on the reference side), it never falls back to the CPU"""
    import subprocess

    import pytest
    import torch

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    binary = os.path.join(root, "oracle", "_ref", "adaptor_check")
    if not os.path.exists(binary):
        pytest.skip("compiled reference-side adaptor (oracle/_ref) not present")
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by tests/test_gpu_adaptor.py")
    g = os.path.join(root, "tests", "golden")
    out = subprocess.run([binary, "--lib", os.path.join(root, "quantit_b200", "libqtb.so"), "tdot",
                          os.path.join(g, "tdot1_A.qtbt"), os.path.join(g, "tdot1_B.qtbt"), "3", "0"],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 3 and "no usable CUDA device" in out.stderr and "no CPU fallback" in out.stderr
