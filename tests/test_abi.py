"""The C-ABI library loads, exports every symbol include/yalps_b200.h declares, and refuses to run without a
CUDA device instead of falling back to the CPU.  No compute calls here."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from yalps_b200 import _ffi


def header_functions():
    text = open(os.path.join(ROOT, "include", "yalps_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(yalps_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_ffi.LIB_PATH)
    names = header_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/yalps_b200.h but not exported"


def test_binding_table_covers_header():
    assert sorted(_ffi.SIGNATURES) == header_functions()


def test_default_options_match_reference_defaults():
    lib = _ffi.load()
    o = _ffi.Options()
    lib.yalps_default_options(ctypes.byref(o))
    assert (o.precision, o.max_pivots, o.tolerance, o.max_iterations, o.check_cycles) == (1e-8, 8192, 0, 32768, 0)
    assert o.timeout_ms == float("inf")  # src/YALPS.ts:52-60


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import yalps_b200
    with pytest.raises(_ffi.YalpsError) as e:
        yalps_b200.Engine(0)
    assert "no CPU fallback" in str(e.value)
    with pytest.raises(_ffi.YalpsError):
        yalps_b200.solve({"variables": {}, "constraints": {}})


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under yalps_b200/ may reference it."""
    pkg = os.path.join(ROOT, "yalps_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text, f
