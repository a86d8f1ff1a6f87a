"""CPU-only checks of the boundary: the library builds, loads, and exports exactly the symbols
include/comet_b200.h declares; compute entry points fail loudly without a GPU."""
import os
import re

import numpy as np
import pytest

from comet_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "comet_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(cm_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_are_exported_and_bound():
    L = capi.lib()
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), f"{s} declared in comet_b200.h but not exported"
        assert s in capi.SIGNATURES, f"{s} has no ctypes signature"
    for s in capi.SIGNATURES:
        assert s in syms, f"{s} bound in capi.py but not declared in the header"


def test_version_and_rounding_mode():
    L = capi.lib()
    assert b"sm_100a" in L.cm_version()
    assert L.cm_get_rounding() == capi.ROUND_SEPARATE
    assert L.cm_set_rounding(7) == capi.ERR_INVALID_ARG
    assert L.cm_set_rounding(capi.ROUND_FMA) == capi.OK and L.cm_get_rounding() == capi.ROUND_FMA
    L.cm_set_rounding(capi.ROUND_SEPARATE)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.CometError) as e:
        capi.FlatIndex(8, capi.L2)
    assert e.value.code == capi.ERR_CUDA and "no CPU fallback" in e.value.msg
    with pytest.raises(capi.CometError):
        capi.distance_pairs(capi.L2, np.ones(4, np.float32), np.ones(4, np.float32))


def test_argument_validation_needs_no_gpu():
    L = capi.lib()
    import ctypes as C
    h = C.c_void_p()
    assert L.cm_flat_create(0, capi.L2, C.byref(h)) == capi.ERR_INVALID_ARG
    assert b"dimension must be positive" in L.cm_last_error()
    assert L.cm_flat_create(8, 9, C.byref(h)) == capi.ERR_INVALID_ARG


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "comet_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dp, f), errors="replace").read()
                assert "oracle" not in src.lower() or f == "__init__.py" and "oracle" not in src.lower(), \
                    f"{f} mentions the oracle"
