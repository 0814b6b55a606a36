"""CPU: the C-ABI library builds/loads and exports exactly what include/*.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    syms = set()
    inc = os.path.join(ROOT, "include")
    for f in os.listdir(inc):
        txt = open(os.path.join(inc, f)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        syms |= set(re.findall(r"\b(rf_[a-z0-9_]+)\s*\(", txt))
    return syms


@pytest.fixture(scope="module")
def built_lib():
    from refign_b200 import build
    return build.build_library()


def test_library_exports_every_declared_symbol(built_lib):
    L = ctypes.CDLL(built_lib)
    declared = header_symbols()
    assert len(declared) >= 15
    for s in declared:
        assert hasattr(L, s), "symbol %s declared in include/ but not exported" % s


def test_ctypes_table_matches_header(built_lib):
    from refign_b200 import _lib
    assert set(_lib.exported_symbols()) == header_symbols()
    L = _lib.lib()
    assert L.rf_version() == 1
    assert isinstance(L.rf_last_error(), bytes)


def test_no_cpu_fallback():
    """The product path refuses CPU tensors instead of silently computing elsewhere."""
    import torch
    from refign_b200 import ops
    a = torch.randn(1, 4, 8, 8)
    with pytest.raises(RuntimeError):
        ops.spatial_correlation_sample(a, a, patch_size=9)
    with pytest.raises(RuntimeError):
        ops.warp(a, torch.zeros(1, 2, 8, 8))
    with pytest.raises(RuntimeError):
        ops.refine_fused(a, a)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "refign_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "librefign_oracle" not in txt and "oracle/_" not in txt, f
