"""CPU: the C-ABI library loads and exports every symbol include/b200lic.h declares; the product path refuses to run
without the GPU (no fallback).  No compute call is made here."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "b200lic.h")).read()
    return sorted(set(re.findall(r"B200LIC_API[^;(]*?\b(b200lic_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from rdo_ptq_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    L = ctypes.CDLL(_lib.LIB_PATH)
    declared = header_symbols()
    assert len(declared) >= 35
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(_lib.exported_names()) == declared            # the ctypes table covers the whole header
    L.b200lic_version.restype = ctypes.c_int
    assert L.b200lic_version() >= 100


def test_conv_desc_mirror_matches_header():
    from rdo_ptq_b200._lib import ConvDesc
    src = open(os.path.join(ROOT, "include", "b200lic.h")).read()
    body = src[src.index("typedef struct {"):src.index("} b200lic_conv_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = [f.strip() for decl in re.findall(r"(?:int|float)\s+([^;]+);", body) for f in decl.split(",")]
    assert fields == [n for n, _ in ConvDesc._fields_]


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback():
    from rdo_ptq_b200 import ops, _lib
    assert _lib.lib().b200lic_device_check() == -2              # B200LIC_ERR_ARCH: no sm_100 device
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.conv2d(torch.randn(1, 3, 8, 8), torch.randn(4, 3, 3, 3), None, 1, 1)
    with pytest.raises(RuntimeError):
        ops.act_quant(torch.randn(1, 3, 8, 8))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "rdo_ptq_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(d, f)


def test_calib_sched_mirror_matches_header():
    """The device-resident schedule struct: ctypes mirror == header field order, 16 bytes."""
    import ctypes
    import re
    from rdo_ptq_b200 import _lib
    src = open(os.path.join(ROOT, "include", "b200lic.h")).read()
    body = re.search(r"typedef struct \{([^}]*)\} b200lic_calib_sched;", src).group(1)
    fields = re.findall(r"\b(?:int|float)\s+(\w+)\s*;", body)
    assert fields == [f[0] for f in _lib.CalibSched._fields_]
    assert ctypes.sizeof(_lib.CalibSched) == 16
