"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares (and nothing the header does not), and the product refuses to run without CUDA."""
import os
import re
import subprocess

import pytest
import torch

from conftest import ROOT
from se3conv3d_b200 import _lib

HEADER = os.path.join(ROOT, "include", "se3conv3d_b200.h")


def header_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(se3_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(_lib.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "run `python -m se3conv3d_b200.build` (or __graft_entry__.build())"
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = set(re.findall(r"\sT\s+(se3_[a-z0-9_]+)", out))
    assert exported == set(header_symbols())
    L = _lib.lib()
    assert L.se3_abi_version() == 1
    for s in header_symbols():
        assert hasattr(L, s)


def test_library_is_sm100a_and_torch_free():
    out = subprocess.check_output(["cuobjdump", "--list-elf", _lib.LIB_PATH], text=True)
    assert "sm_100a" in out
    ldd = subprocess.check_output(["ldd", _lib.LIB_PATH], text=True)
    assert "torch" not in ldd and "c10" not in ldd


def test_sizes_are_computable_without_a_gpu():
    L = _lib.lib()
    assert L.se3_ball_query_workspace_bytes(1000, 500) > 0
    assert L.se3_knn_workspace_bytes(1000) > 0
    assert L.se3_csr_transpose_workspace_bytes(5000, 1000) > 0


def test_no_cpu_fallback():
    from se3conv3d_b200 import point_cloud_lib_ops as ops
    pts = torch.rand(10, 3)
    with pytest.raises(_lib.Se3Error):
        ops.knn_query(pts, torch.zeros(10, dtype=torch.int32), 4)
    with pytest.raises(_lib.Se3Error):
        ops.compute_keys(pts, torch.zeros(10, dtype=torch.int32), torch.zeros(1, 3), torch.ones(3, dtype=torch.int32),
                         torch.ones(3))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "se3conv3d_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f
                assert "libse3_oracle" not in text and "oracle._" not in text, f
