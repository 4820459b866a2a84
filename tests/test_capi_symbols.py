"""CPU: libmplb.so loads and exports every symbol include/mplb.h declares; without a GPU the
compute entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import mplambda_b200 as M
from mplambda_b200 import _capi
from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mplb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mplb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    lib = ctypes.CDLL(M.lib_path())
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "libmplb.so does not export %s" % n


def test_python_binding_covers_header():
    assert set(declared_symbols()) <= set(_capi._SIGNATURES)


def test_no_cpu_fallback_without_gpu():
    if M.lib().mplb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    tri = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0.0]]])
    with pytest.raises(M.MplbError) as ei:
        M.SE3RigidBodyScenario(tri, tri, [0, 0, 0, 1, 0, 0, 0], [0, 0, 0], [1, 1, 1])
    assert ei.value.code == _capi.ERR_CUDA
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "mplambda_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|#include\s+[\"<].*oracle", src, flags=re.M), f
