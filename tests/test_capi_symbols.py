"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/bq_b200.h declares (no compute calls — those need a GPU)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_functions():
    with open(os.path.join(ROOT, "include", "bq_b200.h")) as fh:
        src = fh.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bqb_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    names = declared_functions()
    for must in ("bqb_batch_create", "bqb_batch_setup", "bqb_batch_info", "bqb_score_device", "bqb_score_host",
                 "bqb_expected_var_host", "bqb_expected_var_device", "bqb_mean_neg_device", "bqb_argmin_device",
                 "bqb_batch_destroy", "bqb_last_error"):
        assert must in names


def test_library_builds_and_exports_every_declared_symbol():
    from bayesian_quadrature_b200 import build, _lib
    build.build()                                    # nvcc cross-compiles sm_100a without a GPU
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), "libbq_b200.so does not export %s" % name
    assert sorted(_lib.SYMBOLS) == declared_functions()     # the Python binding covers the header exactly
    assert lib.bqb_version() >= 100


def test_capacity_and_argument_errors_without_gpu():
    from bayesian_quadrature_b200 import _lib
    L = _lib.load()
    assert L.bqb_ns_capacity(1) == 16 and L.bqb_ns_capacity(16) == 16
    assert L.bqb_ns_capacity(17) == 64 and L.bqb_ns_capacity(64) == 64
    assert L.bqb_ns_capacity(65) == 128 and L.bqb_ns_capacity(129) == 160 and L.bqb_ns_capacity(161) == 256
    assert L.bqb_ns_capacity(256) == 256
    assert L.bqb_ns_capacity(257) == 512 and L.bqb_ns_capacity(512) == 512      # generic scoring kernel only
    assert L.bqb_ns_capacity(513) == _lib.EUNSUPPORTED
    assert L.bqb_ns_capacity(0) == _lib.EINVAL
    assert L.bqb_batch_create(None, 0, 1, 8) == _lib.EINVAL
    assert L.bqb_batch_set_approx(None, 0, None, None, None, 0, 0, 0) == _lib.EINVAL
    assert b"bad arguments" in L.bqb_last_error()


def test_missing_library_fails_loudly(monkeypatch):
    from bayesian_quadrature_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", os.path.join(ROOT, "does_not_exist.so"))
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.load()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "bayesian_quadrature_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, f)) as fh:
                    src = fh.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "bq_oracle" not in src, f
