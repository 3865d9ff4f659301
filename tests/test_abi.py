"""CPU: the C-ABI library builds, loads, and exports every symbol include/smt_b200.h declares (no compute calls)."""
import ctypes

import pytest

from stereo_match_traditional_b200 import _lib


@pytest.fixture(scope="module")
def lib():
    if not _lib.LIB_PATH.exists():
        _lib.build()
    return _lib.load()


def test_header_and_binding_table_agree():
    hs = set(_lib.header_symbols())
    assert hs, "no SMT_API declarations found"
    assert hs == set(_lib.SIGNATURES), (hs ^ set(_lib.SIGNATURES))


def test_library_exports_every_declared_symbol(lib):
    for name in _lib.header_symbols():
        assert hasattr(lib, name), name
    assert lib.smt_abi_version() == 2


def test_status_strings_and_param_defaults(lib):
    assert lib.smt_status_string(0) == b"ok"
    assert b"CUDA" in lib.smt_status_string(4)
    p = _lib.AdCensusParams()
    lib.smt_adcensus_default_params(ctypes.byref(p), 375, 450, 64)
    assert (p.rows, p.cols, p.D) == (375, 450, 64)
    assert (p.sigmaC, p.sigmaS) == (10.0, 30.0)          # main.cpp:25-26
    assert (p.L1, p.L2, p.t1, p.t2) == (34, 17, 20, 6)    # adcensus_types.h:69-70
    assert (p.p1, p.p2, p.gate, p.agg_iters) == (10, 150, 2, 4)
    assert p.speckle_invalid == 2147483647 and p.median_size == 3


def test_no_cpu_fallback(lib):
    """Without a CUDA device smt_create must fail loudly (SMT_ERR_CUDA), never fall back."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    assert lib.smt_create(0, None, ctypes.byref(h)) == 4
    assert not h.value
    from stereo_match_traditional_b200.api import Engine, SmtError

    with pytest.raises(SmtError):
        Engine(0)


def test_product_package_never_imports_oracle():
    import pathlib
    import re

    pkg = pathlib.Path(_lib.HERE)
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        assert not re.search(r"^\s*(from|import)\s+oracle|oracle_port\.h|libsmt_oracle|libsmt_ref", f.read_text(), re.M), f
