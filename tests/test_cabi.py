"""The C-ABI library: loads, exports every symbol include/skb.h declares, fails loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from skysampler_b200 import _cabi
from oracle import brute64


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "skb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(skb_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    lib = _cabi.lib()
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), "libskb.so does not export %s" % n
        assert n in _cabi.SIGNATURES, "no ctypes signature for %s" % n
    assert set(_cabi.SIGNATURES) == set(names)


def test_constants_match_header():
    text = open(os.path.join(ROOT, "include", "skb.h")).read()
    for name in ("SKB_MAXD", "SKB_MAX_GROUP", "SKB_MAX_JOBS", "SKB_F64", "SKB_F32", "SKB_EINVAL", "SKB_ECUDA"):
        m = re.search(r"#define\s+%s\s+(-?\d+)" % name, text)
        assert m and int(m.group(1)) == getattr(_cabi, name)


def test_log_knorm_matches_oracle():
    lib = _cabi.lib()
    for d, h in ((1, 0.05), (4, 0.1), (8, 0.1), (16, 0.3)):
        assert abs(lib.skb_log_knorm(d, h) - brute64.log_knorm(d, h)) < 1e-12


def test_argument_validation_needs_no_gpu():
    lib = _cabi.lib()
    h = C.c_void_p()
    x = np.zeros((4, 2))
    rc = lib.skb_create(_cabi.ptr(x), 0, 4, 0, 2, 1, None, None, 0.1, None, None, None, 0, C.byref(h))
    assert rc == _cabi.SKB_EINVAL and b"dimension" in lib.skb_last_error()
    rc = lib.skb_create(_cabi.ptr(x), 0, 4, 2, 2, 1, None, None, -1.0, None, None, None, 0, C.byref(h))
    assert rc == _cabi.SKB_EINVAL and b"bandwidth" in lib.skb_last_error()
    rc = lib.skb_create(_cabi.ptr(x), 0, 0, 2, 2, 1, None, None, 0.1, None, None, None, 0, C.byref(h))
    assert rc == _cabi.SKB_EINVAL
    with pytest.raises(ValueError):
        _cabi.check(rc)


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, not compute on the host."""
    lib = _cabi.lib()
    if lib.skb_device_count() > 0:
        pytest.skip("a GPU is present")
    from skysampler_b200 import GaussianKDE
    with pytest.raises(RuntimeError, match="no CUDA device"):
        GaussianKDE(0.1).fit(np.zeros((10, 2)))
    fma, ex2 = C.c_double(), C.c_double()
    assert lib.skb_measure_pipes(0, 1.0, C.byref(fma), C.byref(ex2)) == _cabi.SKB_ECUDA


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "skysampler_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert not re.search(r"^\s*(from|import)\s+sklearn\.neighbors", src, flags=re.M), f
                assert "KernelDensity(" not in src.replace("KernelDensity(...)", ""), f


def test_diagnostic_options_round_trip_without_a_gpu():
    """skb_set_option / skb_get_option are host-side switches (include/skb.h): every documented name is known, values
    are normalised, unknown names fail with SKB_EINVAL and -1."""
    lib = _cabi.lib()
    header = open(os.path.join(ROOT, "include", "skb.h")).read()
    for name, probe, back in (("cull", 0, 0), ("cull", 7, 1), ("force_nsplit", 5, 5), ("force_nsplit", -3, 0),
                              ("stats", 1, 1), ("stats", 0, 0), ("kernel", 2, 2), ("kernel", 0, 0),
                              ("ff2_cells", 0, 0), ("ff2_cells", 9, 1), ("ff2_cells", -1, -1),
                              ("host_stage_rows", 4096, 4096), ("host_stage_rows", 0, 0)):
        assert '"%s"' % name in header, name
        old = lib.skb_get_option(name.encode())
        assert lib.skb_set_option(name.encode(), probe) == 0
        assert lib.skb_get_option(name.encode()) == back, (name, probe)
        assert lib.skb_set_option(name.encode(), old) == 0
    assert lib.skb_set_option(b"no_such_switch", 1) == _cabi.SKB_EINVAL and b"unknown option" in lib.skb_last_error()
    assert lib.skb_get_option(b"no_such_switch") == -1
    with _cabi.option("cull", 0):
        assert lib.skb_get_option(b"cull") == 0
    assert lib.skb_get_option(b"cull") == 1
