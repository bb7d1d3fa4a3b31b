"""ctypes front end of ``oracle/brute64.c`` — the fp64 brute-force oracle at full training-set sizes.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  ``score`` / ``partial_lse`` have the signatures
of ``oracle.brute64`` and are pinned against it by ``tests/test_oracle.py``; the shared object is built
by ``__graft_entry__.build()`` (``gcc -O2 -fopenmp``) and, if it is missing, on first use.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import brute64

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "brute64.c")
LIB = os.path.join(_HERE, "_brute64.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"])
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        h = C.CDLL(LIB)
        h.brute64_partial_lse.restype = None
        h.brute64_partial_lse.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_double,
                                          C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        _lib = h
    return _lib


def partial_lse(train, weights, h, query):
    """Same contract as ``brute64.partial_lse`` (natural-log ``m``, ``s``), all host cores."""
    train = np.ascontiguousarray(train, dtype=np.float64)
    query = np.ascontiguousarray(query, dtype=np.float64)
    n, d = train.shape
    logw = None
    if weights is not None:
        with np.errstate(divide="ignore"):
            logw = np.ascontiguousarray(np.log(np.asarray(weights, dtype=np.float64)))
    m_out = np.empty(len(query))
    s_out = np.empty(len(query))
    lib().brute64_partial_lse(train.ctypes.data, None if logw is None else logw.ctypes.data, n, d, float(h),
                              query.ctypes.data, len(query), m_out.ctypes.data, s_out.ctypes.data)
    return m_out, s_out


def score(train, weights, h, query):
    """Same contract as ``brute64.score``."""
    train = np.asarray(train, dtype=np.float64)
    n, d = train.shape
    m, s = partial_lse(train, weights, h, query)
    sum_w = float(n) if weights is None else float(np.sum(np.asarray(weights, dtype=np.float64)))
    with np.errstate(divide="ignore"):
        return brute64.log_knorm(d, h) + m + np.log(s) - np.log(sum_w)
