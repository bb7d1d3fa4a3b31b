"""ctypes binding of ``libskb.so`` (C ABI in ``include/skb.h``).

This is the stub a maintainer of the reference would add next to ``skysampler/emulator.py`` to
reach the B200 backend (INTEGRATION.md shows it in reference terms).  It holds no compute: every
function forwards plain pointers and sizes.  There is no fallback — if the shared library is
missing the import fails loudly, and if no CUDA device is present every compute call raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SKB_LIB_PATH") or os.path.join(_HERE, "csrc", "libskb.so")

SKB_MAXD = 16
SKB_MAX_GROUP = 8
SKB_MAX_JOBS = 16
SKB_F64, SKB_F32 = 0, 1
SKB_EINVAL, SKB_ECUDA, SKB_ENOMEM, SKB_ESTATE = -1, -2, -3, -4


class SkbJob(C.Structure):
    _fields_ = [
        ("kde", C.c_void_p),
        ("q", C.c_void_p),
        ("q_dtype", C.c_int32),
        ("q_is_whitened", C.c_int32),
        ("M", C.c_int64),
        ("ldq", C.c_int64),
        ("cols", C.c_void_p),
        ("out_logp", C.c_void_p),
    ]


class SkbFeatureSpec(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("b", C.c_void_p),
        ("a_stride", C.c_int64), ("b_stride", C.c_int64),
        ("a_const", C.c_double), ("b_const", C.c_double),
        ("lo", C.c_double), ("hi", C.c_double),
        ("a_dtype", C.c_int32), ("b_dtype", C.c_int32),
        ("op", C.c_int32), ("log10", C.c_int32),
    ]


# name -> (restype, argtypes); the single source of truth the symbol-export test walks too
SIGNATURES = {
    "skb_last_error": (C.c_char_p, []),
    "skb_version": (C.c_int, []),
    "skb_device_count": (C.c_int, []),
    "skb_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int64, C.c_int,
                             C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                             C.c_int, C.POINTER(C.c_void_p)]),
    "skb_destroy": (C.c_int, [C.c_void_p]),
    "skb_n_train": (C.c_int64, [C.c_void_p]),
    "skb_ndim": (C.c_int, [C.c_void_p]),
    "skb_sum_weight": (C.c_double, [C.c_void_p]),
    "skb_device": (C.c_int, [C.c_void_p]),
    "skb_fit_ms": (C.c_double, [C.c_void_p]),
    "skb_set_option": (C.c_int, [C.c_char_p, C.c_int]),
    "skb_get_option": (C.c_int, [C.c_char_p]),
    "skb_score": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_int,
                            C.c_void_p, C.c_void_p]),
    "skb_score_partial": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "skb_log_knorm": (C.c_double, [C.c_int, C.c_double]),
    "skb_score_batch": (C.c_int, [C.POINTER(SkbJob), C.c_int, C.c_void_p]),
    "skb_ratio_accept": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_void_p, C.c_double,
                                   C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p,
                                   C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p]),
    "skb_draw": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64,
                           C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int,
                           C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                           C.POINTER(C.c_int64), C.c_void_p]),
    "skb_compact": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p,
                              C.POINTER(C.c_int64), C.c_void_p]),
    "skb_features": (C.c_int, [C.POINTER(SkbFeatureSpec), C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "skb_measure_pipes": (C.c_int, [C.c_int, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "skb_debug_counters": (C.c_int, [C.c_int, C.POINTER(C.c_uint64), C.c_int]),
    "skb_launch_count": (C.c_int64, []),
}

_lib = None


def lib():
    """The loaded library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "skysampler_b200: %s is missing — build it with `make -C skysampler_b200/csrc` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`); there is no CPU fallback"
                % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


class SkbError(RuntimeError):
    pass


def check(rc):
    """0 -> None; SKB_EINVAL -> ValueError (as sklearn's validation would); else RuntimeError."""
    if rc == 0:
        return
    msg = lib().skb_last_error().decode("utf-8", "replace")
    if rc == SKB_EINVAL:
        raise ValueError(msg)
    if rc == SKB_ENOMEM:
        raise MemoryError(msg)
    raise SkbError("libskb error %d: %s" % (rc, msg))


def _is_torch(x):
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


def ptr(x):
    """Raw address of a numpy array / torch tensor (host or CUDA), or None."""
    if x is None:
        return None
    if _is_torch(x):
        return C.c_void_p(x.data_ptr())
    if isinstance(x, np.ndarray):
        return C.c_void_p(x.ctypes.data)
    raise TypeError("expected a numpy array or a torch tensor, got %r" % type(x))


def dtype_code(x):
    name = str(x.dtype).replace("torch.", "")
    if name == "float64":
        return SKB_F64
    if name == "float32":
        return SKB_F32
    raise TypeError("only float64 / float32 matrices are supported, got %s" % x.dtype)


def stream_ptr(stream=None):
    """cudaStream_t of a torch stream (or of torch's current stream) as void*; None -> legacy default."""
    if stream is None:
        return None
    return C.c_void_p(int(stream.cuda_stream))


def cols_array(cols, d, ld=None):
    if cols is None:
        return None
    a = np.ascontiguousarray(np.asarray(cols, dtype=np.int32))
    if a.shape != (d,):
        raise ValueError("need %d column indices, got shape %s" % (d, a.shape))
    if ld is not None and (a.min() < 0 or a.max() >= ld):
        raise ValueError("column indices %s out of range for a table with %d columns" % (a.tolist(), ld))
    return a


def set_option(name, value):
    """Process-wide diagnostic switch of the library (include/skb.h: "cull", "force_nsplit", "stats")."""
    check(lib().skb_set_option(name.encode(), int(value)))


class option(object):
    """``with option("cull", 0): ...`` — set a diagnostic switch for a block and restore it afterwards."""

    def __init__(self, name, value):
        self.name, self.value = name, int(value)

    def __enter__(self):
        self.prev = lib().skb_get_option(self.name.encode())
        set_option(self.name, self.value)
        return self

    def __exit__(self, *exc):
        set_option(self.name, self.prev)
        return False
