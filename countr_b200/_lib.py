"""ctypes binding of libcountr_sm100.so (the C ABI declared in include/countr_b200.h).

There is no fallback: if the shared library is missing or the device is not sm_100 the
import-time / call-time checks raise.  Nothing here touches the CPU test infrastructure.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# COUNTR_B200_LIB points at an alternative build of the same ABI (A/B runs of two kernel versions on one box)
LIB_PATH = os.environ.get("COUNTR_B200_LIB") or os.path.join(_HERE, "lib", "libcountr_sm100.so")


class CountrError(RuntimeError):
    pass


class GemmDesc(Structure):
    """Mirror of `countr_gemm_desc` (include/countr_b200.h)."""

    _fields_ = [
        ("a", c_void_p), ("b", c_void_p),
        ("lda", c_int64), ("sa1", c_int64), ("sa2", c_int64),
        ("ldb", c_int64), ("sb1", c_int64), ("sb2", c_int64),
        ("a_mn", c_int32), ("b_mn", c_int32),
        ("M", c_int32), ("N", c_int32), ("K", c_int32),
        ("nb1", c_int32), ("nb2", c_int32),
        ("bf16", c_int32),
        ("bn", c_int32), ("split_k", c_int32), ("cluster", c_int32), ("cta_pair", c_int32),
        ("conv_h", c_int32), ("conv_w", c_int32), ("conv_cin", c_int32), ("conv_bx", c_int32), ("conv_by", c_int32),
        ("conv_dw", c_int32), ("conv_batch", c_int32),
        ("c", c_void_p),
        ("ldc", c_int64), ("sc1", c_int64), ("sc2", c_int64),
        ("out_f32", c_int32), ("atomic", c_int32),
        ("alpha", c_float),
        ("bias", c_void_p),
        ("act", c_int32),
        ("aux", c_void_p),
        ("ldaux", c_int64),
        ("residual", c_void_p),
        ("ldr", c_int64),
        ("res_mod", c_int32),
        ("gn_stats", c_void_p),
        ("ln_x16", c_void_p), ("ld_x16", c_int64), ("ln_stats", c_void_p), ("ln_colsum", c_void_p), ("ln_dim", c_int32), ("ln_eps", c_float),
    ]


_lib = None


def _declare(lib):
    lib.countr_last_error.restype = c_char_p
    lib.countr_version.restype = c_char_p
    lib.countr_check_device.restype = c_int32
    lib.countr_num_sms.restype = c_int32
    lib.countr_set_sm_budget.argtypes = [c_int32]
    lib.countr_set_sm_budget.restype = c_int32
    lib.countr_gemm.argtypes = [POINTER(GemmDesc), c_void_p]
    lib.countr_gemm.restype = c_int32
    lib.countr_weight_refresh_blocks.argtypes = [c_int32, c_int64, c_int64]
    lib.countr_weight_refresh_blocks.restype = c_int64
    from . import _sigs  # noqa: WPS433  (plain-argument entry points)
    _sigs.declare(lib)


def lib():
    """Return the loaded library (loads on first use)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CountrError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C countr_b200/csrc). countr_b200 has no CPU or PyTorch fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        _declare(handle)
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise CountrError(f"countr_b200 kernel call failed ({rc}): {lib().countr_last_error().decode()}")


def require_device():
    check(lib().countr_check_device())
