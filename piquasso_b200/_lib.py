"""ctypes binding of ``libpqperm.so`` (the C ABI in ``include/pqperm.h``).

The library is built in-tree by :mod:`piquasso_b200.build`.  There is no
Python or CPU implementation behind these calls: if the shared object is
missing, or no CUDA device is usable, the compute entry points raise.
"""

from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PQ_LIB_PATH") or os.path.join(_HERE, "libpqperm.so")

PQ_OK = 0
PQ_ERR_SUM_MISMATCH = 1
PQ_ERR_BAD_ARG = 2
PQ_ERR_NO_DEVICE = 3
PQ_ERR_CUDA = 4
PQ_ERR_TOO_LARGE = 5

c_double_p = ctypes.POINTER(ctypes.c_double)
c_float_p = ctypes.POINTER(ctypes.c_float)
c_int32_p = ctypes.POINTER(ctypes.c_int32)
c_int64_p = ctypes.POINTER(ctypes.c_int64)


class PlanInfo(ctypes.Structure):
    """``pq_plan_info`` of include/pqperm.h."""

    _fields_ = [
        ("idx_max", ctypes.c_int64),
        ("seg_len", ctypes.c_int64),
        ("nseg", ctypes.c_int64),
        ("active_rows", ctypes.c_int32),
        ("active_cols", ctypes.c_int32),
        ("low_digits", ctypes.c_int32),
        ("kernel", ctypes.c_int32),
        ("cols_padded", ctypes.c_int32),
        ("sum_rows", ctypes.c_int32),
        ("trivial", ctypes.c_int32),
        ("flops_per_term", ctypes.c_double),
    ]


# every symbol include/pqperm.h declares: (name, restype, argtypes)
SIGNATURES = [
    ("pq_last_error", ctypes.c_char_p, []),
    ("pq_device_count", ctypes.c_int, []),
    ("pq_set_devices", ctypes.c_int, [c_int32_p, ctypes.c_int]),
    ("pq_perm_c128", ctypes.c_int,
     [c_double_p, ctypes.c_int, ctypes.c_int, c_int32_p, c_int32_p, c_double_p]),
    ("pq_perm_c64", ctypes.c_int,
     [c_float_p, ctypes.c_int, ctypes.c_int, c_int32_p, c_int32_p, c_float_p]),
    ("pq_perm_laplace_c128", ctypes.c_int,
     [c_double_p, ctypes.c_int, ctypes.c_int, c_int32_p, c_int32_p, c_double_p,
      ctypes.POINTER(ctypes.c_int)]),
    ("pq_perm_laplace_partial_c128", ctypes.c_int,
     [c_double_p, ctypes.c_int, ctypes.c_int, c_int32_p, c_int32_p, ctypes.c_int, ctypes.c_int,
      c_double_p, ctypes.POINTER(ctypes.c_int)]),
    ("pq_perm_laplace_partial_dev_c128", ctypes.c_int,
     [c_double_p, ctypes.c_int, ctypes.c_int, c_int32_p, c_int32_p, ctypes.c_int, ctypes.c_int,
      ctypes.c_int, ctypes.c_void_p, c_double_p, ctypes.POINTER(ctypes.c_int)]),
    ("pq_perm_laplace_c64", ctypes.c_int,
     [c_float_p, ctypes.c_int, ctypes.c_int, c_int32_p, c_int32_p, c_float_p,
      ctypes.POINTER(ctypes.c_int)]),
    ("pq_perm_laplace_batch_c128", ctypes.c_int,
     [ctypes.c_int, c_double_p, c_int64_p, c_int32_p, c_int32_p, c_int32_p, c_int64_p,
      c_int32_p, c_int64_p, c_double_p, c_int64_p, c_int32_p]),
    ("pq_perm_batch_c128", ctypes.c_int,
     [c_double_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_int32_p, c_int32_p, c_double_p]),
    ("pq_sampler_pmf_c128", ctypes.c_int,
     [c_double_p, ctypes.c_int, ctypes.c_int, c_int32_p, c_int32_p, c_double_p]),
    ("pq_sampler_draw_c128", ctypes.c_int,
     [c_double_p, ctypes.c_int, ctypes.c_int, c_int32_p, c_int32_p, c_double_p, c_int32_p]),
    ("pq_pcg64_streams", ctypes.c_int,
     [ctypes.c_uint64, ctypes.c_int64, ctypes.c_int, ctypes.POINTER(ctypes.c_uint64)]),
    ("pq_last_sampler_profile", None, [c_double_p]),
    ("pq_sampler_draw_dev_c128", ctypes.c_int,
     [ctypes.c_int, c_double_p, ctypes.c_int, ctypes.c_int, c_int32_p, c_int32_p, c_double_p,
      c_int32_p]),
    ("pq_sampler_pmf_dev_c128", ctypes.c_int,
     [ctypes.c_int, c_double_p, ctypes.c_int, ctypes.c_int, c_int32_p, c_int32_p, c_double_p]),
    ("pq_perm_partial_c128", ctypes.c_int,
     [c_double_p, ctypes.c_int, ctypes.c_int, c_int32_p, c_int32_p, ctypes.c_int,
      ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
      ctypes.POINTER(ctypes.c_int), c_double_p]),
    ("pq_perm_job_create_c128", ctypes.c_int,
     [c_double_p, ctypes.c_int, ctypes.c_int, c_int32_p, c_int32_p, ctypes.c_int,
      ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p),
      ctypes.POINTER(ctypes.c_int), c_double_p]),
    ("pq_perm_job_launch", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    ("pq_perm_job_info", ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(PlanInfo)]),
    ("pq_perm_job_terms", ctypes.c_int64, [ctypes.c_void_p]),
    ("pq_perm_job_destroy", ctypes.c_int, [ctypes.c_void_p]),
    ("pq_kernel_ms_history", ctypes.c_int, [ctypes.c_int, c_double_p, ctypes.c_int]),
    ("pq_perm_combine", ctypes.c_int, [c_double_p, ctypes.c_int, c_double_p]),
    ("pq_perm_finish", ctypes.c_int, [c_double_p, ctypes.c_int, c_double_p]),
    ("pq_perm_plan", ctypes.c_int,
     [ctypes.c_int, ctypes.c_int, c_int32_p, c_int32_p, ctypes.POINTER(PlanInfo)]),
    ("pq_perm_batch_plan", ctypes.c_int,
     [ctypes.c_int, ctypes.c_int, c_int32_p, c_int32_p, ctypes.c_int, ctypes.POINTER(PlanInfo)]),
    ("pq_perm_gray_of_offset", ctypes.c_int,
     [ctypes.c_int, c_int32_p, ctypes.c_int64, c_int32_p]),
    ("pq_perm_segment_sums_c128", ctypes.c_int,
     [c_double_p, ctypes.c_int, ctypes.c_int, c_int32_p, c_int32_p, ctypes.c_int64,
      ctypes.c_int64, c_double_p]),
    ("pq_last_kernel_ms", ctypes.c_double, [ctypes.c_int]),
    ("pq_launch_count", ctypes.c_int64, []),
    ("pq_perm_arbiter_c128", ctypes.c_int, [c_double_p, ctypes.c_int, ctypes.c_int, c_int32_p,
                                            c_int32_p, c_double_p]),
    ("pq_last_sampler_detail", None, [c_double_p]),
    ("pq_sampler_work", None, [c_double_p]),
    ("pq_sampler_work_reset", None, []),
    ("pq_fp64_peak_tflops", ctypes.c_double, [ctypes.c_int, ctypes.c_int]),
    ("pq_set_kernel_choice", ctypes.c_int, [ctypes.c_int]),
    ("pq_set_timing", ctypes.c_int, [ctypes.c_int]),
    ("pq_set_seg_len_hint", ctypes.c_int, [ctypes.c_int64]),
]

_lib = None


class PqPermError(RuntimeError):
    """A libpqperm call failed (``code`` is the PQ_ERR_* value)."""

    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code


def load() -> ctypes.CDLL:
    """Load libpqperm.so and bind every declared symbol (raises if absent)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s is missing: build it with `python -m piquasso_b200.build` "
                "(there is no CPU fallback for the permanent path)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, restype, argtypes in SIGNATURES:
            fn = getattr(lib, name)  # AttributeError = ABI drift, fail loudly
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def last_error() -> str:
    return load().pq_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != PQ_OK:
        raise PqPermError(rc, last_error() or ("libpqperm error %d" % rc))
