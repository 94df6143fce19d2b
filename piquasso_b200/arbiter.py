"""Accuracy arbiter: the permanent in double-double arithmetic on the GPU
(``pq_perm_arbiter_c128``, csrc/pqperm_arbiter.cu).

Not a fast path and not part of the drop-in surface: the tests and the accuracy
reports use it to arbitrate Haar-random matrices beyond n = 32, where the
reference's own C++ is wrong (src/n_aryGrayCodeCounter.hpp:179) and a CPU
computation in extended precision takes hours."""

from __future__ import annotations

import numpy as np

from . import _lib
from ._math.permanent import _check_shapes, _raise, _resolve_matrix, _resolve_mult


def permanent_dd(matrix, rows, cols):
    """``(hi, lo)``: two complex128 numbers whose exact sum is the permanent as
    computed in ~106-bit arithmetic (relative error ~1e-25 for the sizes at hand)."""
    lib = _lib.load()
    a = np.ascontiguousarray(_resolve_matrix(matrix), dtype=np.complex128)
    r = _resolve_mult(rows, "rows")
    c = _resolve_mult(cols, "cols")
    _check_shapes(a, r, c)
    out = np.zeros(4)
    rc = lib.pq_perm_arbiter_c128(
        a.ctypes.data_as(_lib.c_double_p), a.shape[0], a.shape[1],
        r.ctypes.data_as(_lib.c_int32_p), c.ctypes.data_as(_lib.c_int32_p),
        out.ctypes.data_as(_lib.c_double_p))
    _raise(rc)
    return complex(out[0], out[2]), complex(out[1], out[3])


def relerr_vs(value, hi, lo):
    """|value - (hi + lo)| / |hi + lo| without losing the low part."""
    d = (complex(value) - hi) - lo
    return abs(d) / abs(hi)
