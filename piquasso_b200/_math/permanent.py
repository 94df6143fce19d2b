"""Drop-in for the reference's native module ``piquasso._math.permanent``.

Same two callables, same keyword names (``matrix``, ``rows``, ``cols``), same
overload resolution and return types as the pybind11 binding
``piquasso/_math/permanent.cpp:26-85`` of the reference, but the arithmetic
runs on the GPU through ``libpqperm.so`` (``include/pqperm.h``).

Overload rules mirrored from the binding (``py::array_t<std::complex<T>,
c_style>`` without ``forcecast``, float overload registered first):

* a ``complex64`` ndarray takes the complex64 entry, a ``complex128`` ndarray
  the complex128 entry;
* any other ndarray is converted by numpy *safe* casting, complex64 first
  (float32, float16, small ints, bool), else complex128 (float64, int32/64);
  anything that casts safely to neither raises ``TypeError``;
* a non-ndarray (nested list) is built directly as complex64, as
  ``PyArray_FromAny`` does for the first overload;
* ``rows`` / ``cols`` are force-cast to int32 (lists and tuples accepted).

The complex64 entry computes in FP64 on the device and rounds once.
"""

from __future__ import annotations

import ctypes

import numpy as np

from .. import _lib

__all__ = ["permanent", "permanent_laplace"]


def _resolve_matrix(matrix):
    if isinstance(matrix, np.ndarray):
        if matrix.dtype == np.complex64 or matrix.dtype == np.complex128:
            dtype = matrix.dtype
        elif np.can_cast(matrix.dtype, np.complex64, casting="safe"):
            dtype = np.dtype(np.complex64)
        elif np.can_cast(matrix.dtype, np.complex128, casting="safe"):
            dtype = np.dtype(np.complex128)
        else:
            raise TypeError(
                "permanent(): incompatible function arguments: matrix dtype %s does "
                "not cast safely to complex64 or complex128" % matrix.dtype)
        a = np.ascontiguousarray(matrix, dtype=dtype)
    else:
        try:
            a = np.ascontiguousarray(np.array(matrix, dtype=np.complex64))
        except (TypeError, ValueError) as exc:
            raise TypeError("permanent(): incompatible function arguments") from exc
    if a.ndim != 2:
        raise ValueError("matrix must be 2-dimensional, got ndim=%d" % a.ndim)
    return a


def _resolve_mult(v, name):
    try:
        a = np.asarray(v)
        if a.dtype != np.int32:
            a = a.astype(np.int32, casting="unsafe")
    except (TypeError, ValueError) as exc:
        raise TypeError("%s must be convertible to an int32 array" % name) from exc
    a = np.ascontiguousarray(a)
    if a.ndim != 1:
        raise ValueError("%s must be 1-dimensional" % name)
    return a


def _check_shapes(a, r, c):
    if r.shape[0] != a.shape[0] or c.shape[0] != a.shape[1]:
        raise ValueError(
            "multiplicity lengths (%d, %d) do not match the matrix shape %s"
            % (r.shape[0], c.shape[0], a.shape))


def _raise(rc):
    if rc == _lib.PQ_OK:
        return
    msg = _lib.last_error()
    if rc in (_lib.PQ_ERR_BAD_ARG, _lib.PQ_ERR_TOO_LARGE):
        raise ValueError(msg)
    # PQ_ERR_SUM_MISMATCH surfaces as RuntimeError like the reference's
    # `throw std::string` (src/permanent.cpp:100-104)
    raise _lib.PqPermError(rc, msg)


def permanent(matrix, rows, cols):
    """Calculates the permanent of a matrix, based on Eq. (8) of
    https://arxiv.org/abs/2309.07027.

    Returns a 0-d numpy array of the matrix's complex dtype (as the reference's
    ``create_numpy_scalar``, src/numpy_utils.hpp:36-49)."""
    lib = _lib.load()
    a = _resolve_matrix(matrix)
    r = _resolve_mult(rows, "rows")
    c = _resolve_mult(cols, "cols")
    _check_shapes(a, r, c)
    rp = r.ctypes.data_as(_lib.c_int32_p)
    cp = c.ctypes.data_as(_lib.c_int32_p)
    if a.dtype == np.complex64:
        out = np.zeros(2, dtype=np.float32)
        rc = lib.pq_perm_c64(a.ctypes.data_as(_lib.c_float_p), a.shape[0], a.shape[1],
                             rp, cp, out.ctypes.data_as(_lib.c_float_p))
        _raise(rc)
        return np.array(np.complex64(complex(out[0], out[1])))
    out = np.zeros(2, dtype=np.float64)
    rc = lib.pq_perm_c128(a.ctypes.data_as(_lib.c_double_p), a.shape[0], a.shape[1],
                          rp, cp, out.ctypes.data_as(_lib.c_double_p))
    _raise(rc)
    return np.array(np.complex128(complex(out[0], out[1])))


def permanent_laplace(matrix, rows, cols):
    """Calculates the permanents of the submatrices corresponding to the Laplace
    expansion, corresponding to Eq. (8) of https://arxiv.org/abs/2309.07027 and
    Lemma 1 of https://arxiv.org/abs/2005.04214.

    Returns a fresh 1-d numpy array with one entry per column (length 1 on the
    reference's empty-problem early-out, src/permanent_laplace.cpp:52-57)."""
    lib = _lib.load()
    a = _resolve_matrix(matrix)
    r = _resolve_mult(rows, "rows")
    c = _resolve_mult(cols, "cols")
    _check_shapes(a, r, c)
    rp = r.ctypes.data_as(_lib.c_int32_p)
    cp = c.ctypes.data_as(_lib.c_int32_p)
    n = ctypes.c_int(0)
    width = max(a.shape[1], 1)
    if a.dtype == np.complex64:
        out = np.zeros(2 * width, dtype=np.float32)
        rc = lib.pq_perm_laplace_c64(a.ctypes.data_as(_lib.c_float_p), a.shape[0],
                                     a.shape[1], rp, cp,
                                     out.ctypes.data_as(_lib.c_float_p), ctypes.byref(n))
        _raise(rc)
        return out[: 2 * n.value].view(np.complex64).copy()
    out = np.zeros(2 * width, dtype=np.float64)
    rc = lib.pq_perm_laplace_c128(a.ctypes.data_as(_lib.c_double_p), a.shape[0],
                                  a.shape[1], rp, cp,
                                  out.ctypes.data_as(_lib.c_double_p), ctypes.byref(n))
    _raise(rc)
    return out[: 2 * n.value].view(np.complex128).copy()
