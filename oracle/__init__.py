"""CPU checker for the permanent hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package, and only
as the checker.  ``piquasso_b200`` never imports it.

Three independent statements of the same function live here:

* :func:`permanent` / :func:`permanent_laplace` -- ctypes over
  ``oracle/perm_oracle.c`` (C restatement of ``src/permanent.cpp``,
  ``src/permanent_laplace.cpp``, ``src/n_aryGrayCodeCounter.hpp`` with 64-bit
  offsets; ``precision=1`` is the long-double arbiter);
* :func:`ref_permanent` / :func:`ref_permanent_laplace` -- the unmodified
  reference C++ compiled into ``oracle/_ref/libpqref.so`` (valid for
  ``idx_max <= 2**31``);
* :func:`permanent_definition` -- the textbook sum over permutations (numpy,
  tiny sizes), which shares nothing with the Glynn/Gray-code machinery.
"""

from __future__ import annotations

import ctypes
import itertools
import os

import numpy as np

from . import build as _build

_c_int_p = ctypes.POINTER(ctypes.c_int)
_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_float_p = ctypes.POINTER(ctypes.c_float)

_oracle_lib = None
_ref_lib = None


def _load_oracle():
    global _oracle_lib
    if _oracle_lib is None:
        path = _build.build_oracle()
        lib = ctypes.CDLL(path)
        lib.pqo_permanent_c128.restype = ctypes.c_int
        lib.pqo_permanent_c128.argtypes = [
            _c_double_p, ctypes.c_int, ctypes.c_int, _c_int_p, _c_int_p,
            ctypes.c_int, ctypes.c_int, _c_double_p]
        lib.pqo_permanent_c128_hilo.restype = ctypes.c_int
        lib.pqo_permanent_c128_hilo.argtypes = [
            _c_double_p, ctypes.c_int, ctypes.c_int, _c_int_p, _c_int_p,
            ctypes.c_int, ctypes.c_int, _c_double_p]
        lib.pqo_permanent_laplace_c128.restype = ctypes.c_int
        lib.pqo_permanent_laplace_c128.argtypes = [
            _c_double_p, ctypes.c_int, ctypes.c_int, _c_int_p, _c_int_p,
            ctypes.c_int, ctypes.c_int, _c_double_p, _c_int_p]
        lib.pqo_partial_c128.restype = ctypes.c_int
        lib.pqo_partial_c128.argtypes = [
            _c_double_p, ctypes.c_int, ctypes.c_int, _c_int_p, _c_int_p,
            ctypes.c_int64, ctypes.c_int64, ctypes.c_int, _c_double_p,
            ctypes.POINTER(ctypes.c_int64)]
        lib.pqo_gray_of_offset.restype = ctypes.c_int
        lib.pqo_gray_of_offset.argtypes = [
            ctypes.c_int, _c_int_p, ctypes.c_int64, _c_int_p, _c_int_p,
            ctypes.POINTER(ctypes.c_int64)]
        lib.pqo_gray_init.restype = None
        lib.pqo_gray_init.argtypes = [
            _c_int_p, ctypes.c_int, ctypes.c_int64, _c_int_p, _c_int_p]
        lib.pqo_gray_next.restype = ctypes.c_int
        lib.pqo_gray_next.argtypes = [
            _c_int_p, ctypes.c_int, _c_int_p, _c_int_p, _c_int_p, _c_int_p, _c_int_p]
        lib.pqo_num_threads.restype = ctypes.c_int
        _oracle_lib = lib
    return _oracle_lib


def ref_available() -> bool:
    return _build.build_ref() is not None


def _load_ref():
    global _ref_lib
    if _ref_lib is None:
        path = _build.build_ref()
        if path is None:
            raise RuntimeError(
                "oracle/_ref/libpqref.so is absent and the reference sources are "
                "not present to build it")
        lib = ctypes.CDLL(path)
        lib.pqref_permanent_c128.restype = ctypes.c_int
        lib.pqref_permanent_c128.argtypes = [
            _c_double_p, ctypes.c_int, ctypes.c_int, _c_int_p, _c_int_p, _c_double_p]
        lib.pqref_permanent_c64.restype = ctypes.c_int
        lib.pqref_permanent_c64.argtypes = [
            _c_float_p, ctypes.c_int, ctypes.c_int, _c_int_p, _c_int_p, _c_float_p]
        lib.pqref_permanent_laplace_c128.restype = ctypes.c_int
        lib.pqref_permanent_laplace_c128.argtypes = [
            _c_double_p, ctypes.c_int, ctypes.c_int, _c_int_p, _c_int_p,
            _c_double_p, _c_int_p]
        lib.pqref_permanent_laplace_c64.restype = ctypes.c_int
        lib.pqref_permanent_laplace_c64.argtypes = [
            _c_float_p, ctypes.c_int, ctypes.c_int, _c_int_p, _c_int_p,
            _c_float_p, _c_int_p]
        lib.pqref_gray_trace.restype = ctypes.c_int
        lib.pqref_gray_trace.argtypes = [
            _c_int_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int, _c_int_p, _c_int_p]
        _ref_lib = lib
    return _ref_lib


def _prep(matrix, rows, cols, dtype=np.complex128):
    a = np.ascontiguousarray(np.asarray(matrix), dtype=dtype)
    if a.ndim != 2:
        raise ValueError("matrix must be 2-D")
    r = np.ascontiguousarray(np.asarray(rows), dtype=np.int32).reshape(-1)
    c = np.ascontiguousarray(np.asarray(cols), dtype=np.int32).reshape(-1)
    return a, r, c


def _ip(a):
    return a.ctypes.data_as(_c_int_p)


def num_threads() -> int:
    return int(_load_oracle().pqo_num_threads())


def permanent(matrix, rows, cols, njobs: int = 32, precision: int = 0) -> complex:
    """C restatement of ``permanent_cpp<double>`` (src/permanent.cpp:49-264)."""
    a, r, c = _prep(matrix, rows, cols)
    out = np.zeros(2)
    rc = _load_oracle().pqo_permanent_c128(
        a.ctypes.data_as(_c_double_p), a.shape[0], a.shape[1], _ip(r), _ip(c),
        njobs, precision, out.ctypes.data_as(_c_double_p))
    if rc == 1:
        raise RuntimeError("Number of input and output states should be equal")
    if rc:
        raise ValueError("bad arguments")
    return complex(out[0], out[1])


def permanent_hilo(matrix, rows, cols, njobs: int = 32, precision: int = 2):
    """The permanent in extended precision as a (hi, lo) pair of complex128
    (value = hi + lo): ``precision=1`` long double, ``precision=2`` software
    binary128 (``__float128``, ~50x slower than double -- fixtures only)."""
    a, r, c = _prep(matrix, rows, cols)
    out = np.zeros(4)
    rc = _load_oracle().pqo_permanent_c128_hilo(
        a.ctypes.data_as(_c_double_p), a.shape[0], a.shape[1], _ip(r), _ip(c),
        njobs, precision, out.ctypes.data_as(_c_double_p))
    if rc == 1:
        raise RuntimeError("Number of input and output states should be equal")
    if rc:
        raise ValueError("bad arguments")
    return complex(out[0], out[2]), complex(out[1], out[3])


def permanent_laplace(matrix, rows, cols, njobs: int = 32, precision: int = 0):
    """C restatement of ``permanent_laplace_cpp<double>``
    (src/permanent_laplace.cpp:43-240)."""
    a, r, c = _prep(matrix, rows, cols)
    out = np.zeros(2 * max(a.shape[1], 1))
    n = ctypes.c_int(0)
    rc = _load_oracle().pqo_permanent_laplace_c128(
        a.ctypes.data_as(_c_double_p), a.shape[0], a.shape[1], _ip(r), _ip(c),
        njobs, precision, out.ctypes.data_as(_c_double_p), ctypes.byref(n))
    if rc:
        raise ValueError("bad arguments")
    return out[: 2 * n.value].view(np.complex128).copy()


def partial(matrix, rows, cols, begin: int, end: int, laplace: bool = False):
    """Unscaled long-double sum of term(offset), offset in [begin, end).

    Returns (values, idx_max): values is complex (hi+lo folded) of length 1 or C
    together with the raw hi/lo quadruples."""
    a, r, c = _prep(matrix, rows, cols)
    nacc = a.shape[1] if laplace else 1
    out = np.zeros(4 * nacc)
    idx_max = ctypes.c_int64(0)
    rc = _load_oracle().pqo_partial_c128(
        a.ctypes.data_as(_c_double_p), a.shape[0], a.shape[1], _ip(r), _ip(c),
        begin, end, int(laplace), out.ctypes.data_as(_c_double_p),
        ctypes.byref(idx_max))
    if rc:
        raise ValueError("bad arguments")
    q = out.reshape(nacc, 4)
    vals = (q[:, 0] + q[:, 1]) + 1j * (q[:, 2] + q[:, 3])
    return vals, q, int(idx_max.value)


def gray_of_offset(rows, offset: int):
    """Gray digits (after the reference's row split) of ``offset``; also the
    digit limits and idx_max.  Digit i belongs to split-row i+1."""
    r = np.ascontiguousarray(np.asarray(rows), dtype=np.int32).reshape(-1)
    gray = np.zeros(max(len(r), 1), dtype=np.int32)
    limits = np.zeros(max(len(r), 1), dtype=np.int32)
    idx_max = ctypes.c_int64(0)
    n = _load_oracle().pqo_gray_of_offset(
        len(r), _ip(r), offset, _ip(gray), _ip(limits), ctypes.byref(idx_max))
    if n < 0:
        raise ValueError("no non-zero row multiplicity")
    return gray[:n].copy(), limits[:n].copy(), int(idx_max.value)


def gray_trace(limits, offset: int, nsteps: int):
    """(gray0, [(changed, prev, value)...]) from the restated counter."""
    lib = _load_oracle()
    lim = np.ascontiguousarray(np.asarray(limits), dtype=np.int32)
    d = len(lim)
    chain = np.zeros(d, dtype=np.int32)
    gray = np.zeros(d, dtype=np.int32)
    lib.pqo_gray_init(_ip(lim), d, offset, _ip(chain), _ip(gray))
    gray0 = gray.copy()
    trace = []
    ch, pv, vl = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    for _ in range(nsteps):
        if lib.pqo_gray_next(_ip(lim), d, _ip(chain), _ip(gray),
                             ctypes.byref(ch), ctypes.byref(pv), ctypes.byref(vl)):
            break
        trace.append((ch.value, pv.value, vl.value))
    return gray0, trace


# ---------------------------------------------------------------- reference


def ref_permanent(matrix, rows, cols) -> complex:
    """The unmodified reference ``permanent_cpp<double>``."""
    a, r, c = _prep(matrix, rows, cols)
    out = np.zeros(2)
    rc = _load_ref().pqref_permanent_c128(
        a.ctypes.data_as(_c_double_p), a.shape[0], a.shape[1], _ip(r), _ip(c),
        out.ctypes.data_as(_c_double_p))
    if rc:
        raise RuntimeError("Caught an unknown exception!")
    return complex(out[0], out[1])


def ref_permanent_c64(matrix, rows, cols) -> complex:
    a, r, c = _prep(matrix, rows, cols, dtype=np.complex64)
    out = np.zeros(2, dtype=np.float32)
    rc = _load_ref().pqref_permanent_c64(
        a.ctypes.data_as(_c_float_p), a.shape[0], a.shape[1], _ip(r), _ip(c),
        out.ctypes.data_as(_c_float_p))
    if rc:
        raise RuntimeError("Caught an unknown exception!")
    return complex(out[0], out[1])


def ref_permanent_laplace(matrix, rows, cols):
    """The unmodified reference ``permanent_laplace_cpp<double>``."""
    a, r, c = _prep(matrix, rows, cols)
    out = np.zeros(2 * max(a.shape[1], 1))
    n = ctypes.c_int(0)
    _load_ref().pqref_permanent_laplace_c128(
        a.ctypes.data_as(_c_double_p), a.shape[0], a.shape[1], _ip(r), _ip(c),
        out.ctypes.data_as(_c_double_p), ctypes.byref(n))
    return out[: 2 * n.value].view(np.complex128).copy()


def ref_gray_trace(limits, offset: int, nsteps: int):
    """Same as :func:`gray_trace` from the reference's own counter class."""
    lim = np.ascontiguousarray(np.asarray(limits), dtype=np.int32)
    d = len(lim)
    gray0 = np.zeros(d, dtype=np.int32)
    trace = np.zeros(3 * max(nsteps, 1), dtype=np.int32)
    done = _load_ref().pqref_gray_trace(_ip(lim), d, offset, nsteps, _ip(gray0), _ip(trace))
    return gray0, [tuple(int(x) for x in trace[3 * i: 3 * i + 3]) for i in range(done)]


# ------------------------------------------------- independent definition


def assym_reduce(matrix, row_mult, col_mult):
    """Repeat rows/columns by multiplicity (what piquasso/_math/linalg.py:98-117
    does with np.repeat); host helper for the definition-based check."""
    a = np.asarray(matrix)
    return np.repeat(np.repeat(a, np.asarray(row_mult), axis=0),
                     np.asarray(col_mult), axis=1)


def permanent_definition(matrix, rows=None, cols=None) -> complex:
    """perm(A) = sum over permutations of prod a[i, sigma(i)] on the expanded
    matrix.  O(n!) -- n <= 8."""
    a = np.asarray(matrix, dtype=np.complex128)
    if rows is not None:
        a = assym_reduce(a, rows, cols)
    n, m = a.shape
    if n != m:
        raise ValueError("expanded matrix must be square")
    if n == 0:
        return 1.0 + 0.0j
    total = 0.0j
    idx = np.arange(n)
    for sigma in itertools.permutations(range(n)):
        total += np.prod(a[idx, list(sigma)])
    return complex(total)
