"""Introspection of the partition the GPU path uses (no GPU needed).

Thin wrappers over ``pq_perm_plan`` / ``pq_perm_gray_of_offset``
(include/pqperm.h): which Gray-code terms exist, how they are cut into
segments and which rank owns which segments.
"""

from __future__ import annotations

import ctypes

import numpy as np

from . import _lib


def _i32(v):
    return np.ascontiguousarray(np.asarray(v).astype(np.int32, casting="unsafe")).reshape(-1)


def plan(rows, cols):
    """Partition plan of ``permanent(A, rows, cols)`` as a dict (fields of
    ``pq_plan_info``).  Raises like ``permanent`` for a sum mismatch."""
    lib = _lib.load()
    r, c = _i32(rows), _i32(cols)
    info = _lib.PlanInfo()
    _lib.check(lib.pq_perm_plan(len(r), len(c), r.ctypes.data_as(_lib.c_int32_p),
                                c.ctypes.data_as(_lib.c_int32_p), ctypes.byref(info)))
    return {name: getattr(info, name) for name, _ in info._fields_}


def batch_plan(rows, cols, nprob=1):
    """Plan of ONE problem of a ``permanent_batch`` call of ``nprob`` problems
    (``pq_perm_batch_plan``): ``kernel`` 3 = term-by-term batched walk, 4 = hypercube
    flavour; ``cols_padded`` = columns after the column multiplicities were written out."""
    lib = _lib.load()
    r, c = _i32(rows), _i32(cols)
    info = _lib.PlanInfo()
    _lib.check(lib.pq_perm_batch_plan(len(r), len(c), r.ctypes.data_as(_lib.c_int32_p),
                                      c.ctypes.data_as(_lib.c_int32_p), int(nprob),
                                      ctypes.byref(info)))
    return {name: getattr(info, name) for name, _ in info._fields_}


def gray_of_offset(rows, offset):
    """Gray digits (reference digit order, one per row) the GPU path assigns to
    ``offset``: enumeration parity with src/n_aryGrayCodeCounter.hpp:170-194."""
    lib = _lib.load()
    r = _i32(rows)
    gray = np.zeros(len(r), dtype=np.int32)
    _lib.check(lib.pq_perm_gray_of_offset(len(r), r.ctypes.data_as(_lib.c_int32_p),
                                          int(offset), gray.ctypes.data_as(_lib.c_int32_p)))
    return gray


def segment_range(nseg, part, nparts):
    """Segments [begin, end) owned by rank ``part`` of ``nparts``: the
    contiguous split ``nseg*part//nparts`` used by pq_perm_partial_c128 (the
    hierarchical form of src/permanent.cpp:158-164)."""
    if not (0 <= part < nparts):
        raise ValueError("part must be in [0, nparts)")
    return (nseg * part) // nparts, (nseg * (part + 1)) // nparts
