"""Batched front door of the Laplace path and the lock-step Clifford-Clifford
sampler that drives it.

``permanent_laplace_batch`` packs many independent ``permanent_laplace``
problems into ONE ``pq_perm_laplace_batch_c128`` call (one kernel launch per
kernel variant).  ``generate_samples`` restates the reference's sampler
(``piquasso/_simulators/passive/sampling.py:149-236, 711-753``) shot-parallel:
all shots advance one photon at a time, every shot keeping its own
``np.random.default_rng(seed_sequence + idx)`` and drawing from it in exactly
the reference's order, so the samples are identical to the reference's for the
same seed.
"""

from __future__ import annotations

import numpy as np

from . import _lib

__all__ = ["permanent_laplace_batch", "generate_samples"]

# wall-clock split of generate_samples (seconds), for tools/sampler_bench.py
TIMERS = {}


def _tick(name, t0):
    import time
    TIMERS[name] = TIMERS.get(name, 0.0) + (time.perf_counter() - t0)



def permanent_laplace_batch(matrices, rows_list, cols_list):
    """``[permanent_laplace(m, r, c) for m, r, c in zip(...)]`` in one call.

    All problems are computed in complex128.  Returns a list of 1-d complex128
    arrays (length ``len(c)``, or 1 on the reference's early-out)."""
    lib = _lib.load()
    n = len(matrices)
    if not (len(rows_list) == n and len(cols_list) == n):
        raise ValueError("matrices, rows_list and cols_list must have equal length")
    if n == 0:
        return []
    R = np.empty(n, dtype=np.int32)
    C = np.empty(n, dtype=np.int32)
    mats = []
    rws = []
    cls = []
    for b in range(n):
        a = np.ascontiguousarray(matrices[b], dtype=np.complex128)
        if a.ndim != 2:
            raise ValueError("problem %d: matrix must be 2-dimensional" % b)
        r = np.ascontiguousarray(np.asarray(rows_list[b]).astype(np.int32, casting="unsafe"))
        c = np.ascontiguousarray(np.asarray(cols_list[b]).astype(np.int32, casting="unsafe"))
        if r.shape != (a.shape[0],) or c.shape != (a.shape[1],):
            raise ValueError("problem %d: multiplicities do not match the matrix" % b)
        R[b], C[b] = a.shape
        mats.append(a.reshape(-1))
        rws.append(r)
        cls.append(c)
    a_sizes = R.astype(np.int64) * C.astype(np.int64)
    a_off = np.concatenate(([0], np.cumsum(a_sizes)[:-1])).astype(np.int64)
    r_off = np.concatenate(([0], np.cumsum(R.astype(np.int64))[:-1])).astype(np.int64)
    c_off = np.concatenate(([0], np.cumsum(C.astype(np.int64))[:-1])).astype(np.int64)
    widths = np.maximum(C.astype(np.int64), 1)
    o_off = np.concatenate(([0], np.cumsum(widths)[:-1])).astype(np.int64)
    A = np.concatenate(mats) if a_sizes.sum() else np.zeros(1, dtype=np.complex128)
    rows = np.concatenate(rws) if R.sum() else np.zeros(1, dtype=np.int32)
    cols = np.concatenate(cls) if C.sum() else np.zeros(1, dtype=np.int32)
    out = np.zeros(int(widths.sum()), dtype=np.complex128)
    out_len = np.zeros(n, dtype=np.int32)
    rc = lib.pq_perm_laplace_batch_c128(
        n, A.ctypes.data_as(_lib.c_double_p), a_off.ctypes.data_as(_lib.c_int64_p),
        R.ctypes.data_as(_lib.c_int32_p), C.ctypes.data_as(_lib.c_int32_p),
        rows.ctypes.data_as(_lib.c_int32_p), r_off.ctypes.data_as(_lib.c_int64_p),
        cols.ctypes.data_as(_lib.c_int32_p), c_off.ctypes.data_as(_lib.c_int64_p),
        out.ctypes.data_as(_lib.c_double_p), o_off.ctypes.data_as(_lib.c_int64_p),
        out_len.ctypes.data_as(_lib.c_int32_p))
    if rc in (_lib.PQ_ERR_BAD_ARG, _lib.PQ_ERR_TOO_LARGE):
        raise ValueError(_lib.last_error())
    _lib.check(rc)
    return [out[o_off[b]: o_off[b] + out_len[b]].copy() for b in range(n)]


def _to_first_quantized(occupation):
    # piquasso/_math/indices.py:105-115
    out = []
    for mode, count in enumerate(occupation):
        out.extend([mode] * int(count))
    return np.array(out, dtype=int)


def generate_samples(input, shots, interferometer, seed_sequence, reject_condition=None,
                     batch_shots=None):
    """Clifford & Clifford algorithm B, all shots in lock step.

    Restates ``_generate_samples`` / ``_generate_sample`` / ``_calculate_pmf``
    (``piquasso/_simulators/passive/sampling.py:149-236, 723-753``) with the
    loops interchanged: outer loop over the n photons, inner (batched) loop over
    shots.  Shot ``idx`` owns ``np.random.default_rng(seed_sequence + idx)`` and
    draws ``choice(len(to_shrink))`` then ``choice(arange(d), p=pmf)`` per photon
    exactly as the reference does, so the returned tuples are identical.

    ``reject_condition`` (uniform losses, ``simulation_steps.py:350-360``) is a
    state-independent callable the reference evaluates once per photon per shot
    in shot-major order, possibly drawing from a shared generator; it is
    therefore evaluated up front in that same order.
    """
    input = np.asarray(input, dtype=int)
    U = np.ascontiguousarray(interferometer, dtype=np.complex128)
    d = len(input)
    n = int(np.sum(input))
    first_quantized = _to_first_quantized(input)
    if batch_shots is None:
        batch_shots = shots
    if reject_condition is None:
        rejected = np.zeros((shots, n), dtype=bool)
    else:
        rejected = np.array([[bool(reject_condition()) for _ in range(n)]
                             for _ in range(shots)], dtype=bool).reshape(shots, n)
    samples_all = []
    for start in range(0, shots, max(1, batch_shots)):
        stop = min(shots, start + max(1, batch_shots))
        nb = stop - start
        rngs = [np.random.default_rng(seed=seed_sequence + idx) for idx in range(start, stop)]
        sample = np.zeros((nb, d), dtype=int)
        current_input = np.zeros((nb, d), dtype=int)
        to_shrink = [np.copy(first_quantized) for _ in range(nb)]
        arange_d = np.arange(d)
        import time
        for photon in range(n):
            t0 = time.perf_counter()
            mats, rws, cls, nz = [], [], [], []
            live = [s for s in range(nb) if not rejected[start + s, photon]]
            for s in live:
                # _grow_current_input (sampling.py:197-205)
                ridx = rngs[s].choice(len(to_shrink[s]))
                mode = to_shrink[s][ridx]
                current_input[s, mode] += 1
                to_shrink[s] = np.delete(to_shrink[s], ridx)
                # _filter_zeros (sampling.py:711-720)
                in_nz = current_input[s] > 0
                out_nz = sample[s] > 0
                mats.append(U[np.ix_(out_nz, in_nz)])
                rws.append(sample[s][out_nz])
                cls.append(current_input[s][in_nz])
                nz.append(arange_d[in_nz])
            _tick("host: grow input + filter", t0)
            t0 = time.perf_counter()
            partials = permanent_laplace_batch(mats, rws, cls)
            _tick("permanent_laplace_batch (pack + plan + GPU)", t0)
            t0 = time.perf_counter()
            for i, s in enumerate(live):
                # _calculate_pmf (sampling.py:736-749): pmf[m] = |sum_j in_j p_j U[m, nz_j]|^2
                weights = current_input[s][nz[i]] * partials[i]
                amp = U[:, nz[i]] @ weights
                pmf = np.abs(amp) ** 2
                pmf = pmf / pmf.sum()
                index = rngs[s].choice(arange_d, p=pmf)
                sample[s, index] += 1
            _tick("host: pmf + rng.choice", t0)
        samples_all.extend(tuple(int(x) for x in row) for row in sample)
    return samples_all
