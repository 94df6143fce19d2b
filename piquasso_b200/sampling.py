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

import ctypes

import numpy as np

from . import _lib
from .shot_rng import ShotStreams

__all__ = ["permanent_laplace_batch", "permanent_batch", "detection_probabilities",
           "grad_perm", "sampler_pmf", "sampler_draw", "generate_samples",
           "generate_lossy_samples"]

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


def permanent_batch(matrix, rows_batch, cols_batch):
    """``[permanent(matrix, r, c) for r, c in zip(rows_batch, cols_batch)]`` in one
    call (``pq_perm_batch_c128``): one matrix, many multiplicity vectors.
    ``rows_batch`` is (B, R), ``cols_batch`` (B, C) or a single (C,) vector that
    is broadcast.  Returns a complex128 array of length B."""
    lib = _lib.load()
    a = np.ascontiguousarray(matrix, dtype=np.complex128)
    if a.ndim != 2:
        raise ValueError("matrix must be 2-dimensional")
    R, C = a.shape
    # (no copy when the caller already holds contiguous int32 arrays)
    rb = np.ascontiguousarray(np.asarray(rows_batch), dtype=np.int32).reshape(-1, R)
    cb = np.asarray(np.asarray(cols_batch), dtype=np.int32)
    if cb.ndim == 1:
        cb = np.broadcast_to(cb, (rb.shape[0], C))
    cb = np.ascontiguousarray(cb).reshape(-1, C)
    if cb.shape[0] != rb.shape[0]:
        raise ValueError("rows_batch and cols_batch describe different numbers of problems")
    out = np.zeros(rb.shape[0], dtype=np.complex128)
    rc = lib.pq_perm_batch_c128(
        a.ctypes.data_as(_lib.c_double_p), R, C, rb.shape[0],
        rb.ctypes.data_as(_lib.c_int32_p), cb.ctypes.data_as(_lib.c_int32_p),
        out.ctypes.data_as(_lib.c_double_p))
    if rc in (_lib.PQ_ERR_BAD_ARG, _lib.PQ_ERR_TOO_LARGE):
        raise ValueError(_lib.last_error())
    _lib.check(rc)
    return out


def detection_probabilities(interferometer, input, outputs):
    """Particle-detection probabilities of many output occupations for one input:
    ``|permanent(U, rows=output, cols=input)|^2 / (prod output! * prod input!)``
    (``piquasso/_simulators/passive/probabilities.py:26-54`` and
    ``utils.py:131-138`` of the reference), all outputs in one batched call."""
    outputs = np.atleast_2d(np.asarray(outputs, dtype=int))
    input = np.asarray(input, dtype=int)
    amps = permanent_batch(interferometer, outputs, input)
    norm = np.prod(_factorials(outputs), axis=1) * np.prod(_factorials(input))
    return np.abs(amps) ** 2 / norm


_FACTORIAL_TABLE = None


def _factorials(occupations):
    """``scipy.special.factorial`` of an occupation array through a table of its own
    values (0! .. 170!, the same doubles): the element-wise scipy call was half the wall
    time of a 2000-output probability table (3 ms of 6.5)."""
    from scipy.special import factorial

    global _FACTORIAL_TABLE
    if _FACTORIAL_TABLE is None:
        _FACTORIAL_TABLE = factorial(np.arange(171))
    occupations = np.asarray(occupations)
    if occupations.size and (occupations.min() < 0 or occupations.max() > 170):
        return factorial(occupations)
    return _FACTORIAL_TABLE[occupations]


def grad_perm(matrix, rows, cols):
    """Gradient of the permanent with respect to every matrix element, what the
    reference's ``grad_perm`` (``src/permanent.cpp:271-300``, used by its JAX
    VJP) computes with one ``permanent_cpp`` call per element:

        grad[i, j] = rows[i] * cols[j] * perm(A, rows - e_i, cols - e_j)

    Here row i is ONE Laplace problem -- ``permanent_laplace(A, rows - e_i,
    cols)[j]`` is exactly ``perm(A, rows - e_i, cols - e_j)`` -- so the whole
    gradient is a single batched call of at most ``len(rows)`` problems.  Entries
    with ``rows[i] == 0`` or ``cols[j] == 0`` are zero, as in the reference."""
    a = np.ascontiguousarray(matrix, dtype=np.complex128)
    r = np.asarray(rows).astype(np.int64)
    c = np.asarray(cols).astype(np.int64)
    if a.shape != (len(r), len(c)):
        raise ValueError("multiplicities do not match the matrix")
    grad = np.zeros(a.shape, dtype=np.complex128)
    live = [i for i in range(len(r)) if r[i] > 0]
    if not live or c.sum() == 0:
        return grad
    rws = []
    for i in live:
        ri = r.copy()
        ri[i] -= 1
        rws.append(ri)
    parts = permanent_laplace_batch([a] * len(live), rws, [c] * len(live))
    for i, part in zip(live, parts):
        if len(part) == len(c):
            grad[i, :] = r[i] * c * part
        else:
            # early-out [1] (all remaining rows empty): perm of the empty minor is 1
            # for the single photon's column
            grad[i, :] = r[i] * c * (c > 0) * part[0] * (c.sum() == 1)
    grad[:, c == 0] = 0.0
    return grad


def _to_first_quantized(occupation):
    # piquasso/_math/indices.py:105-115
    out = []
    for mode, count in enumerate(occupation):
        out.extend([mode] * int(count))
    return np.array(out, dtype=int)


def sampler_pmf(interferometer, out_occ, in_occ, device=None):
    """Unnormalised pmf rows of one photon step for many shots at once
    (``pq_sampler_pmf_c128``): row s is ``_calculate_pmf(in_occ[s], out_occ[s],
    permanent_laplace, interferometer)`` of the reference before normalisation
    (``piquasso/_simulators/passive/sampling.py:723-749``).  ``device`` selects a
    CUDA device explicitly (``pq_sampler_pmf_dev_c128``)."""
    lib = _lib.load()
    U = np.ascontiguousarray(interferometer, dtype=np.complex128)
    d = U.shape[0]
    if U.shape != (d, d):
        raise ValueError("interferometer must be square")
    oo = np.ascontiguousarray(out_occ, dtype=np.int32).reshape(-1, d)
    io = np.ascontiguousarray(in_occ, dtype=np.int32).reshape(-1, d)
    if oo.shape != io.shape:
        raise ValueError("out_occ and in_occ must have the same shape")
    pmf = np.empty(oo.shape, dtype=np.float64)
    args = (U.ctypes.data_as(_lib.c_double_p), d, oo.shape[0],
            oo.ctypes.data_as(_lib.c_int32_p), io.ctypes.data_as(_lib.c_int32_p),
            pmf.ctypes.data_as(_lib.c_double_p))
    if device is None:
        rc = lib.pq_sampler_pmf_c128(*args)
    else:
        rc = lib.pq_sampler_pmf_dev_c128(int(device), *args)
    if rc in (_lib.PQ_ERR_BAD_ARG, _lib.PQ_ERR_TOO_LARGE):
        raise ValueError(_lib.last_error())
    _lib.check(rc)
    TIMERS["  of which GPU kernels (CUDA events)"] = (
        TIMERS.get("  of which GPU kernels (CUDA events)", 0.0)
        + max(lib.pq_last_kernel_ms(0 if device is None else int(device)), 0.0) * 1e-3)
    return pmf


def sampler_draw(interferometer, out_occ, in_occ, uniforms, device=None):
    """One photon step including the draw (``pq_sampler_draw_c128``): entry s is
    ``rng.choice(arange(d), p=_calculate_pmf(in_occ[s], out_occ[s], ...))`` of the
    reference (sampling.py:723-753) for the generator state in which
    ``rng.random()`` returns ``uniforms[s]``.  The pmf rows never leave the device:
    normalisation and numpy's cdf search are repeated there operation by
    operation.  ``device`` selects a CUDA device explicitly
    (``pq_sampler_draw_dev_c128``; calls for different devices from different
    threads run concurrently), ``None`` the library's first device."""
    lib = _lib.load()
    U = np.ascontiguousarray(interferometer, dtype=np.complex128)
    d = U.shape[0]
    if U.shape != (d, d):
        raise ValueError("interferometer must be square")
    oo = np.ascontiguousarray(out_occ, dtype=np.int32).reshape(-1, d)
    io = np.ascontiguousarray(in_occ, dtype=np.int32).reshape(-1, d)
    u = np.ascontiguousarray(uniforms, dtype=np.float64).reshape(-1)
    if oo.shape != io.shape or u.shape[0] != oo.shape[0]:
        raise ValueError("out_occ, in_occ and uniforms describe different numbers of shots")
    index = np.empty(oo.shape[0], dtype=np.int32)
    args = (U.ctypes.data_as(_lib.c_double_p), d, oo.shape[0],
            oo.ctypes.data_as(_lib.c_int32_p), io.ctypes.data_as(_lib.c_int32_p),
            u.ctypes.data_as(_lib.c_double_p), index.ctypes.data_as(_lib.c_int32_p))
    if device is None:
        rc = lib.pq_sampler_draw_c128(*args)
    else:
        rc = lib.pq_sampler_draw_dev_c128(int(device), *args)
    if rc in (_lib.PQ_ERR_BAD_ARG, _lib.PQ_ERR_TOO_LARGE):
        raise ValueError(_lib.last_error())
    _lib.check(rc)
    if (index < 0).any():
        raise ValueError("probabilities contain NaN")  # numpy's message for such a row
    prof = (ctypes.c_double * 4)()
    lib.pq_last_sampler_profile(prof)
    for name, ms in zip(("  of which planning (host threads)", "  of which waiting for the device",
                         "  of which device phase", "  of which GPU kernels (CUDA events)"), prof):
        TIMERS[name] = TIMERS.get(name, 0.0) + ms * 1e-3
    detail = (ctypes.c_double * 8)()
    lib.pq_last_sampler_detail(detail)
    for name, ms in zip(("    device phase: scratch growth", "    device phase: staging descriptors",
                         "    device phase: enqueue", "    device phase: waiting for the stream",
                         "    device phase: scatter", "    device phase: interferometer upload"),
                        detail):
        TIMERS[name] = TIMERS.get(name, 0.0) + ms * 1e-3
    return index


def _generate_samples_by_coroutines(input, shots, interferometer, seed_sequence,
                                    reject_condition, postselect_data,
                                    uniform_particle_overlap, pmf_rows):
    """The sampler variants (post-selection, uniform particle overlap) through
    the lock-step shot engine of :mod:`piquasso_b200.shot_engine`."""
    from . import shot_engine

    input = np.asarray(input, dtype=int)
    U = np.ascontiguousarray(interferometer, dtype=np.complex128)
    d = len(input)
    n = int(np.sum(input))
    first_quantized = _to_first_quantized(input)
    postselected = postselect_data is not None and len(postselect_data[0]) > 0

    def coroutine(idx, reject):
        rng = np.random.default_rng(seed=seed_sequence + idx)
        return shot_engine.shot_coroutine(d, n, first_quantized, rng, reject, U,
                                          postselect_data, uniform_particle_overlap)

    if reject_condition is not None and postselected:
        # The number of reject_condition() calls of a post-selected shot depends
        # on its retries, and the reference's condition draws from one generator
        # shared by all shots (simulation_steps.py:350-360): only the reference's
        # own shot-after-shot order reproduces it.
        results = []
        for idx in range(shots):
            results.extend(shot_engine.run_shots([coroutine(idx, reject_condition)], U,
                                                 pmf_rows))
    else:
        if reject_condition is None:
            rejects = [lambda: False] * shots
        else:
            # exactly n calls per shot, shot after shot: evaluate them up front
            drawn = [[bool(reject_condition()) for _ in range(n)] for _ in range(shots)]
            rejects = [iter(row).__next__ for row in drawn]
        results = shot_engine.run_shots([coroutine(idx, rejects[idx])
                                         for idx in range(shots)], U, pmf_rows)
    return [tuple(int(x) for x in sample) for sample in results]


def generate_lossy_samples(input, shots, interferometer, seed_sequence, postselect_data=None,
                           pmf_rows=None, devices=None):
    """Non-uniform losses (``generate_lossy_samples``, sampling.py:110-146 of the
    reference): sample the 2d-mode unitary dilation of the lossy transfer matrix
    and keep the first d modes."""
    from . import shot_engine

    input = np.asarray(input, dtype=int)
    expanded = shot_engine.expanded_interferometer(np.asarray(interferometer))
    expanded_input = np.concatenate([input, np.zeros_like(input)])
    samples = generate_samples(expanded_input, shots, expanded, seed_sequence,
                               reject_condition=None, postselect_data=postselect_data,
                               pmf_rows=pmf_rows, devices=devices)
    return [s[: len(input)] for s in samples]


def generate_samples(input, shots, interferometer, seed_sequence, reject_condition=None,
                     batch_shots=None, postselect_data=None, uniform_particle_overlap=None,
                     pmf_rows=None, overlap=None, devices=None, as_array=False):
    """Clifford & Clifford algorithm B, all shots in lock step.

    Restates ``_generate_samples`` / ``_generate_sample`` / ``_calculate_pmf``
    (``piquasso/_simulators/passive/sampling.py:149-236, 723-753``) with the
    loops interchanged: outer loop over the n photons, inner (batched) loop over
    shots.  Per photon step ONE library call (``pq_sampler_draw_c128``) filters
    the zeros, walks all shots' Laplace problems on the GPU, assembles the pmf
    rows on the device and draws from them there.

    Host RNG order is the reference's: shot ``idx`` owns
    ``np.random.default_rng(seed_sequence + idx)`` and per photon draws
    ``choice(len(to_shrink))`` and then ``choice(arange(d), p=pmf)``.  The two
    draws are issued here as ``integers(0, len)`` and ``random()`` +
    ``cdf.searchsorted(u, side="right")``, which is what ``Generator.choice``
    does internally and consumes the bit stream identically
    (``tests/test_host.py::test_numpy_choice_equivalences`` pins that), and both
    are derived for all shots at once from the shots' raw PCG64 streams
    (:mod:`piquasso_b200.shot_rng`, pinned against real generators by
    ``test_shot_streams_replay_numpy_generators``), so the returned tuples are
    identical to the reference's for the same seed.

    ``reject_condition`` (uniform losses, ``simulation_steps.py:350-360``) is a
    state-independent callable the reference evaluates once per photon per shot
    in shot-major order, possibly drawing from a shared generator; it is
    therefore evaluated up front in that same order.

    ``postselect_data = (modes, photons, max_trials)`` and
    ``uniform_particle_overlap`` select the reference's other per-shot
    algorithms (sampling.py:73-97); those run through the coroutine engine of
    :mod:`piquasso_b200.shot_engine`, still one batched GPU call per round.
    ``pmf_rows`` replaces the GPU call (tests inject the oracle there to exercise
    the host logic without a GPU); without it the plain sampler also draws on the
    device (:func:`sampler_draw`) and only the chosen modes come back.
    ``overlap`` > 1 runs shot batches in that many worker threads (host
    bookkeeping and planning of one batch under the GPU time of another; the
    library serialises the device phases).  The default is two equal halves from
    2000 shots on (config 4 on B200: 10^4 shots 1.52 -> 1.47 s, 5000 shots 0.735 -> 0.69 s,
    2500 shots 0.37 -> 0.35 s; the GPU time sits in the last three photons, so the gain is
    a few percent) and a single batch below (1250 shots gain 6 % on an idle host, but that
    is a rank's share at 8 GPUs, where eight processes already share the host's cores);
    ``batch_shots`` fixes the batch size instead.  ``devices`` (CUDA device indices) shards the shots over several
    GPUs inside this process, one host thread per device.  The result does not
    depend on any of them.  ``as_array`` returns the samples as one (shots, d) int32
    array instead of the reference's list of tuples (the sharded driver gathers
    arrays, not pickled tuples).
    """
    import time

    if ((postselect_data is not None and len(postselect_data[0]) > 0)
            or uniform_particle_overlap is not None):
        if pmf_rows is None:
            # the coroutine engine issues one batched call per round: it runs on the
            # first of `devices` (shots at different photon numbers share the call)
            if devices is not None and len(devices) >= 1:
                first = int(devices[0])
                pmf_rows = lambda u, o, i: sampler_pmf(u, o, i, device=first)  # noqa: E731
            else:
                pmf_rows = sampler_pmf
        out = _generate_samples_by_coroutines(input, shots, interferometer, seed_sequence,
                                              reject_condition, postselect_data,
                                              uniform_particle_overlap, pmf_rows)
        return np.array(out, dtype=np.int32).reshape(len(out), -1) if as_array else out
    input = np.asarray(input, dtype=int)
    U = np.ascontiguousarray(interferometer, dtype=np.complex128)
    d = len(input)
    n = int(np.sum(input))
    first_quantized = _to_first_quantized(input)
    if reject_condition is None:
        rejected = np.zeros((shots, n), dtype=bool)
    else:
        rejected = np.array([[bool(reject_condition()) for _ in range(n)]
                             for _ in range(shots)], dtype=bool).reshape(shots, n)
    cols = np.arange(max(n, 1))

    def run_batch(start, stop, device=None):
        nb = stop - start
        t0 = time.perf_counter()
        # shot idx owns default_rng(seed_sequence + idx); shot_rng replays numpy's
        # integers() / random() on the raw streams of all shots at once
        streams = ShotStreams(seed_sequence, start, stop, draws_per_shot=2 * n + 2)
        _tick("host: per-shot generators", t0)
        sample = np.zeros((nb, d), dtype=np.int32)
        current_input = np.zeros((nb, d), dtype=np.int32)
        # to_shrink of every shot as rows of one array, `remaining` entries valid
        shrink = np.tile(first_quantized, (nb, 1)) if n else np.zeros((nb, 0), dtype=int)
        remaining = np.full(nb, n, dtype=np.int64)
        for photon in range(n):
            t0 = time.perf_counter()
            live = np.flatnonzero(~rejected[start:stop, photon])
            if live.size == 0:
                continue
            # _grow_current_input (sampling.py:197-205): rng.choice(len(to_shrink)),
            # take that mode, np.delete it (later entries shift left)
            ridx = streams.integers(live, remaining[live])
            modes = shrink[live, ridx]
            current_input[live, modes] += 1
            if n > 1:
                if live.size == nb:
                    # no rejected shot: plain slices instead of two gathered copies
                    np.copyto(shrink[:, : n - 1], shrink[:, 1:],
                              where=cols[None, : n - 1] >= ridx[:, None])
                else:
                    shifted = np.where(cols[None, : n - 1] >= ridx[:, None],
                                       shrink[live, 1:], shrink[live, : n - 1])
                    shrink[live, : n - 1] = shifted
            remaining[live] -= 1
            _tick("host: grow input", t0)
            t0 = time.perf_counter()
            # _sample_from_pmf: Generator.choice(a, p=p) = cdf.searchsorted(random(),
            # side="right"); the variate does not depend on the pmf, so it is drawn
            # first and the search can happen next to the pmf, on the device
            u = streams.random(live)
            _tick("host: rng.random", t0)
            t0 = time.perf_counter()
            everyone = live.size == nb  # no rejected shot: hand the arrays over as they are
            if pmf_rows is None:
                index = sampler_draw(U, sample if everyone else sample[live],
                                     current_input if everyone else current_input[live], u,
                                     device=device)
                _tick("pq_sampler_draw_c128 (filter + plan + GPU walk + pmf + draw)", t0)
            else:
                pmf = pmf_rows(U, sample[live], current_input[live])
                _tick("pmf_rows", t0)
                t0 = time.perf_counter()
                # _calculate_pmf normalisation (sequential sum), then numpy's choice
                p = pmf / np.cumsum(pmf, axis=1)[:, -1:]
                cdf = np.cumsum(p, axis=1)
                cdf /= cdf[:, -1:]
                index = (cdf <= u[:, None]).sum(axis=1)
                _tick("host: normalise + search", t0)
            sample[live, index] += 1
        return sample

    def finish(parts):
        arr = np.concatenate(parts, axis=0) if parts else np.zeros((0, d), dtype=np.int32)
        return arr if as_array else [tuple(row) for row in arr.tolist()]

    if devices is not None and len(devices) > 1 and pmf_rows is None and shots >= len(devices):
        # one process, several GPUs: equal contiguous shot ranges, one host thread per
        # device (the library serialises per device, not globally); no exchange step
        from concurrent.futures import ThreadPoolExecutor
        g = len(devices)
        jobs = [((shots * i) // g, (shots * (i + 1)) // g, int(dev))
                for i, dev in enumerate(devices)]
        with ThreadPoolExecutor(max_workers=g) as pool:
            parts = list(pool.map(lambda job: run_batch(*job), jobs))
        return finish(parts)
    # fewer shots than devices, or a single device: everything on the first one
    device = int(devices[0]) if devices is not None and len(devices) >= 1 else None
    if overlap is None:
        overlap = 2 if (pmf_rows is None and batch_shots is None
                        and shots >= _OVERLAP_DEFAULT_SHOTS) else 1
    # Shots are independent, so batches may also run concurrently on ONE device:
    # while one batch waits for the GPU inside the library call (GIL released,
    # device phases serialised by the library), another does its host bookkeeping
    # and planning.  Unequal batch sizes keep the threads out of step -- one in
    # its host-bound early photons while the other is in its GPU-bound late ones.
    bounds = _batch_bounds(shots, batch_shots, overlap if pmf_rows is None else 1)
    if len(bounds) <= 1 or overlap <= 1 or pmf_rows is not None:
        parts = [run_batch(b, e, device) for b, e in bounds]
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=overlap) as pool:
            parts = list(pool.map(lambda be: run_batch(be[0], be[1], device), bounds))
    return finish(parts)


# shots below which one batch is not worth splitting for host/GPU overlap
_OVERLAP_MIN_SHOTS = 2000
# shots from which two overlapping halves are the default
_OVERLAP_DEFAULT_SHOTS = 2000
# batch sizes (fractions of the shots) handed to more than two worker threads, in this order
_OVERLAP_FRACTIONS = (0.375, 0.25, 0.25, 0.125)


def _batch_bounds(shots, batch_shots, overlap):
    """[(begin, end)] of the shot batches of generate_samples."""
    if batch_shots is not None:
        step = max(1, int(batch_shots))
        return [(b, min(shots, b + step)) for b in range(0, shots, step)]
    if overlap <= 1 or shots < _OVERLAP_MIN_SHOTS:
        return [(0, shots)] if shots > 0 else []
    if overlap == 2 and len(_OVERLAP_FRACTIONS) == 4:
        return [(0, shots // 2), (shots // 2, shots)]
    cuts = np.rint(np.cumsum((0.0,) + _OVERLAP_FRACTIONS) * shots).astype(int)
    cuts[-1] = shots
    return [(int(b), int(e)) for b, e in zip(cuts[:-1], cuts[1:]) if e > b]

