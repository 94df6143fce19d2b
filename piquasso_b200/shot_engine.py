"""Lock-step shot engine for the sampler variants around ``permanent_laplace``
(SURVEY.md section 8 f-4): post-selected, partially distinguishable (uniform
overlap) and non-uniformly lossy Clifford-Clifford sampling.

The reference runs these one shot at a time
(``piquasso/_simulators/passive/sampling.py:110-146, 239-529``), every photon
step being one ``permanent_laplace`` call.  Here each shot is a *coroutine*: it
does its own host-side bookkeeping and RNG draws in the reference's order and
``yield``s an ``(input occupation, output occupation)`` pair whenever it needs a
pmf row.  The engine collects the requests of all live shots -- which may be
at different photon numbers, on different retries -- and answers them with ONE
batched ``pq_sampler_pmf_c128`` call per round.  Each shot owns
``np.random.default_rng(seed_sequence + idx)``, so the samples are identical
to the reference's for the same seed.

The plain (no post-selection, indistinguishable) sampler keeps its vectorised
implementation in :mod:`piquasso_b200.sampling`; its coroutine here is what the
variants delegate to.
"""

from __future__ import annotations

import numpy as np

__all__ = ["InvalidSimulation", "run_shots", "shot_coroutine", "expanded_interferometer"]


class InvalidSimulation(Exception):
    """Same role as ``piquasso.api.exceptions.InvalidSimulation``: the
    post-selection criteria were not met within the allowed number of trials."""


# --------------------------------------------------------------------------
# host helpers shared by the shot coroutines

def _first_quantized(occupation):
    # piquasso/_math/indices.py:105-115
    return np.repeat(np.arange(len(occupation)), np.asarray(occupation, dtype=int))


def _second_quantized(first_quantized, d):
    # piquasso/_math/indices.py:118-124
    return np.bincount(np.asarray(first_quantized, dtype=int), minlength=d).astype(np.int64)


def _draw_input_photon(current_input, to_shrink, rng):
    """``_grow_current_input`` (sampling.py:197-205): one more input photon, picked
    uniformly among those not used yet."""
    pick = rng.choice(len(to_shrink))
    current_input[to_shrink[pick]] += 1
    return np.delete(to_shrink, pick)


def _times_linear_truncated(poly, constant, linear, out):
    """``(constant + sum_j linear[j] x_j) * poly`` truncated to poly's shape
    (``piquasso/_math/polynomial.py:19-40``).  As in the reference, ``out`` may
    alias ``poly``; the shifted terms are then read from the already updated
    array (sampling.py:547-550 calls it that way), which is reproduced here
    because the draws that follow depend on the value."""
    out[...] = constant * poly
    for axis, coefficient in enumerate(linear):
        np.moveaxis(out, axis, 0)[1:] += coefficient * np.moveaxis(poly, axis, 0)[:-1]
    return out


# --------------------------------------------------------------------------
# shot coroutines: ``pmf = yield (current_input, sample)``

def _indistinguishable(d, n, first_quantized_input, rng, reject):
    """``_generate_sample`` (sampling.py:208-236)."""
    sample = np.zeros(d, dtype=int)
    current_input = np.zeros(d, dtype=int)
    to_shrink = np.copy(first_quantized_input)
    for _ in range(int(n)):
        if reject():
            continue
        to_shrink = _draw_input_photon(current_input, to_shrink, rng)
        pmf = yield current_input, sample
        sample[rng.choice(np.arange(d), p=pmf)] += 1
    return sample


def _postselected(d, n, first_quantized_input, rng, reject, postselect_data, trim=True):
    """``_generate_sample_with_postselect`` (sampling.py:239-325): restart the shot
    as soon as a post-selected mode overflows or can no longer be filled.  With
    ``trim=False`` (the reference's ``track_photons_needed=False``) the second
    check is off and the post-selected modes stay in the returned sample."""
    modes, photons, max_trials = postselect_data
    modes = tuple(int(m) for m in modes)
    wanted_total = int(np.sum(photons))
    n = int(n)
    trials = 0
    while True:
        trials += 1
        if trials > max_trials:
            raise InvalidSimulation(
                "Too many trials during sample generation: the post-selection criteria "
                "may be very unlikely (max_sample_generation_trials = %s)." % (max_trials,))
        sample = np.zeros(d, dtype=int)
        current_input = np.zeros(d, dtype=int)
        to_shrink = np.copy(first_quantized_input)
        still_needed = wanted_total
        room = np.array(photons, dtype=int)
        failed = False
        for k in range(1, n + 1):
            if reject():
                continue
            to_shrink = _draw_input_photon(current_input, to_shrink, rng)
            pmf = yield current_input, sample
            index = int(rng.choice(np.arange(d), p=pmf))
            sample[index] += 1
            if index in modes:
                still_needed -= 1
                room[modes.index(index)] -= 1
                if np.any(room < 0):
                    failed = True
                    break
            if trim and still_needed > n - k:
                failed = True
                break
        if not failed:
            break
    if not trim:
        return sample
    return np.delete(sample, list(modes))


def _split_by_overlap(occupation, overlap, rng):
    """``_separate_particles`` (sampling.py:367-394): how many photons of each
    input mode take part in the interference (Renema et al., arXiv:1707.02793)."""
    from scipy.special import comb, factorial

    d = len(occupation)
    indist = np.zeros(d, dtype=int)
    dist = np.zeros(d, dtype=int)
    for j, n_j in enumerate(occupation):
        if n_j == 0:
            continue
        weights = np.array(
            [comb(n_j, k) * (overlap ** k) * ((1.0 - overlap) ** (n_j - k)) * factorial(k)
             for k in range(n_j + 1)], dtype=float)
        k_j = rng.choice(n_j + 1, p=weights / weights.sum())
        indist[j] = k_j
        dist[j] = n_j - k_j
    return indist, dist


def _classical_particles(interferometer, particles, rng, reject):
    """``_sample_distinguishable_particles`` (sampling.py:397-417)."""
    d = len(particles)
    output = np.zeros(d, dtype=int)
    for input_mode, count in enumerate(particles):
        probabilities = np.abs(interferometer[:, input_mode]) ** 2
        probabilities = probabilities / np.sum(probabilities)  # lossy: not unitary
        for _ in range(count):
            mode = rng.choice(d, p=probabilities)
            if not reject():
                output[mode] += 1
    return output


def _uniform_overlap(d, n, first_quantized_input, rng, reject, interferometer, overlap):
    """``_generate_sample_with_uniform_overlap`` (sampling.py:328-364)."""
    occupation = _second_quantized(first_quantized_input, d)
    indist, dist = _split_by_overlap(occupation, overlap, rng)
    dist_output = _classical_particles(interferometer, dist, rng, reject)
    indist_output = yield from _indistinguishable(
        d, np.sum(indist), _first_quantized(indist), rng, reject)
    return tuple(dist_output + indist_output)


def _dist_postselection_probability(interferometer, dist, modes, photons):
    # sampling.py:532-552
    poly = np.zeros(tuple(photons + 1), dtype=float)
    poly[(0,) * len(photons)] = 1.0
    for input_mode, multiplicity in enumerate(dist):
        probabilities = np.abs(interferometer[modes, input_mode]) ** 2
        for _ in range(multiplicity):
            _times_linear_truncated(poly, 1.0 - probabilities.sum(), probabilities, out=poly)
    return poly[tuple(photons)]


def _dist_postselection_table(interferometer, dist, modes, bound):
    # sampling.py:555-583
    dist_input = _first_quantized(dist)
    table = [np.zeros(tuple(bound + 1), dtype=float) for _ in range(len(dist_input) + 1)]
    table[-1][(0,) * len(bound)] = 1.0
    for photon in range(len(dist_input) - 1, -1, -1):
        probabilities = np.abs(interferometer[modes, dist_input[photon]]) ** 2
        _times_linear_truncated(table[photon + 1], 1.0 - probabilities.sum(), probabilities,
                                out=table[photon])
    return table


def _dist_output_given_postselection(interferometer, dist, modes, photons, table, rng):
    # sampling.py:586-657
    d = interferometer.shape[0]
    dist_input = _first_quantized(dist)
    free_modes = np.delete(np.arange(d), modes)
    sample = np.zeros(d, dtype=int)
    remaining = np.array(photons, dtype=int)
    for photon, input_mode in enumerate(dist_input):
        nxt = table[photon + 1]
        future = nxt[tuple(remaining)]
        p_free = np.abs(interferometer[free_modes, input_mode]) ** 2
        p_post = np.abs(interferometer[modes, input_mode]) ** 2
        p_loss = 1.0 - p_free.sum() - p_post.sum()
        w_post = np.zeros(len(modes), dtype=float)
        for axis, probability in enumerate(p_post):
            if remaining[axis] == 0:
                continue
            remaining[axis] -= 1
            w_post[axis] = probability * nxt[tuple(remaining)]
            remaining[axis] += 1
        weights = np.concatenate([p_free * future, np.array([p_loss * future], dtype=float),
                                  w_post])
        weights /= np.sum(weights)
        index = rng.choice(len(weights), p=weights)
        if index < len(free_modes):
            sample[free_modes[index]] += 1
        elif index > len(free_modes):
            axis = index - len(free_modes) - 1
            sample[modes[axis]] += 1
            remaining[axis] -= 1
    return sample


def _postselected_uniform_overlap(d, n, first_quantized_input, rng, reject, interferometer,
                                  postselect_data, overlap):
    """``_generate_sample_with_postselect_and_uniform_overlap``
    (sampling.py:420-529): greedy retry loop around the two halves."""
    modes, photons, max_trials = postselect_data
    modes = np.asarray(modes, dtype=int)
    photons = np.asarray(photons, dtype=int)
    occupation = _second_quantized(first_quantized_input, d)
    trials = 0
    while True:
        trials += 1
        if trials > max_trials:
            raise InvalidSimulation(
                "Too many trials during sample generation: the post-selection criteria "
                "may be very unlikely (max_sample_generation_trials = %s)." % (max_trials,))
        indist, dist = _split_by_overlap(occupation, overlap, rng)
        try:
            indist_output = yield from _postselected(
                d, np.sum(indist), _first_quantized(indist), rng, reject,
                (modes, photons, 1), trim=False)
        except InvalidSimulation:
            continue
        remaining = photons - indist_output[modes]
        if rng.random() > _dist_postselection_probability(interferometer, dist, modes,
                                                           remaining):
            continue
        table = _dist_postselection_table(interferometer, dist, modes, photons)
        output = _dist_output_given_postselection(interferometer, dist, modes, remaining,
                                                  table, rng)
        return tuple(np.delete(output + indist_output, modes))


def shot_coroutine(d, n, first_quantized_input, rng, reject, interferometer,
                   postselect_data=None, uniform_particle_overlap=None):
    """The reference's choice of per-shot algorithm (sampling.py:73-97)."""
    postselected = postselect_data is not None and len(postselect_data[0]) > 0
    if postselected and uniform_particle_overlap is None:
        return _postselected(d, n, first_quantized_input, rng, reject, postselect_data)
    if postselected:
        return _postselected_uniform_overlap(d, n, first_quantized_input, rng, reject,
                                             interferometer, postselect_data,
                                             uniform_particle_overlap)
    if uniform_particle_overlap is None:
        return _indistinguishable(d, n, first_quantized_input, rng, reject)
    return _uniform_overlap(d, n, first_quantized_input, rng, reject, interferometer,
                            uniform_particle_overlap)


# --------------------------------------------------------------------------
# the engine

def run_shots(coroutines, interferometer, pmf_rows):
    """Drive shot coroutines to completion, one batched pmf call per round.

    ``pmf_rows(U, out_occ, in_occ)`` returns the UNNORMALISED pmf rows
    (:func:`piquasso_b200.sampling.sampler_pmf`); rows are normalised here the
    way ``_calculate_pmf`` does (sequential sum, sampling.py:736-749)."""
    results = [None] * len(coroutines)
    waiting = {}
    for idx, co in enumerate(coroutines):
        try:
            waiting[idx] = next(co)
        except StopIteration as done:
            results[idx] = done.value
    while waiting:
        order = list(waiting)
        in_occ = np.stack([waiting[i][0] for i in order]).astype(np.int32)
        out_occ = np.stack([waiting[i][1] for i in order]).astype(np.int32)
        pmf = pmf_rows(interferometer, out_occ, in_occ)
        norm = np.cumsum(pmf, axis=1)[:, -1]
        for row, idx in enumerate(order):
            try:
                waiting[idx] = coroutines[idx].send(pmf[row] / norm[row])
            except StopIteration as done:
                results[idx] = done.value
                del waiting[idx]
    return results


def expanded_interferometer(interferometer):
    """2d x 2d dilation of a lossy d x d transfer matrix from its SVD (isometric
    on the first d input modes, where the photons are)
    (``_prepare_interferometer_matrix_in_expanded_space``, sampling.py:756-786):
    ``[[V,0],[0,1]] @ [[S, C],[C, S]] @ [[W,0],[0,1]]`` with ``C = sqrt(1 - S^2)``."""
    v, s, w = np.linalg.svd(interferometer)
    d = len(v)
    zeros = np.zeros_like(v)
    eye = np.eye(d)
    c = np.diag(np.sqrt(1.0 - np.power(s, 2)))
    middle = np.block([[np.diag(s), c], [c, np.diag(s)]])
    return (np.block([[v, zeros], [zeros, eye]]) @ middle
            @ np.block([[w, zeros], [zeros, eye]]))
