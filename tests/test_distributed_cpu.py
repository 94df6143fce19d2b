"""The N>1 host path on CPU: two ranks over gloo.

The GPU partial of each rank is stood in for by the oracle's long-double walk
of exactly the segment range the library assigns to that rank
(``plan.plan`` + ``plan.segment_range``); everything else -- argument
resolution, rank -> range mapping, the one all-gather of four doubles per rank and their error-free sum, the
final scaling -- is the product code path of ``piquasso_b200.distributed``.
"""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from conftest import haar


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_partial(a, r, c, part, nparts, device_index):
    from piquasso_b200 import plan
    p = plan.plan(r, c)
    if p["trivial"]:
        return complex(1.0, 0.0)
    b, e = plan.segment_range(p["nseg"], part, nparts)
    _, quads, _ = oracle.partial(a, r, c, b * p["seg_len"], e * p["seg_len"])
    return torch.tensor(quads[0], dtype=torch.float64)


def _worker(rank, world, port, cases, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from piquasso_b200 import distributed
    distributed._device_partial = _oracle_partial
    out = []
    for a, rows, cols in cases:
        out.append(complex(distributed.permanent_allgather(a, rows, cols, device_index=0)))
    results[rank] = out
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_gloo_sum_to_the_permanent():
    rng = np.random.default_rng(41)
    cases = [(haar(11, 11), np.ones(11, np.int32), np.ones(11, np.int32)),
             (haar(3, 2), np.zeros(3, np.int32), np.zeros(3, np.int32))]
    for trial in range(4):
        d = int(rng.integers(2, 7))
        nph = int(rng.integers(2, 9))
        cases.append((haar(d, 50 + trial),
                      rng.multinomial(nph, np.ones(d) / d).astype(np.int32),
                      rng.multinomial(nph, np.ones(d) / d).astype(np.int32)))
    world = 2
    manager = mp.get_context("spawn").Manager()
    results = manager.dict()
    mp.spawn(_worker, args=(world, _free_port(), cases, results), nprocs=world, join=True)
    assert set(results.keys()) == {0, 1}
    for i, (a, rows, cols) in enumerate(cases):
        want = oracle.permanent(a, rows, cols, precision=1)
        for rank in range(world):
            got = results[rank][i]
            assert abs(got - want) <= 1e-12 * max(abs(want), 1e-3), (i, rank, got, want)
        assert results[0][i] == results[1][i]  # every rank holds the same value


def _oracle_sampler_pmf(interferometer, out_occ, in_occ):
    """Stand-in for pq_sampler_pmf_c128 built on the oracle (CPU ranks)."""
    u = np.asarray(interferometer)
    d = u.shape[0]
    out = np.zeros((len(out_occ), d))
    for s, (oo, io) in enumerate(zip(np.asarray(out_occ), np.asarray(in_occ))):
        inz, onz = io > 0, oo > 0
        part = oracle.permanent_laplace(u[np.ix_(onz, inz)], oo[onz], io[inz])
        idx = np.arange(d)[inz]
        for m in range(d):
            amp = sum(io[idx[j]] * part[j] * u[m, idx[j]] for j in range(len(part)))
            out[s, m] = abs(amp) ** 2
    return out


def _sampler_worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from piquasso_b200 import distributed
    u = haar(6, 3)
    results[rank] = distributed.generate_samples_sharded([1, 1, 0, 2, 0, 0], 7, u, 11,
                                                         pmf_rows=_oracle_sampler_pmf)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sampler_shots_sharded_over_two_ranks():
    """Shots are sharded with no data-path collective; both ranks end with the
    single-process sample list (same per-shot seeds)."""
    from piquasso_b200 import sampling
    world = 2
    manager = mp.get_context("spawn").Manager()
    results = manager.dict()
    mp.spawn(_sampler_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    want = sampling.generate_samples([1, 1, 0, 2, 0, 0], 7, haar(6, 3), 11,
                                     pmf_rows=_oracle_sampler_pmf)
    assert results[0] == want and results[1] == want
    assert len(want) == 7 and all(sum(s) == 4 for s in want)


def _oracle_laplace_partial(a, r, c, part, nparts):
    """Stand-in for pq_perm_laplace_partial_c128 on CPU ranks: the oracle's walk of
    this rank's contiguous share of the term space, scaled like the library's."""
    if a.shape[0] == 0 or a.shape[1] == 0 or r.sum() == 0 or c.sum() == 0:
        return np.array([1.0 + 0j if part == 0 else 0j])
    _, _, idx_max = oracle.partial(a, r, c, 0, 0, laplace=True)
    b, e = (idx_max * part) // nparts, (idx_max * (part + 1)) // nparts
    vals, _, _ = oracle.partial(a, r, c, b, e, laplace=True)
    return vals / 2.0 ** (int(r.sum()) - 1)


def _laplace_worker(rank, world, port, cases, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from piquasso_b200 import distributed
    distributed._laplace_device_partial = _oracle_laplace_partial
    results[rank] = [distributed.permanent_laplace_allgather(a, rows, cols) for a, rows, cols in cases]
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_gloo_sum_to_one_permanent_laplace():
    """One Laplace problem, term space split over two ranks, one all-gather of C
    complex numbers and a fixed-order sum: both ranks hold permanent_laplace."""
    rng = np.random.default_rng(9)
    cases = []
    for k, rows in ((9, np.ones(8, np.int32)), (7, np.array([2, 1, 3], np.int32)),
                    (5, np.array([1, 0, 2, 1], np.int32))):
        a = (rng.normal(size=(len(rows), k)) + 1j * rng.normal(size=(len(rows), k))) / 2
        cases.append((a, rows, np.ones(k, np.int32)))
    cases.append((haar(3, 1), np.zeros(3, np.int32), np.ones(3, np.int32)))  # early-out [1]
    world = 2
    manager = mp.get_context("spawn").Manager()
    results = manager.dict()
    mp.spawn(_laplace_worker, args=(world, _free_port(), cases, results), nprocs=world, join=True)
    for i, (a, rows, cols) in enumerate(cases):
        want = oracle.permanent_laplace(a, rows, cols, precision=1)
        for rank in range(world):
            assert results[rank][i].shape == want.shape
            assert np.allclose(results[rank][i], want, rtol=1e-12, atol=1e-14), (i, rank)
        assert np.array_equal(results[0][i], results[1][i])
