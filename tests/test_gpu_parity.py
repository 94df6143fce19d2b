"""Parity of the CUDA path (through the C ABI) with the oracle, the
reference's golden vectors and size-independent properties.

Tolerances (north_star): complex128 permanents within RELATIVE 1e-10 of the
reference up to n=32 (asserted here at 1e-10 against the long-double arbiter and
the compiled-reference goldens); beyond n=32 the reference itself is wrong
(int offset truncation), so closed forms with relative 1e-9 are used.
complex64 inputs: relative 1e-5."""

import ctypes
import math

import numpy as np
import pytest

import oracle
from conftest import golden_complex, golden_matrix, haar, load_golden, relerr
from piquasso_b200 import _lib, plan
from piquasso_b200._math.permanent import permanent, permanent_laplace

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def close(got, want, rtol=RTOL, atol=1e-13):
    return abs(got - want) <= rtol * abs(want) + atol


@pytest.fixture(autouse=True)
def _reset_choices(lib):
    lib.pq_set_kernel_choice(0)
    lib.pq_set_seg_len_hint(0)
    yield
    lib.pq_set_kernel_choice(0)
    lib.pq_set_seg_len_hint(0)


def test_device_present(lib):
    assert lib.pq_device_count() >= 1


def test_reference_test_goldens():
    for case in load_golden("permanent_reference_tests.json"):
        m = golden_matrix(case["matrix"])
        want = golden_complex(case["value"])
        got = permanent(m, rows=case["rows"], cols=case["cols"])
        assert got.shape == ()
        if m.dtype == np.complex64:
            assert got.dtype == np.complex64
            assert close(complex(got), want, rtol=1e-5, atol=1e-6), case["source"]
        else:
            assert got.dtype == np.complex128
            assert close(complex(got), want), case["source"]


def test_reference_literal_goldens():
    assert np.isclose(permanent(np.array([[4.2]]), cols=np.ones(1, int), rows=np.ones(1, int)), 4.2)
    u = np.array([[1, 1j], [1, -1j]]) / np.sqrt(2)
    assert np.isclose(permanent(u, cols=np.array([0, 2]), rows=np.array([2, 0])), -1)
    assert np.isclose(permanent(np.array([[0j]]), [1], [1]), 0.0)
    assert np.isclose(permanent(np.array([[1, -2], [-3, 4]], dtype=complex), [1, 1], [1, 1]), 10.0)
    big = permanent(np.full((3, 3), 1e10, dtype=complex), [1, 1, 1], [1, 1, 1])
    assert np.isfinite(big) and np.isclose(complex(big), 6e30, rtol=1e-10)
    assert np.isclose(complex(permanent(np.array([[2, 3], [4, 5]], dtype=complex), [3, 0], [2, 1])), 72.0)
    assert np.isclose(permanent(np.eye(8, dtype=complex), np.ones(8, int), np.ones(8, int)), 1.0)


def test_haar_goldens_from_compiled_reference():
    for case in load_golden("permanent_haar.json"):
        m = golden_matrix(case["matrix"])
        want = golden_complex(case["value"])
        got = complex(permanent(m, case["rows"], case["cols"]))
        assert close(got, want), (case["source"], got, want)


def test_laplace_goldens_from_compiled_reference():
    for case in load_golden("laplace.json"):
        m = golden_matrix(case["matrix"])
        want = np.array([golden_complex(z) for z in case["value"]])
        got = permanent_laplace(m, rows=case["rows"], cols=case["cols"])
        assert got.shape == want.shape and got.dtype == np.complex128, case["source"]
        assert np.allclose(got, want, rtol=RTOL, atol=1e-13), case["source"]


@pytest.mark.parametrize("choice", [1, 2, 22, 32, 42])
def test_every_kernel_variant_matches_the_arbiter(lib, choice):
    for n in (9, 13, 17, 21):
        a = haar(n, n)
        ones = np.ones(n, np.int32)
        want = oracle.permanent(a, ones, ones, precision=1, njobs=64)
        lib.pq_set_kernel_choice(choice)
        assert relerr(complex(permanent(a, ones, ones)), want) < RTOL, (n, choice)


def test_random_multiplicities_against_the_arbiter(lib):
    rng = np.random.default_rng(17)
    for trial in range(60):
        d = int(rng.integers(1, 10))
        nph = int(rng.integers(1, 12))
        rows = rng.multinomial(nph, np.ones(d) / d)
        cols = rng.multinomial(nph, np.ones(d) / d)
        a = haar(d, 700 + trial)
        want = oracle.permanent(a, rows, cols, precision=1)
        for hint in (0, 1, 5, 4096):
            lib.pq_set_seg_len_hint(hint)
            assert close(complex(permanent(a, rows, cols)), want), (rows, cols, hint)


def test_rectangular_and_assym_reduce_property():
    """permanent(U, rows, cols) == permanent(assym_reduce(U, rows, cols), 1, 1)
    (tests/_math/test_permanent.py:231-249, 304-320 of the reference)."""
    rng = np.random.default_rng(23)
    for trial in range(20):
        d1, d2 = int(rng.integers(1, 7)), int(rng.integers(1, 7))
        nph = int(rng.integers(1, 9))
        rows = rng.multinomial(nph, np.ones(d1) / d1)
        cols = rng.multinomial(nph, np.ones(d2) / d2)
        a = rng.normal(size=(d1, d2)) + 1j * rng.normal(size=(d1, d2))
        big = oracle.assym_reduce(a, rows, cols)
        ones = np.ones(nph, int)
        lhs = complex(permanent(a, rows, cols))
        rhs = complex(permanent(np.ascontiguousarray(big), ones, ones))
        assert close(lhs, rhs, rtol=1e-9, atol=1e-10)
        assert close(lhs, oracle.permanent(a, rows, cols, precision=1), rtol=1e-9, atol=1e-10)


def test_partition_indexing_segment_sums(lib):
    """Segment s of the GPU partition is exactly the reference's offset range
    [s*W, (s+1)*W): per-segment sums against the long-double walk of the same
    offsets."""
    rng = np.random.default_rng(29)
    cases = [(haar(12, 3), np.ones(12, int), np.ones(12, int), 16, 1),
             (haar(12, 3), np.ones(12, int), np.ones(12, int), 16, 2),
             (haar(14, 4), np.ones(14, int), np.ones(14, int), 8, 22),
             (haar(15, 6), np.ones(15, int), np.ones(15, int), 64, 32),
             (haar(17, 7), np.ones(17, int), np.ones(17, int), 256, 42)]
    for trial in range(6):
        d = int(rng.integers(3, 8))
        nph = int(rng.integers(4, 11))
        rows = rng.multinomial(nph, np.ones(d) / d)
        cols = rng.multinomial(nph, np.ones(d) / d)
        cases.append((haar(d, 40 + trial), rows, cols, int(rng.integers(1, 40)), 0))
    for a, rows, cols, hint, choice in cases:
        lib.pq_set_kernel_choice(choice)
        lib.pq_set_seg_len_hint(hint)
        p = plan.plan(rows, cols)
        nseg, W = p["nseg"], p["seg_len"]
        take = int(min(nseg, 37))
        begin = int(max(0, nseg - take) // 2)
        ac = np.ascontiguousarray(a, dtype=np.complex128)
        r32 = np.ascontiguousarray(rows, dtype=np.int32)
        c32 = np.ascontiguousarray(cols, dtype=np.int32)
        out = np.zeros(2 * take)
        _lib.check(lib.pq_perm_segment_sums_c128(
            ac.ctypes.data_as(_lib.c_double_p), ac.shape[0], ac.shape[1],
            r32.ctypes.data_as(_lib.c_int32_p), c32.ctypes.data_as(_lib.c_int32_p),
            begin, take, out.ctypes.data_as(_lib.c_double_p)))
        got = out.view(np.complex128)
        for s in range(take):
            want, _, idx_max = oracle.partial(a, rows, cols, (begin + s) * W, (begin + s + 1) * W)
            assert idx_max == p["idx_max"]
            assert close(got[s], want[0], rtol=1e-11, atol=1e-12), (rows, cols, s)


def test_partials_of_all_ranks_sum_to_the_whole(lib):
    import torch
    from piquasso_b200.distributed import _device_partial, combine, finish
    a = np.ascontiguousarray(haar(22, 22), dtype=np.complex128)
    ones = np.ones(22, np.int32)
    whole = complex(permanent(a, ones, ones))
    for nparts in (1, 2, 3, 8):
        quads = []
        for g in range(nparts):
            part = _device_partial(a, ones, ones, g, nparts, 0)
            torch.cuda.synchronize()
            quads.append(part.cpu().numpy())
        assert relerr(finish(combine(quads), 22), whole) < 1e-13


def test_closed_forms_up_to_n32():
    """rank-1 unit-modulus matrices: perm = n! prod(u) prod(v); the reference's
    own error on these is 4.9e-11 (n=30) and 1.3e-10 (n=32), SURVEY.md 0."""
    for n in (24, 28, 30, 32):
        rng = np.random.default_rng(0)
        u = np.exp(2j * np.pi * rng.random(n))
        v = np.exp(2j * np.pi * rng.random(n))
        exact = math.factorial(n) * np.prod(u) * np.prod(v)
        ones = np.ones(n, np.int32)
        assert relerr(complex(permanent(np.outer(u, v), ones, ones)), exact) < RTOL, n


def test_full_size_properties_n36():
    """Beyond the reference's validity range (idx_max > 2^31): closed forms and
    algebraic properties at full size."""
    n = 36
    ones = np.ones(n, np.int32)
    rng = np.random.default_rng(0)
    u = np.exp(2j * np.pi * rng.random(n))
    v = np.exp(2j * np.pi * rng.random(n))
    exact = math.factorial(n) * np.prod(u) * np.prod(v)
    assert relerr(complex(permanent(np.outer(u, v), ones, ones)), exact) < 1e-9
    assert relerr(complex(permanent(np.ones((n, n), dtype=complex), ones, ones)),
                  math.factorial(n)) < 1e-9
    assert np.isclose(permanent(np.eye(n, dtype=complex), ones, ones), 1.0)


def test_full_size_n40_closed_form():
    """BASELINE config 5 at full size (2^39 terms, ~8 s on one B200): rank-1
    unit-modulus matrix against n! prod(u) prod(v)."""
    n = 40
    rng = np.random.default_rng(0)
    u = np.exp(2j * np.pi * rng.random(n))
    v = np.exp(2j * np.pi * rng.random(n))
    exact = math.factorial(n) * np.prod(u) * np.prod(v)
    got = complex(permanent(np.outer(u, v), np.ones(n, np.int32), np.ones(n, np.int32)))
    assert relerr(got, exact) < 1e-9


def test_invariances_haar_n26():
    n = 26
    a = haar(n, 5)
    ones = np.ones(n, np.int32)
    base = complex(permanent(a, ones, ones))
    rng = np.random.default_rng(1)
    assert relerr(complex(permanent(np.ascontiguousarray(a[rng.permutation(n)]), ones, ones)), base) < 1e-9
    assert relerr(complex(permanent(np.ascontiguousarray(a.T), ones, ones)), base) < 1e-9
    assert relerr(complex(permanent(1.5j * a, ones, ones)), (1.5j) ** n * base) < 1e-9


def test_complex64_entry():
    a = haar(10, 10).astype(np.complex64)
    ones = np.ones(10, np.int32)
    got = permanent(a, ones, ones)
    assert got.dtype == np.complex64
    want = oracle.permanent(a.astype(np.complex128), ones, ones, precision=1)
    assert relerr(complex(got), want) < 1e-5
    lp = permanent_laplace(a[:9], np.ones(9, np.int32), ones)
    assert lp.dtype == np.complex64 and lp.shape == (10,)


def test_laplace_random_against_the_arbiter():
    rng = np.random.default_rng(31)
    for trial in range(80):
        d = int(rng.integers(1, 10))
        k = int(rng.integers(1, 11))
        rows = rng.multinomial(k - 1, np.ones(d) / d) if k > 1 else np.zeros(d, dtype=int)
        cols = rng.multinomial(k, np.ones(d) / d)
        a = haar(d, 900 + trial)
        if trial % 2:
            a = np.ascontiguousarray(a[np.ix_(rows > 0, cols > 0)])
            rows, cols = rows[rows > 0], cols[cols > 0]
        want = oracle.permanent_laplace(a, rows, cols, precision=1)
        got = permanent_laplace(a, rows, cols)
        assert got.shape == want.shape
        assert np.allclose(got, want, rtol=RTOL, atol=1e-13), (rows, cols)


def test_laplace_minor_identity_at_sampler_sizes():
    """laplace[l] == permanent with column l removed, at the photon steps of
    BASELINE config 4 (k = 20..25 unit columns)."""
    for k in (20, 23, 25):
        a = np.ascontiguousarray(haar(30, k)[: k - 1, :k])
        rows = np.ones(k - 1, np.int32)
        cols = np.ones(k, np.int32)
        lp = permanent_laplace(a, rows, cols)
        for l in (0, k // 2, k - 1):
            minor = complex(permanent(np.ascontiguousarray(np.delete(a, l, axis=1)), rows,
                                      np.ones(k - 1, np.int32)))
            assert relerr(lp[l], minor) < 1e-9, (k, l)


def test_laplace_batch_equals_single_calls():
    from piquasso_b200.sampling import permanent_laplace_batch
    rng = np.random.default_rng(37)
    mats, rws, cls = [], [], []
    for b in range(300):
        d = 12
        k = int(rng.integers(1, 12))
        out = rng.multinomial(k - 1, np.ones(d) / d) if k > 1 else np.zeros(d, int)
        inp = rng.multinomial(k, np.ones(d) / d)
        u = haar(d, b % 7)
        mats.append(np.ascontiguousarray(u[np.ix_(out > 0, inp > 0)]))
        rws.append(out[out > 0])
        cls.append(inp[inp > 0])
    res = permanent_laplace_batch(mats, rws, cls)
    for b in range(0, 300, 7):
        want = oracle.permanent_laplace(mats[b], rws[b], cls[b], precision=1)
        assert res[b].shape == want.shape
        assert np.allclose(res[b], want, rtol=RTOL, atol=1e-13), b


def test_sampler_goldens_identical_samples():
    """Seeded Clifford-Clifford goldens of the reference
    (tests/_simulators/passive/test_measurements.py:237-561): same seed, same
    host RNG draws => identical samples."""
    from piquasso_b200.sampling import generate_samples
    cases = load_golden("sampler.json")
    assert len(cases) >= 4
    for case in cases:
        rejects = list(case["rejects"])
        it = iter(rejects)
        got = generate_samples(case["input"], case["shots"],
                               golden_matrix(case["interferometer"]),
                               case["seed_sequence"], reject_condition=lambda: next(it))
        assert [list(s) for s in got] == case["samples"], case["source"]


def test_pybind_dropin_module():
    """The pybind11 module `permanent` (what replaces piquasso/_math/permanent*.so)
    returns the same values and types as the ctypes mirror."""
    from piquasso_b200.native import permanent as native
    for case in load_golden("permanent_reference_tests.json"):
        m = golden_matrix(case["matrix"])
        want = golden_complex(case["value"])
        got = native.permanent(m, rows=case["rows"], cols=case["cols"])
        assert got.shape == () and got.dtype == (np.complex64 if m.dtype == np.complex64
                                                 else np.complex128)
        assert close(complex(got), want, rtol=1e-5 if m.dtype == np.complex64 else RTOL,
                     atol=1e-6 if m.dtype == np.complex64 else 1e-13), case["source"]
    for case in load_golden("laplace.json")[:40]:
        m = golden_matrix(case["matrix"])
        want = np.array([golden_complex(z) for z in case["value"]])
        got = native.permanent_laplace(matrix=m, rows=case["rows"], cols=case["cols"])
        assert got.shape == want.shape
        assert np.allclose(got, want, rtol=RTOL, atol=1e-13)
    with pytest.raises(RuntimeError):
        native.permanent(np.eye(2, dtype=complex), [1, 1], [1, 0])
    with pytest.raises(TypeError):  # no forcecast: complex128 -> complex64 is refused, and
        native.permanent("not a matrix", [1], [1])  # junk matches no overload


def test_sampler_pmf_rows_match_the_reference_formula():
    """pq_sampler_pmf_c128 against _calculate_pmf (sampling.py:723-749) evaluated
    with the oracle's permanent_laplace: zero filtering, early-out, collisions."""
    from piquasso_b200.sampling import sampler_pmf
    rng = np.random.default_rng(43)
    d = 9
    u = haar(d, 77)
    outs, ins = [], []
    for trial in range(40):
        k = int(rng.integers(1, 8))
        ins.append(rng.multinomial(k, np.ones(d) / d))
        outs.append(rng.multinomial(k - 1, np.ones(d) / d) if k > 1 else np.zeros(d, int))
    got = sampler_pmf(u, np.array(outs), np.array(ins))
    for s in range(len(outs)):
        inz, onz = ins[s] > 0, outs[s] > 0
        part = oracle.permanent_laplace(u[np.ix_(onz, inz)], outs[s][onz], ins[s][inz], precision=1)
        idx = np.arange(d)[inz]
        want = np.zeros(d)
        for m in range(d):
            amp = sum(ins[s][idx[j]] * part[j] * u[m, idx[j]] for j in range(len(part)))
            want[m] = abs(amp) ** 2
        assert np.allclose(got[s], want, rtol=1e-10, atol=1e-15), s


def test_sampler_identical_to_sequential_reference_algorithm():
    """Lock-step sampler == the reference's per-shot loop (sampling.py:208-236)
    driven by the oracle's permanent_laplace, at the shape of BASELINE config 4
    scaled down (40 modes, 10 photons)."""
    from piquasso_b200.sampling import generate_samples
    d, n, seed = 40, 10, 7
    u = haar(d, 40)
    inp = np.array([1] * n + [0] * (d - n))
    got = generate_samples(inp, 6, u, seed)

    def ref_shot(sd):
        rng = np.random.default_rng(sd)
        sample = np.zeros(d, dtype=int)
        cur = np.zeros(d, dtype=int)
        shrink = np.repeat(np.arange(d), inp)
        for _ in range(n):
            ri = rng.choice(len(shrink))
            cur[shrink[ri]] += 1
            shrink = np.delete(shrink, ri)
            nz, oz = cur > 0, sample > 0
            part = oracle.permanent_laplace(u[np.ix_(oz, nz)], sample[oz], cur[nz])
            idx = np.arange(d)[nz]
            pmf = np.empty(d)
            norm = 0.0
            for m in range(d):
                p = 0.0
                for j in range(len(part)):
                    p += cur[idx[j]] * part[j] * u[m, idx[j]]
                pmf[m] = np.abs(p) ** 2
                norm += pmf[m]
            sample[rng.choice(np.arange(d), p=pmf / norm)] += 1
        return tuple(int(x) for x in sample)

    assert got == [ref_shot(seed + i) for i in range(6)]


def test_single_process_multi_device_split(lib):
    """pq_set_devices: one permanent split over the visible devices of one process."""
    import ctypes
    ndev = lib.pq_device_count()
    if ndev < 2:
        pytest.skip("needs two devices")
    a = haar(27, 27)
    ones = np.ones(27, np.int32)
    base = complex(permanent(a, ones, ones))
    ids = (ctypes.c_int32 * ndev)(*range(ndev))
    try:
        _lib.check(lib.pq_set_devices(ids, ndev))
        assert relerr(complex(permanent(a, ones, ones)), base) < 1e-12
    finally:
        _lib.check(lib.pq_set_devices((ctypes.c_int32 * 1)(0), 1))


def test_wide_problems_against_the_arbiter():
    """Widths at the edges of the kernel families: 50-64 active columns (generic
    walk beyond the binary kernel's 48; Laplace lane split S=4 up to NCL=16)."""
    rng = np.random.default_rng(47)
    # permanent: few n-ary rows, many unit columns
    for rows in ([50], [30, 24], [13, 13, 13, 13], [20, 20, 20, 4]):
        rows = np.array(rows)
        n = int(rows.sum())
        a = rng.normal(size=(len(rows), n)) + 1j * rng.normal(size=(len(rows), n))
        cols = np.ones(n, int)
        # high multiplicities cancel catastrophically (sum_g (-1)^g C(r,g) (r+1-2g)^n):
        # the bar is the reference-style double computation measured against the
        # long-double arbiter, not 1e-10
        want = oracle.permanent(a, rows, cols, precision=1)
        ref_err = relerr(oracle.permanent(a, rows, cols), want)
        assert relerr(complex(permanent(a, rows, cols)), want) <= max(1e-10, 20 * ref_err), rows
    # permanent_laplace: NC = 14 (S=2), 30 (S=4), 52, 64 (S=4, NCL=13/16), n-ary rows
    for nc, rows in ((14, [4, 4, 3, 2]), (30, [6, 6, 6, 6, 5]), (52, [13, 13, 13, 12]),
                     (64, [21, 21, 21])):
        rows = np.array(rows)
        assert rows.sum() == nc - 1
        a = (rng.normal(size=(len(rows), nc)) + 1j * rng.normal(size=(len(rows), nc))) / 3
        cols = np.ones(nc, int)
        want = oracle.permanent_laplace(a, rows, cols, precision=1)
        ref_err = np.max(np.abs(oracle.permanent_laplace(a, rows, cols) - want) / np.abs(want))
        got = permanent_laplace(a, rows, cols)
        assert got.shape == (nc,)
        assert np.max(np.abs(got - want) / np.abs(want)) <= max(1e-10, 20 * ref_err), nc
    # column multiplicities on a wide Laplace problem
    a = (rng.normal(size=(3, 20)) + 1j * rng.normal(size=(3, 20))) / 3
    cols = np.array([2, 1] * 10)
    rows = np.array([10, 10, 9])
    want = oracle.permanent_laplace(a, rows, cols, precision=1)
    ref_err = np.max(np.abs(oracle.permanent_laplace(a, rows, cols) - want) / np.abs(want))
    got = permanent_laplace(a, rows, cols)
    assert np.max(np.abs(got - want) / np.abs(want)) <= max(1e-10, 20 * ref_err)


def test_grad_perm_matches_elementwise_permanents():
    """grad[i,j] = rows[i] cols[j] perm(A, rows-e_i, cols-e_j)  (src/permanent.cpp:271-300)."""
    from piquasso_b200.sampling import grad_perm
    rng = np.random.default_rng(53)
    for trial in range(8):
        d = int(rng.integers(1, 6))
        nph = int(rng.integers(1, 7))
        rows = rng.multinomial(nph, np.ones(d) / d)
        cols = rng.multinomial(nph, np.ones(d) / d)
        a = haar(d, 60 + trial)
        got = grad_perm(a, rows, cols)
        for i in range(d):
            for j in range(d):
                if rows[i] == 0 or cols[j] == 0:
                    assert got[i, j] == 0
                    continue
                r2, c2 = rows.copy(), cols.copy()
                r2[i] -= 1
                c2[j] -= 1
                want = rows[i] * cols[j] * oracle.permanent(a, r2, c2, precision=1)
                assert close(got[i, j], want, rtol=1e-10, atol=1e-13), (rows, cols, i, j)
    # finite-difference check of the meaning of "gradient": d perm / d a_ij
    a = haar(4, 9)
    ones = np.ones(4, int)
    g = grad_perm(a, ones, ones)
    eps = 1e-6
    b = a.copy()
    b[1, 2] += eps
    fd = (oracle.permanent(b, ones, ones) - oracle.permanent(a, ones, ones)) / eps
    assert abs(g[1, 2] - fd) < 1e-6


def test_permanent_batch_and_detection_probabilities():
    """One interferometer, many occupation vectors in one call (probability
    tables; the reference loops connector.permanent, passive/utils.py:131-138)."""
    import itertools
    from piquasso_b200.sampling import detection_probabilities, permanent_batch
    rng = np.random.default_rng(59)
    d = 7
    u = haar(d, 21)
    rows, cols = [], []
    for trial in range(64):
        n = int(rng.integers(0, 7))
        rows.append(rng.multinomial(n, np.ones(d) / d))
        cols.append(rng.multinomial(n, np.ones(d) / d))
    got = permanent_batch(u, np.array(rows), np.array(cols))
    for b in range(len(rows)):
        want = oracle.permanent(u, rows[b], cols[b], precision=1)
        assert close(got[b], want, rtol=1e-10, atol=1e-14), (rows[b], cols[b])
    # a full probability table: 3 photons in 5 modes sums to one
    d = 5
    u = haar(d, 5)
    inp = np.array([1, 1, 0, 1, 0])
    outs = [o for o in itertools.product(range(4), repeat=d) if sum(o) == 3]
    p = detection_probabilities(u, inp, outs)
    assert len(outs) == 35 and abs(p.sum() - 1.0) < 1e-12
    # (the reference's golden probabilities, tests/_simulators/passive/test_preparations.py:231-282,
    # run through this entry in test_gpu_envelope.py::test_detection_probabilities_reference_goldens)
    # a sum mismatch in any problem fails the call like the single permanent (src/permanent.cpp:97-104)
    with pytest.raises(RuntimeError):
        permanent_batch(u, [[1, 0, 0, 0, 0]], [[1, 1, 0, 0, 0]])


@pytest.mark.parametrize("photons", [9, 12, 17, 20, 21, 26])
def test_permanent_batch_one_lane_kernels_against_the_arbiter(photons):
    """Batched permanents (probability tables, passive/utils.py:131-138 of the
    reference) walk one lane per Gray segment up to 32 active columns: unit and
    general column flavours, problems of very different sizes in one call (short
    segments for the small ones), against the long-double oracle."""
    from piquasso_b200.sampling import permanent_batch
    rng = np.random.default_rng(100 + photons)
    d = 40
    u = haar(d, 40)
    # unit columns: `photons` single-photon inputs, outputs with collisions
    inp = np.zeros(d, dtype=np.int32)
    inp[rng.choice(d, photons, replace=False)] = 1
    outs = rng.multinomial(photons, np.ones(d) / d, size=6).astype(np.int32)
    outs[0] = 0
    outs[0, :3] = (photons - 2, 1, 1)           # few digits: a handful of terms
    got = permanent_batch(u, outs, inp)
    for b in range(len(outs)):
        want = oracle.permanent(u, outs[b], inp, precision=1)
        assert close(got[b], want, rtol=1e-10, atol=1e-16), (photons, b)
    # general flavour: the same photon number through inputs with multiplicities
    inp2 = rng.multinomial(photons, np.ones(12) / 12).astype(np.int32)
    inp2 = np.concatenate([inp2, np.zeros(d - 12, dtype=np.int32)])
    got = permanent_batch(u, outs[:4], inp2)
    for b in range(4):
        want = oracle.permanent(u, outs[b], inp2, precision=1)
        assert close(got[b], want, rtol=1e-10, atol=1e-16), (photons, b, "general")


@pytest.mark.parametrize("ncols,extra", [(7, 30), (15, 20), (21, 13)])
def test_permanent_batch_general_columns_one_lane(ncols, extra):
    """Batched permanents whose column multiplicities do NOT fit 32 unit columns keep
    the general flavour (run-time powers s_j^c_j) of the one-lane walk: up to 8, 9..20
    and 21..32 distinct columns are separate instantiations."""
    from piquasso_b200.sampling import permanent_batch
    rng = np.random.default_rng(7 * ncols)
    d = 30
    u = haar(d, 30)
    inp = np.zeros(d, dtype=np.int32)
    inp[:ncols] = 1
    inp[:extra] += 1 if extra <= ncols else 0
    if extra > ncols:                      # pile the surplus onto the first columns
        inp[: extra % ncols] += extra // ncols + 1
        inp[extra % ncols: ncols] += extra // ncols
    photons = int(inp.sum())
    assert photons > 32 and np.count_nonzero(inp) == ncols
    outs = np.zeros((3, d), dtype=np.int32)
    for b in range(3):
        outs[b, rng.choice(d, 6, replace=False)] = rng.multinomial(photons, np.ones(6) / 6)
    got = permanent_batch(u, outs, inp)
    for b in range(3):
        want = oracle.permanent(u, outs[b], inp, precision=1)
        assert close(got[b], want, rtol=1e-9, atol=1e-300), (ncols, b, got[b], want)


def test_haar_submatrices_up_to_n28_against_the_arbiter():
    """north_star: relative 1e-10 on complex128 Haar-random unitary submatrices.
    The arbiter is the long-double restatement; the reference's own double
    arithmetic is measured beside it (it is ~1e-10..1e-9 off at these sizes,
    SURVEY.md section 0), so agreement with the reference itself can only be
    asserted at the reference's own error."""
    big = haar(60, 60)
    for n in (16, 22, 26, 28):
        a = np.ascontiguousarray(big[:n, 7:7 + n])
        ones = np.ones(n, np.int32)
        truth = oracle.permanent(a, ones, ones, precision=1, njobs=256)
        got = complex(permanent(a, ones, ones))
        assert relerr(got, truth) < 1e-10, n
        if n <= 26:
            ref_like = oracle.permanent(a, ones, ones, njobs=64)  # the reference's arithmetic
            assert relerr(got, ref_like) < max(1e-10, 3 * relerr(ref_like, truth)), n


def test_concurrent_callers_are_serialised_correctly():
    """dask worker threads call the connector concurrently (Config.use_dask,
    sampling.py:174-188 of the reference); the library serialises them."""
    import threading
    mats = [haar(10 + (i % 5), 200 + i) for i in range(12)]
    want = [oracle.permanent(m, np.ones(len(m), int), np.ones(len(m), int), precision=1) for m in mats]
    got = [None] * len(mats)

    def work(tid):
        for i in range(tid, len(mats), 4):
            for _ in range(5):
                got[i] = complex(permanent(mats[i], np.ones(len(mats[i]), np.int32),
                                           np.ones(len(mats[i]), np.int32)))

    threads = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for g, w in zip(got, want):
        assert relerr(g, w) < 1e-10


def test_sampler_variant_goldens_identical_samples():
    """Post-selected, partially distinguishable and lossy samplers
    through the shot engine with the CUDA pmf call: identical samples to the
    reference run with the same seeds (tests/golden/sampler_variants.json)."""
    from conftest import run_sampler_variant
    for case in load_golden("sampler_variants.json"):
        got = run_sampler_variant(case)
        assert [list(s) for s in got] == case["samples"], case["label"]


def test_sampler_draw_equals_numpy_choice_on_the_pmf_rows():
    """pq_sampler_draw_c128 (normalisation + numpy's cdf search on the device) picks,
    for every shot, exactly the mode the host picks from pq_sampler_pmf_c128's row
    with the same uniform variate -- including variates at the cdf's own values."""
    from piquasso_b200.sampling import sampler_draw, sampler_pmf
    rng = np.random.default_rng(5)
    d = 12
    u = haar(d, 12)
    outs, ins = [], []
    for trial in range(400):
        k = int(rng.integers(1, 9))
        ins.append(rng.multinomial(k, np.ones(d) / d))
        outs.append(rng.multinomial(k - 1, np.ones(d) / d) if k > 1 else np.zeros(d, int))
    outs, ins = np.array(outs), np.array(ins)
    pmf = sampler_pmf(u, outs, ins)
    p = pmf / np.cumsum(pmf, axis=1)[:, -1:]
    cdf = np.cumsum(p, axis=1)
    cdf /= cdf[:, -1:]
    variates = rng.random(len(outs))
    variates[:100] = cdf[np.arange(100), rng.integers(0, d - 1, 100)]  # exactly on a step
    variates[100] = 0.0
    variates[101] = np.nextafter(1.0, 0.0)
    want = (cdf <= variates[:, None]).sum(axis=1)
    for i in range(len(outs)):  # what Generator.choice does with that variate
        assert want[i] == cdf[i].searchsorted(variates[i], side="right")
    got = sampler_draw(u, outs, ins, variates)
    assert np.array_equal(got, want)


def test_sampler_overlapping_batches_give_the_same_samples():
    """Shot batches run by concurrent worker threads (host bookkeeping of one under
    the GPU time of another) return the single-batch sample list."""
    from piquasso_b200.sampling import generate_samples
    d, n, seed = 30, 8, 21
    u = haar(d, 30)
    inp = np.array([1] * n + [0] * (d - n))
    want = generate_samples(inp, 23, u, seed, overlap=1)
    assert generate_samples(inp, 23, u, seed, batch_shots=5, overlap=3) == want
    assert generate_samples(inp, 23, u, seed, batch_shots=4, overlap=2) == want


def test_sampler_step_planned_on_several_host_threads():
    """Thousands of shots per photon step are planned on several host threads
    (merged in shot order): same rows and draws as the same shots in small calls."""
    from piquasso_b200.sampling import sampler_draw, sampler_pmf
    rng = np.random.default_rng(11)
    d, nshots = 10, 5000
    u = haar(d, 10)
    k = rng.integers(1, 7, size=nshots)
    ins = np.array([rng.multinomial(kk, np.ones(d) / d) for kk in k])
    outs = np.array([rng.multinomial(kk - 1, np.ones(d) / d) if kk > 1 else np.zeros(d, int)
                     for kk in k])
    variates = rng.random(nshots)
    big_pmf = sampler_pmf(u, outs, ins)
    big_draw = sampler_draw(u, outs, ins, variates)
    for b in range(0, nshots, 1000):  # below the threading threshold
        sl = slice(b, b + 1000)
        assert np.array_equal(sampler_pmf(u, outs[sl], ins[sl]), big_pmf[sl])
        assert np.array_equal(sampler_draw(u, outs[sl], ins[sl], variates[sl]), big_draw[sl])


def test_laplace_parts_of_all_ranks_sum_to_the_whole(lib):
    """pq_perm_laplace_partial_c128: the term-space split of ONE permanent_laplace
    problem; the parts of 1, 3 and 8 ranks add up to the single call,
    n-ary rows and zero-multiplicity columns included."""
    from piquasso_b200.distributed import _laplace_device_partial
    rng = np.random.default_rng(17)
    cases = [(np.ones(19, np.int32), np.ones(20, np.int32)),
             (np.array([3, 2, 4, 1, 2], np.int32), np.ones(13, np.int32)),
             (np.array([2, 0, 3, 1], np.int32), np.array([2, 0, 1, 3, 1], np.int32))]
    for rows, cols in cases:
        a = (rng.normal(size=(len(rows), len(cols))) + 1j * rng.normal(size=(len(rows), len(cols)))) / 2
        whole = permanent_laplace(a, rows, cols)
        want = oracle.permanent_laplace(a, rows, cols, precision=1)
        assert np.allclose(whole, want, rtol=RTOL, atol=1e-13)
        for nparts in (1, 3, 8):
            total = sum(_laplace_device_partial(a, rows, cols, g, nparts) for g in range(nparts))
            assert np.allclose(total, want, rtol=RTOL, atol=1e-13), (rows, nparts)
    a = haar(3, 3)
    parts = [_laplace_device_partial(a, np.zeros(3, np.int32), np.ones(3, np.int32), g, 2)
             for g in range(2)]
    assert [p.tolist() for p in parts] == [[1.0 + 0j], [0j]]


def test_single_process_sampler_sharded_over_devices(lib):
    """generate_samples(devices=[...]): one host thread per device through
    pq_sampler_draw_dev_c128, same sample list as the single-device run.  With one
    GPU the device is listed twice (two threads, serialised by the device lock)."""
    from piquasso_b200.sampling import generate_samples
    d, n, seed = 24, 7, 5
    u = haar(d, 24)
    inp = np.array([1] * n + [0] * (d - n))
    want = generate_samples(inp, 31, u, seed)
    ndev = lib.pq_device_count()
    devices = list(range(min(ndev, 4))) if ndev > 1 else [0, 0]
    assert generate_samples(inp, 31, u, seed, devices=devices) == want
    assert generate_samples(inp, 31, u, seed, devices=[0]) == want


def test_problems_wider_than_64_columns():
    """65-256 active columns (few rows): permanent, permanent_laplace and the batch
    entries go through the warp-per-segment walk (S = 32 lanes x up to 8 columns).
    Such problems have high row multiplicities and the Glynn sum cancels heavily;
    rank-1 unit-modulus matrices (closed form n! prod u^r prod v) are the
    well-conditioned ones, the bar elsewhere is the reference-style double walk
    measured against the long-double arbiter."""
    from piquasso_b200.sampling import permanent_batch, permanent_laplace_batch
    rng = np.random.default_rng(65)

    def bar(ref_err):
        return max(1e-10, 20 * ref_err)

    def rank1(rows, seed, scale=1.0):
        from fractions import Fraction
        gen = np.random.default_rng(seed)
        rows = np.array(rows)
        n = int(rows.sum())
        u = np.exp(2j * np.pi * gen.random(len(rows)))
        v = np.exp(2j * np.pi * gen.random(n))
        mag = float(Fraction(math.factorial(n)) * Fraction(scale) ** n)
        return scale * np.outer(u, v), rows, np.ones(n, int), mag * np.prod(u ** rows) * np.prod(v)

    # closed forms; 75, 96, 97 and 130 columns = 3, 3, 4 and 5 columns per lane.  How
    # hard the Glynn sum cancels depends on the phases u: the seeds below are ones
    # where the reference-style double walk itself stays below 1e-10.
    for rows, seed in (([40, 35], 10), ([30, 30, 36], 9), ([50, 47], 10), ([70, 60], 10),
                       ([0, 33, 0, 40, 7], 0)):
        a, rows, cols, exact = rank1(rows, seed)
        assert relerr(oracle.permanent(a, rows, cols, precision=1), exact) < 1e-13, rows
        ref_err = relerr(oracle.permanent(a, rows, cols), exact)
        assert ref_err < 1e-10
        assert relerr(complex(permanent(a, rows, cols)), exact) <= bar(ref_err), rows
    # heavier cancellation (the double walk itself is at 5e-6): 65 columns, one row
    a, rows, cols, _ = rank1([65], 1)
    want = oracle.permanent(a, rows, cols, precision=1)
    ref_err = relerr(oracle.permanent(a, rows, cols), want)
    assert relerr(complex(permanent(a, rows, cols)), want) <= bar(ref_err)
    # zero rows / columns in between and column multiplicities (unfiltered call):
    # perm = n! prod u^r prod v^c for the rank-1 matrix
    gen = np.random.default_rng(0)
    rows = np.array([0, 33, 0, 40, 7])
    cols = np.concatenate([np.ones(60, int), [0, 0], 2 * np.ones(10, int), [0]])
    u = np.exp(2j * np.pi * gen.random(5))
    v = np.exp(2j * np.pi * gen.random(len(cols)))
    a = np.outer(u, v)
    exact = float(math.factorial(80)) * np.prod(u ** rows) * np.prod(v ** cols)
    ref_err = relerr(oracle.permanent(a, rows, cols), exact)
    assert ref_err < 1e-10
    assert relerr(complex(permanent(a, rows, cols)), exact) <= bar(ref_err)
    with pytest.raises(RuntimeError):
        permanent(a, rows + 1, cols)  # sum mismatch is still reported
    # batch of wide permanents of one matrix (same closed form, other row vectors)
    rb = np.array([[0, 33, 0, 40, 7], [7, 0, 40, 0, 33], [0, 40, 0, 40, 0]])
    got = permanent_batch(a, rb, cols)
    for b in range(len(rb)):
        w = oracle.permanent(a, rb[b], cols, precision=1)
        e = relerr(oracle.permanent(a, rb[b], cols), w)
        assert relerr(complex(got[b]), w) <= bar(e), b
    # permanent_laplace, 75 / 97 / 130 columns, single call and batch:
    # out_l = (n-1)! prod u^r prod_{j != l} v_j
    probs, exacts = [], []
    for rows, nc in (([40, 34], 75), ([50, 46], 97), ([70, 59], 130)):
        gen = np.random.default_rng(10)
        rows = np.array(rows)
        u = np.exp(2j * np.pi * gen.random(len(rows)))
        v = np.exp(2j * np.pi * gen.random(nc))
        probs.append((np.outer(u, v), rows, np.ones(nc, int)))
        exacts.append(float(math.factorial(int(rows.sum()))) * np.prod(u ** rows) * np.prod(v) / v)
    batch = permanent_laplace_batch(*zip(*probs))
    for (a, rows, cols), exact, from_batch in zip(probs, exacts, batch):
        ref_err = np.max(np.abs(oracle.permanent_laplace(a, rows, cols) - exact) / np.abs(exact))
        assert ref_err < 1e-10
        got = permanent_laplace(a, rows, cols)
        assert got.shape == (len(cols),)
        assert np.max(np.abs(got - exact) / np.abs(exact)) <= bar(ref_err), len(cols)
        assert np.max(np.abs(from_batch - exact) / np.abs(exact)) <= bar(ref_err), len(cols)
    # 256 columns: FP64 cannot resolve such a sum (the reference's own double walk is off
    # by orders of magnitude); what can be checked is that the widest kernel runs and
    # returns finite numbers, and that one column more is a clean error
    a, rows, cols, _ = rank1([86, 85, 85], 2, 0.05)
    assert np.isfinite(complex(permanent(a, rows, cols)))
    with pytest.raises(ValueError):
        permanent(np.ones((1, 257), complex), [257], np.ones(257, int))
