"""The oracle is pinned before it is trusted: the C restatement
(oracle/perm_oracle.c) against the reference's own golden values, against
outputs of the compiled reference (tests/golden/, made by make_golden.py),
against the textbook definition and against the long-double arbiter."""

import math

import numpy as np
import pytest

import oracle
from conftest import golden_complex, golden_matrix, haar, load_golden, relerr


def test_reference_test_goldens():
    """Every permanent() call the reference's tests/_math/test_permanent.py and
    detection-probability tests make (inputs and compiled-reference outputs)."""
    cases = load_golden("permanent_reference_tests.json")
    assert len(cases) >= 10
    for case in cases:
        m = golden_matrix(case["matrix"])
        want = golden_complex(case["value"])
        got = oracle.permanent(m, case["rows"], case["cols"])
        tol = 1e-5 if m.dtype == np.complex64 else 1e-12
        assert abs(got - want) <= tol * max(1.0, abs(want)), case["source"]


def test_reference_literal_goldens():
    """Literal expected values typed in the reference tests
    (tests/_math/test_permanent.py:24-30, 120-126; tests/jax_extensions/
    test_permanent_unit.py:58-131)."""
    assert np.isclose(oracle.permanent(np.array([[4.2]]), [1], [1]), 4.2)
    u = np.array([[1, 1j], [1, -1j]]) / np.sqrt(2)
    assert np.isclose(oracle.permanent(u, [2, 0], [0, 2]), -1)
    assert np.isclose(oracle.permanent(np.zeros((0, 0)), [], []), 1.0)
    assert np.isclose(oracle.permanent(np.array([[0.0]]), [1], [1]), 0.0)
    assert np.isclose(oracle.permanent(np.array([[1, -2], [-3, 4]]), [1, 1], [1, 1]), 10.0)
    assert np.isclose(oracle.permanent(np.full((3, 3), 1e10), [1, 1, 1], [1, 1, 1]), 6e30,
                      rtol=1e-10)
    assert np.isclose(oracle.permanent(np.array([[2, 3], [4, 5]]), [3, 0], [2, 1]), 72.0)
    assert np.isclose(oracle.permanent(np.eye(8), np.ones(8, int), np.ones(8, int)), 1.0)
    assert np.isclose(oracle.permanent(np.random.rand(4, 4), np.zeros(4, int), np.zeros(4, int)),
                      1.0)
    with pytest.raises(RuntimeError):
        oracle.permanent(np.eye(2), [1, 1], [1, 0])


def test_haar_goldens_from_compiled_reference():
    for case in load_golden("permanent_haar.json"):
        m = golden_matrix(case["matrix"])
        want = golden_complex(case["value"])
        got = oracle.permanent(m, case["rows"], case["cols"])
        assert abs(got - want) <= 1e-11 * max(abs(want), 1e-3), case["source"]


def test_laplace_goldens_from_compiled_reference():
    cases = load_golden("laplace.json")
    assert len(cases) > 100
    for case in cases:
        m = golden_matrix(case["matrix"])
        want = np.array([golden_complex(z) for z in case["value"]])
        got = oracle.permanent_laplace(m, case["rows"], case["cols"])
        assert got.shape == want.shape, case["source"]
        assert np.allclose(got, want, rtol=1e-11, atol=1e-14), case["source"]


def test_gray_counter_matches_reference_traces():
    """src/n_aryGrayCodeCounter.hpp initialize()/next() traces of the compiled
    reference counter."""
    for case in load_golden("gray.json"):
        g0, trace = oracle.gray_trace(case["limits"], case["offset"], len(case["trace"]))
        assert g0.tolist() == case["gray0"]
        assert [list(t) for t in trace] == case["trace"]


def test_gray_code_properties():
    rng = np.random.default_rng(0)
    for _ in range(30):
        limits = rng.integers(1, 5, size=rng.integers(1, 6)).tolist()
        total = int(np.prod(limits))
        g0, trace = oracle.gray_trace(limits, 0, total)
        assert len(trace) == total - 1
        seen = {tuple(g0)}
        g = list(g0)
        for changed, prev, value in trace:
            assert abs(prev - value) == 1 and g[changed] == prev
            g[changed] = value
            seen.add(tuple(g))
        assert len(seen) == total  # every code exactly once


def test_against_the_definition():
    rng = np.random.default_rng(1)
    for trial in range(25):
        d = int(rng.integers(1, 5))
        nph = int(rng.integers(1, 6))
        rows = rng.multinomial(nph, np.ones(d) / d)
        cols = rng.multinomial(nph, np.ones(d) / d)
        a = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
        want = oracle.permanent_definition(a, rows, cols)
        got = oracle.permanent(a, rows, cols)
        assert abs(got - want) <= 1e-10 * max(1.0, abs(want))


def test_laplace_is_the_minor_identity():
    """permanent_laplace(A, rows, cols)[l] == permanent(A, rows, cols - e_l)
    (the reference has no direct test of this entry)."""
    rng = np.random.default_rng(2)
    for trial in range(20):
        d = int(rng.integers(2, 6))
        k = int(rng.integers(2, 7))
        rows = rng.multinomial(k - 1, np.ones(d) / d)
        cols = rng.multinomial(k, np.ones(d) / d)
        a = haar(d, trial)
        lap = oracle.permanent_laplace(a, rows, cols)
        for l in range(d):
            if cols[l] == 0:
                continue
            c2 = cols.copy()
            c2[l] -= 1
            assert abs(lap[l] - oracle.permanent(a, rows, c2)) < 1e-11


def test_long_double_arbiter_and_closed_forms():
    """rank-1 A = u v^T has perm = n! prod(u) prod(v); all-ones J_n has n!."""
    for n in (5, 9, 14):
        rng = np.random.default_rng(n)
        u = np.exp(2j * np.pi * rng.random(n))
        v = np.exp(2j * np.pi * rng.random(n))
        a = np.outer(u, v)
        exact = math.factorial(n) * np.prod(u) * np.prod(v)
        ones = np.ones(n, int)
        assert relerr(oracle.permanent(a, ones, ones, precision=1), exact) < 1e-13
        assert relerr(oracle.permanent(a, ones, ones), exact) < 1e-11
        assert relerr(oracle.permanent(np.ones((n, n)), ones, ones, precision=1),
                      math.factorial(n)) < 1e-14


def test_high_multiplicities_binomial_weights_beyond_2_to_63():
    """With multiplicities of a few tens the weight prod C(r_i, g_i) exceeds 2^63
    (C(40,20) * C(34,17) = 3e20): the oracle carries it as a long double, where the
    reference's int overflows already above 2^31.  Closed forms: the rank-1 matrix
    u v^T with multiplicities has perm = n! prod u^r prod v^c; permanent_laplace
    entry l is (n-1)! prod u^r prod_{j != l} v_j."""
    for rows, seed in (([40, 35], 10), ([50, 47], 10), ([0, 33, 0, 40, 7], 0)):
        gen = np.random.default_rng(seed)
        rows = np.array(rows)
        n = int(rows.sum())
        u = np.exp(2j * np.pi * gen.random(len(rows)))
        v = np.exp(2j * np.pi * gen.random(n))
        exact = float(math.factorial(n)) * np.prod(u ** rows) * np.prod(v)
        a = np.outer(u, v)
        assert relerr(oracle.permanent(a, rows, np.ones(n, int), precision=1), exact) < 1e-13
        assert relerr(oracle.permanent(a, rows, np.ones(n, int)), exact) < 1e-10
    gen = np.random.default_rng(10)
    rows = np.array([40, 34])
    u = np.exp(2j * np.pi * gen.random(2))
    v = np.exp(2j * np.pi * gen.random(75))
    exact = float(math.factorial(74)) * np.prod(u ** rows) * np.prod(v) / v
    got = oracle.permanent_laplace(np.outer(u, v), rows, np.ones(75, int), precision=1)
    assert np.max(np.abs(got - exact) / np.abs(exact)) < 1e-13


def test_job_count_does_not_change_the_sum():
    """term(offset) is a pure function of the offset (src/permanent.cpp:158-164):
    any job split sums the same multiset of terms."""
    a = haar(9, 9)
    ones = np.ones(9, int)
    vals = [oracle.permanent(a, ones, ones, njobs=j) for j in (1, 3, 32, 256)]
    for v in vals[1:]:
        assert relerr(v, vals[0]) < 1e-12
    whole, _, idx_max = oracle.partial(a, ones, ones, 0, 256)
    assert idx_max == 256
    halves = oracle.partial(a, ones, ones, 0, 100)[0] + oracle.partial(a, ones, ones, 100, 256)[0]
    assert relerr(halves[0], whole[0]) < 1e-14
    assert relerr(whole[0] / 2 ** 8, vals[0]) < 1e-12


def test_digit_order_does_not_change_the_sum():
    """The sum runs over ALL Gray tuples, so it does not depend on the order of the rows
    (= digits): what the batched hypercube flavour relies on when it moves three rows of
    multiplicity 1 to the lowest digits (piquasso_b200/csrc/pqperm_permhyper.cuh).  Checked
    on the reference's algorithm itself, with repeated rows and columns."""
    rng = np.random.default_rng(17)
    for trial in range(12):
        d = int(rng.integers(4, 9))
        nph = int(rng.integers(3, 10))
        rows = rng.multinomial(nph, np.ones(d) / d)
        cols = rng.multinomial(nph, np.ones(d) / d)
        a = haar(d, 700 + trial)
        base = oracle.permanent(a, rows, cols, precision=1)
        perm = rng.permutation(d)
        moved = oracle.permanent(a[perm], rows[perm], cols, precision=1)
        assert abs(moved - base) <= 1e-13 * abs(base) + 1e-18, (rows, cols, perm)
        if oracle.ref_available():
            ref = oracle.ref_permanent(np.ascontiguousarray(a[perm]), rows[perm], cols)
            assert abs(ref - base) <= 1e-10 * abs(base) + 1e-16


def test_compiled_reference_when_present():
    """Where oracle/_ref exists (it travels to the GPU box) the restatement and
    the unmodified reference agree to the last bits on fresh random inputs."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built and reference sources absent")
    rng = np.random.default_rng(5)
    for trial in range(20):
        d = int(rng.integers(2, 8))
        nph = int(rng.integers(1, 9))
        rows = rng.multinomial(nph, np.ones(d) / d)
        cols = rng.multinomial(nph, np.ones(d) / d)
        a = haar(d, 300 + trial)
        assert relerr(oracle.permanent(a, rows, cols), oracle.ref_permanent(a, rows, cols)) < 1e-12 \
            or abs(oracle.ref_permanent(a, rows, cols)) < 1e-14
