"""Parity at the BASELINE.json shapes and the accuracy envelope (``-m gpu``).

* the double-double arbiter kernel is pinned against CPU binary128 / long-double
  fixtures (tests/golden/arbiter.json, made by make_arbiter_golden.py);
* the production walks are held to the north star's 1e-10 on Haar matrices up to
  n = 32 against those fixtures, and to the stated looser bound beyond, where the
  only arbiter is the pinned GPU kernel (the reference is wrong there,
  src/n_aryGrayCodeCounter.hpp:179);
* BASELINE configs[0..3] at their full shapes against goldens produced by the
  reference itself (tests/golden/baseline_shapes.json, make_baseline_golden.py).

Measured envelope of the production walks on Haar matrices (B200, round 2; the
error comes from the cancellation of the Glynn sum, which grows ~4x per two rows):
n <= 26: <= 3e-14; n = 28/30/32: 1.5e-13 / 2.4e-13 / 1.1e-12; n = 34/36: 7.5e-12 /
4.2e-12.  The bounds asserted below leave an order of magnitude of slack.
"""

import ctypes
import os
import sys
import threading

import numpy as np
import pytest

import oracle
from conftest import ROOT, haar, load_golden, relerr
from piquasso_b200 import _lib, arbiter
from piquasso_b200._math.permanent import permanent

pytestmark = pytest.mark.gpu

# north_star: relative 1e-10 up to n = 32; the looser bound stated beyond
RTOL_NORTH_STAR = 1e-10
RTOL_BEYOND_32 = 1e-9


def _fixture(n, precision):
    for e in load_golden("arbiter.json")["haar"]:
        if e["n"] == n and e["precision"] == precision:
            return complex(*e["hi"]), complex(*e["lo"]), e
    raise KeyError((n, precision))


def _ones(n):
    return np.ones(n, dtype=np.int32)


@pytest.fixture(autouse=True)
def _reset_choices(lib):
    lib.pq_set_kernel_choice(0)
    lib.pq_set_seg_len_hint(0)
    yield
    lib.pq_set_kernel_choice(0)
    lib.pq_set_seg_len_hint(0)


def test_arbiter_is_pinned_against_binary128_and_long_double():
    """pq_perm_arbiter_c128 vs oracle/perm_oracle.c in software binary128
    (n <= 26, agreement far below double precision: 1e-26) and in long double
    (n <= 32, limited by the long double's own 64-bit mantissa: 5e-15)."""
    gold = load_golden("arbiter.json")
    for e in gold["haar"]:
        n = e["n"]
        hi, lo = arbiter.permanent_dd(haar(n, n), _ones(n), _ones(n))
        ghi, glo = complex(*e["hi"]), complex(*e["lo"])
        err = abs((hi - ghi) + (lo - glo)) / abs(ghi)
        assert err < (1e-26 if e["precision"] == 2 else 5e-15), (n, e["precision"], err)
    for e in gold["nary"]:  # repeated rows / columns: the n-ary Gray counter
        rows, cols = np.array(e["rows"], np.int32), np.array(e["cols"], np.int32)
        a = (np.array(e["re"]) + 1j * np.array(e["im"])).reshape(len(rows), len(cols))
        hi, lo = arbiter.permanent_dd(a, rows, cols)
        ghi, glo = complex(*e["hi"]), complex(*e["lo"])
        assert abs((hi - ghi) + (lo - glo)) / abs(ghi) < 1e-26, e["rows"]
        assert arbiter.relerr_vs(complex(permanent(a, rows, cols)), ghi, glo) < RTOL_NORTH_STAR


def test_arbiter_edge_cases_and_coexistence_with_the_walks(lib):
    """Early-outs and the sum check behave like permanent(); a walk that follows an
    arbiter run on the same device finds its segment dispenser re-armed."""
    hi, lo = arbiter.permanent_dd(np.zeros((0, 0), complex), [], [])
    assert hi == 1 and lo == 0
    with pytest.raises(RuntimeError):
        arbiter.permanent_dd(haar(3, 3), [1, 1, 1], [1, 1, 0])
    u = haar(18, 18)
    want = oracle.permanent(u, _ones(18), _ones(18), precision=1)
    for _ in range(2):
        hi, lo = arbiter.permanent_dd(u, _ones(18), _ones(18))
        assert arbiter.relerr_vs(want, hi, lo) < 1e-15
        assert relerr(complex(permanent(u, _ones(18), _ones(18))), want) < 1e-13


def test_haar_up_to_n32_meets_the_north_star():
    """Haar n = 14..32 (BASELINE configs[1] is n = 30, seed 30) against the CPU
    extended-precision fixtures: <= 1e-10 as the north star asks, and <= 2e-11 as
    measured (1.1e-12 at n = 32)."""
    worst = 0.0
    for n in range(14, 33, 2):
        hi, lo, _ = _fixture(n, 1)
        err = arbiter.relerr_vs(complex(permanent(haar(n, n), _ones(n), _ones(n))), hi, lo)
        assert err < RTOL_NORTH_STAR and err < 2e-11, (n, err)
        worst = max(worst, err)
    assert worst > 0.0


def test_baseline_config2_n30_against_the_compiled_reference():
    """configs[1]: the 30x30 Haar permanent (seed 30).  The reference's own C++
    (value committed in tests/golden/arbiter.json) is 1.0e-10 away from the
    long-double arbiter -- its double walk, not ours: we must be closer to the
    arbiter than the reference is, and within the reference's own error of it."""
    hi, lo, e = _fixture(30, 1)
    ref = complex(*e["reference_cpp"])
    ref_err = arbiter.relerr_vs(ref, hi, lo)
    got = complex(permanent(haar(30, 30), _ones(30), _ones(30)))
    got_err = arbiter.relerr_vs(got, hi, lo)
    assert got_err < RTOL_NORTH_STAR
    assert got_err < ref_err               # 2.4e-13 against 1.0e-10
    assert relerr(got, ref) <= ref_err + got_err + 1e-15


def _gpu_fixture(n):
    for e in load_golden("arbiter_gpu.json")["haar"]:
        if e["n"] == n:
            return complex(*e["hi"]), complex(*e["lo"])
    raise KeyError(n)


def test_haar_beyond_n32_against_the_pinned_gpu_arbiter():
    """n = 34: arbiter recomputed here (5 s) and compared with the committed
    fixture of an earlier run (the arbiter is deterministic to ~1e-30), production
    walk within the looser bound.  n = 36: production against the fixture."""
    n = 34
    hi, lo = arbiter.permanent_dd(haar(n, n), _ones(n), _ones(n))
    fhi, flo = _gpu_fixture(n)
    assert abs((hi - fhi) + (lo - flo)) / abs(fhi) < 1e-26
    for n in (34, 36):
        fhi, flo = _gpu_fixture(n)
        err = arbiter.relerr_vs(complex(permanent(haar(n, n), _ones(n), _ones(n))), fhi, flo)
        assert err < RTOL_BEYOND_32 and err < 1e-10, (n, err)   # measured 7.5e-12 / 4.2e-12


def test_full_size_n40_haar_against_the_pinned_gpu_arbiter():
    """BASELINE configs[4]: the bench's own matrix (Haar n = 40, seed 40) against the
    double-double arbiter's value (6 minutes on one B200, committed fixture)."""
    fhi, flo = _gpu_fixture(40)
    got = complex(permanent(haar(40, 40), _ones(40), _ones(40)))
    assert arbiter.relerr_vs(got, fhi, flo) < RTOL_BEYOND_32


def test_cross_partition_agreement_and_its_bound(lib):
    """The same permanent through different cuts of the term space -- segment
    lengths 2^12 / 2^13 / 2^14 and hypercube blocks of 4, 8 and 16 terms -- walks the
    same multiset of terms, but rounds differently: all variants agree with the
    arbiter, and with each other, within the envelope (NOT bit for bit)."""
    n = 32
    hi, lo, _ = _fixture(n, 1)
    u = haar(n, n)
    values = []
    for choice, hint in ((0, 0), (22, 0), (32, 1 << 12), (32, 1 << 13), (32, 1 << 14), (42, 0),
                         (1, 0), (1, 1 << 10)):
        lib.pq_set_kernel_choice(choice)
        lib.pq_set_seg_len_hint(hint)
        v = complex(permanent(u, _ones(n), _ones(n)))
        assert arbiter.relerr_vs(v, hi, lo) < 2e-11, (choice, hint)
        values.append(v)
    spread = max(abs(a - b) for a in values for b in values) / abs(hi)
    assert spread < 4e-11
    # the split over ranks cuts the SAME segments whatever the rank count (the plan
    # no longer depends on nparts): partials of 1, 2, 4 and 8 parts sum to values
    # that agree to the double-double combination, far below the envelope
    lib.pq_set_kernel_choice(0)
    lib.pq_set_seg_len_hint(0)
    import torch
    from piquasso_b200.distributed import combine, finish
    n = 28
    u = np.ascontiguousarray(haar(n, n), dtype=np.complex128)
    sums = []
    for nparts in (1, 2, 4, 8):
        quads = []
        for part in range(nparts):
            out = torch.zeros(4, dtype=torch.float64, device="cuda:0")
            status = ctypes.c_int(0)
            triv = np.zeros(2)
            _lib.check(lib.pq_perm_partial_c128(
                u.ctypes.data_as(_lib.c_double_p), n, n, _ones(n).ctypes.data_as(_lib.c_int32_p),
                _ones(n).ctypes.data_as(_lib.c_int32_p), part, nparts, 0, None,
                ctypes.c_void_p(out.data_ptr()), ctypes.byref(status),
                triv.ctypes.data_as(_lib.c_double_p)))
            quads.append(out.cpu().numpy())
        sums.append(finish(combine(np.array(quads)), n))
    for v in sums[1:]:
        assert relerr(v, sums[0]) < 1e-15


def test_baseline_config1_n20_through_both_bindings():
    """configs[0]: 20x20 Haar (seed 20) via the pybind11 drop-in module and the
    ctypes mirror, against the compiled reference and the long-double arbiter."""
    g = load_golden("baseline_shapes.json")["cfg1_n20"]
    ref, ld = complex(*g["reference_cpp"]), complex(*g["long_double"])
    u = haar(20, 20)
    native = os.path.join(ROOT, "piquasso_b200", "native")
    if native not in sys.path:
        sys.path.insert(0, native)
    import permanent as pyb
    for entry in (pyb.permanent, permanent):
        got = complex(entry(u, _ones(20), _ones(20)))
        assert relerr(got, ld) < 1e-13
        assert relerr(got, ref) < RTOL_NORTH_STAR
    # the literal shape of scripts/permanent_benchmark.py:52-56: symmetrised random
    # matrix, every multiplicity 2 (d = 8 keeps the oracle at 3^7 * 2 terms)
    d = 8
    rng = np.random.default_rng(1)
    a = rng.random((d, d)) + 1j * rng.random((d, d))
    a = a + a.T
    twos = 2 * np.ones(d, dtype=np.int32)
    want = oracle.permanent(a, twos, twos, precision=1)
    assert relerr(complex(pyb.permanent(a, twos, twos)), want) < RTOL_NORTH_STAR


def test_baseline_config3_unfiltered_60_mode_occupations():
    """configs[2]: 60-mode interferometer, 24 photons, the three occupation patterns
    of SURVEY.md 8(d), called UNFILTERED (60x60 with zero multiplicities) as the
    reference does (passive/utils.py:134).  Goldens: the compiled reference and the
    long-double arbiter (make_baseline_golden.py)."""
    g = load_golden("baseline_shapes.json")["cfg3_60modes_24photons"]
    u60 = haar(60, g["haar_seed"])
    assert u60.shape == (60, 60)
    for name, case in g["cases"].items():
        rows, cols = np.array(case["rows"], np.int32), np.array(case["cols"], np.int32)
        assert rows.sum() == 24 and cols.sum() == 24 and len(rows) == 60
        ref, ld = complex(*case["reference_cpp"]), complex(*case["long_double"])
        ref_err = relerr(ref, ld)
        got = complex(permanent(u60, rows, cols))
        # repeated rows cancel harder than unit rows: the bar is the north star, or
        # the reference's own distance from the arbiter where that is larger
        assert relerr(got, ld) <= max(RTOL_NORTH_STAR, 20 * ref_err), name
        assert relerr(got, ref) <= max(RTOL_NORTH_STAR, 20 * ref_err), name


def test_baseline_config4_first_shots_at_100_modes_25_photons():
    """configs[3]: the first shots of the 100-mode / 25-photon Clifford-Clifford run
    are IDENTICAL to the reference's own _generate_samples on its compiled
    permanent_laplace (golden from make_baseline_golden.py), whatever the number
    of shots they are batched with."""
    from piquasso_b200.sampling import generate_samples
    g = load_golden("baseline_shapes.json")["cfg4_sampler_100modes_25photons"]
    u = haar(100, g["haar_seed"])
    want = [tuple(s) for s in g["samples"]]
    assert generate_samples(np.array(g["input"]), len(want), u, g["seed_sequence"]) == want
    more = generate_samples(np.array(g["input"]), 40, u, g["seed_sequence"])
    assert more[: len(want)] == want and all(sum(s) == 25 for s in more)


def test_detection_probabilities_reference_goldens():
    """Golden probabilities of the reference's own tests
    (tests/_simulators/passive/test_preparations.py:231-282) and of a 5-mode state
    evaluated by the reference package (make_baseline_golden.py), through the
    batched entry."""
    from piquasso_b200.sampling import detection_probabilities
    u = np.array([[1, 0, 0],
                  [0, -0.54687158 + 0.07993182j, 0.32028583 - 0.76938896j],
                  [0, 0.78696803 + 0.27426941j, 0.42419041 - 0.35428818j]])
    p = detection_probabilities(u, [1, 1, 0], [[1, 1, 0]])
    assert np.allclose(p, 0.30545762086020883)
    u = np.array([[-0.25022099 + 0.32110177j, -0.69426529 - 0.49960543j, -0.28233272 + 0.15153042j],
                  [-0.69028768 + 0.23351228j, 0.14839865 + 0.49185272j, -0.43153658 - 0.13714903j],
                  [0.41073351 + 0.36681879j, -0.06655274 + 0.00442722j, -0.22146243 - 0.80202711j]])
    p = detection_probabilities(u, [0, 1, 2], [[2, 1, 0]])
    assert np.allclose(p, 0.038483056956364094)
    g = load_golden("baseline_shapes.json")["detection_probabilities"]
    p = detection_probabilities(haar(5, g["haar_seed"]), g["input"], g["outputs"])
    assert np.allclose(p, g["values"], rtol=1e-12, atol=0)


def test_permanent_and_sampler_from_concurrent_threads(lib):
    """One lock story: permanent() and the sampler share a device, its stream and
    its timers; run from two host threads at once they must neither corrupt each
    other's results nor deadlock."""
    from piquasso_b200.sampling import generate_samples
    u20 = haar(20, 20)
    want_perm = oracle.permanent(u20, _ones(20), _ones(20), precision=1)
    u = haar(12, 7)
    inp = np.array([1] * 6 + [0] * 6)
    want_samples = generate_samples(inp, 200, u, 5)
    errors = []

    def perms():
        try:
            for _ in range(150):
                v = complex(permanent(u20, _ones(20), _ones(20)))
                assert relerr(v, want_perm) < 1e-13
                assert lib.pq_last_kernel_ms(0) >= 0.0
        except Exception as exc:  # noqa: BLE001
            errors.append(exc)

    def shots():
        try:
            for _ in range(4):
                assert generate_samples(inp, 200, u, 5) == want_samples
        except Exception as exc:  # noqa: BLE001
            errors.append(exc)

    threads = [threading.Thread(target=perms), threading.Thread(target=shots),
               threading.Thread(target=perms)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert not any(t.is_alive() for t in threads), "deadlock"
    assert not errors, errors


def test_laplace_partials_left_on_the_device_sum_to_the_whole(lib):
    """pq_perm_laplace_partial_dev_c128: the rank shares of one large permanent_laplace
    stay in device memory (what permanent_laplace_allgather hands to NCCL); their sum
    is the single call's result, zero-multiplicity columns and the early-out included."""
    import torch
    from piquasso_b200._math.permanent import permanent_laplace
    rng = np.random.default_rng(11)
    k = 18
    a = np.ascontiguousarray(haar(k + 2, 3)[: k - 1, :k], dtype=np.complex128)
    cases = [(np.ones(k - 1, np.int32), np.ones(k, np.int32))]
    cols0 = np.ones(k, np.int32)
    cols0[[2, 9]] = 0
    rows0 = np.ones(k - 1, np.int32)
    rows0[:2] = 0
    cases.append((rows0, cols0))
    for rows, cols in cases:
        want = permanent_laplace(a, rows, cols)
        total = torch.zeros(2 * k, dtype=torch.float64, device="cuda:0")
        for part in range(3):
            out = torch.zeros(2 * k, dtype=torch.float64, device="cuda:0")
            triv = np.zeros(2)
            n = ctypes.c_int(0)
            _lib.check(lib.pq_perm_laplace_partial_dev_c128(
                a.ctypes.data_as(_lib.c_double_p), k - 1, k,
                rows.ctypes.data_as(_lib.c_int32_p), cols.ctypes.data_as(_lib.c_int32_p),
                part, 3, 0, ctypes.c_void_p(out.data_ptr()),
                triv.ctypes.data_as(_lib.c_double_p), ctypes.byref(n)))
            assert n.value == k and np.isnan(triv[0])
            total += out
        got = total.cpu().numpy().view(np.complex128)
        assert np.allclose(got, want, rtol=1e-12, atol=0)
    out = torch.zeros(2 * k, dtype=torch.float64, device="cuda:0")
    triv = np.zeros(2)
    n = ctypes.c_int(0)
    _lib.check(lib.pq_perm_laplace_partial_dev_c128(
        a.ctypes.data_as(_lib.c_double_p), k - 1, k,
        np.zeros(k - 1, np.int32).ctypes.data_as(_lib.c_int32_p),
        np.ones(k, np.int32).ctypes.data_as(_lib.c_int32_p), 0, 3, 0,
        ctypes.c_void_p(out.data_ptr()), triv.ctypes.data_as(_lib.c_double_p), ctypes.byref(n)))
    assert n.value == 1 and triv[0] == 1.0 and triv[1] == 0.0


def test_round2_abi_entries(lib):
    """pq_set_timing, pq_sampler_work / _detail and the explicit-device pmf entry."""
    from piquasso_b200.sampling import sampler_pmf
    u = haar(16, 16)
    want = complex(permanent(u, _ones(16), _ones(16)))
    assert lib.pq_last_kernel_ms(0) > 0.0
    try:
        lib.pq_set_timing(0)
        assert complex(permanent(u, _ones(16), _ones(16))) == want   # same launch, same bits
        assert lib.pq_last_kernel_ms(0) == -1.0
    finally:
        lib.pq_set_timing(1)
    assert complex(permanent(u, _ones(16), _ones(16))) == want
    assert lib.pq_last_kernel_ms(0) > 0.0
    # sampler work counters: 3 shots with 4 photons placed and a 5th input photon chosen:
    # Laplace problems of 5 columns and 3 binary digits = 8 Gray-code terms each
    d = 9
    U = haar(d, 4)
    out_occ = np.zeros((3, d), np.int32)
    in_occ = np.zeros((3, d), np.int32)
    out_occ[:, [0, 2, 5, 7]] = 1
    in_occ[:, :5] = 1
    lib.pq_sampler_work_reset()
    rows = sampler_pmf(U, out_occ, in_occ)
    rows_dev = sampler_pmf(U, out_occ, in_occ, device=0)
    assert np.array_equal(rows, rows_dev)
    work = (ctypes.c_double * 2)()
    lib.pq_sampler_work(work)
    assert work[0] == 2 * 3 * 8 and work[1] == 2 * 3 * 8 * 22 * 5
    detail = (ctypes.c_double * 8)()
    lib.pq_last_sampler_detail(detail)
    assert all(x >= 0.0 for x in detail) and detail[3] > 0.0
    from conftest import oracle_pmf_rows
    assert np.allclose(rows, oracle_pmf_rows(U, out_occ, in_occ), rtol=1e-10, atol=1e-16)
