"""Host-side logic, no kernel launches: ABI surface, planning, Gray
enumeration, overload resolution and error behaviour of the Python mirror."""

import os
import re

import numpy as np
import pytest

import oracle
from conftest import ROOT, haar
from piquasso_b200 import _lib, plan
from piquasso_b200._math.permanent import permanent, permanent_laplace


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "pqperm.h")).read()
    declared = set(re.findall(r"\b(pq_[a-z0-9_]+)\s*\(", header))
    declared.discard("pq_plan_info")
    bound = {name for name, _, _ in _lib.SIGNATURES}
    assert declared == bound, declared ^ bound
    for name in declared:
        assert hasattr(lib, name)


def test_plan_all_ones():
    for n in (2, 5, 20, 30, 40):
        p = plan.plan(np.ones(n, int), np.ones(n, int))
        assert p["idx_max"] == 2 ** (n - 1)      # src/permanent.cpp:131-142
        assert p["nseg"] * p["seg_len"] == p["idx_max"]
        assert p["active_rows"] == n - 1 and p["active_cols"] == n
        assert p["flops_per_term"] == 8 * n + 2
        assert p["sum_rows"] == n and p["trivial"] == 0
    assert plan.plan(np.ones(40, int), np.ones(40, int))["kernel"] == 2
    assert plan.plan(np.ones(24, int), np.ones(24, int))["kernel"] == 2
    assert plan.plan(np.ones(20, int), np.ones(20, int))["kernel"] == 1  # small: generic walk
    assert plan.plan([2, 1, 0, 3], [1, 1, 4, 0])["kernel"] == 1


def test_plan_idx_max_matches_oracle():
    rng = np.random.default_rng(3)
    for _ in range(50):
        d = int(rng.integers(1, 9))
        n = int(rng.integers(1, 12))
        rows = rng.multinomial(n, np.ones(d) / d)
        cols = rng.multinomial(n, np.ones(d) / d)
        p = plan.plan(rows, cols)
        _, _, idx_max = oracle.gray_of_offset(rows, 0)
        assert p["idx_max"] == idx_max
        assert p["nseg"] * p["seg_len"] == idx_max


def test_plan_trivial_and_errors(lib):
    assert plan.plan([0, 0, 0], [0, 0, 0])["trivial"] == 1
    assert plan.plan([], [])["trivial"] == 1
    with pytest.raises(_lib.PqPermError) as e:
        plan.plan([1, 1], [1, 0])
    assert e.value.code == _lib.PQ_ERR_SUM_MISMATCH
    with pytest.raises(_lib.PqPermError) as e:
        plan.plan(np.ones(70, int), np.ones(70, int))
    assert e.value.code == _lib.PQ_ERR_TOO_LARGE
    with pytest.raises(_lib.PqPermError) as e:
        plan.plan([-1, 2], [1, 0])
    assert e.value.code == _lib.PQ_ERR_BAD_ARG


def test_gray_enumeration_matches_oracle():
    """Identical Gray-code term enumeration: offset -> digits, against the
    restated reference counter (src/n_aryGrayCodeCounter.hpp:170-194)."""
    rng = np.random.default_rng(11)
    for _ in range(60):
        d = int(rng.integers(1, 9))
        n = int(rng.integers(1, 14))
        rows = rng.multinomial(n, np.ones(d) / d)
        _, _, idx_max = oracle.gray_of_offset(rows, 0)
        for off in {0, idx_max - 1, *rng.integers(0, idx_max, size=6).tolist()}:
            want, _, _ = oracle.gray_of_offset(rows, int(off))
            got = plan.gray_of_offset(rows, int(off))
            assert got.tolist() == want.tolist(), (rows, off)


def test_gray_enumeration_beyond_2_pow_31():
    """The reference truncates offsets to int (n_aryGrayCodeCounter.hpp:179);
    the 64-bit restatement and the GPU path's host mirror agree above it."""
    rows = np.ones(40, int)
    for off in (2 ** 31, 2 ** 31 + 12345, 2 ** 38 + 7, 2 ** 39 - 1):
        want, _, idx_max = oracle.gray_of_offset(rows, off)
        assert idx_max == 2 ** 39
        assert plan.gray_of_offset(rows, off).tolist() == want.tolist()
        # binary reflected code: g = c ^ (c >> 1) on digits 1..39 (digit 0 is the
        # radix-1 digit left by the row split)
        g = off ^ (off >> 1)
        assert want[0] == 0 and all(int(want[i + 1]) == (g >> i) & 1 for i in range(39))


def test_segment_ranges_tile_the_term_space():
    for nseg in (1, 7, 1024, 2 ** 25 + 3):
        for nparts in (1, 2, 3, 8):
            edges = [plan.segment_range(nseg, g, nparts) for g in range(nparts)]
            assert edges[0][0] == 0 and edges[-1][1] == nseg
            assert all(edges[i][1] == edges[i + 1][0] for i in range(nparts - 1))
            sizes = [e - b for b, e in edges]
            assert max(sizes) - min(sizes) <= 1


def test_early_outs_need_no_device_and_keep_dtype():
    """src/permanent.cpp:106-108 / src/permanent_laplace.cpp:52-57."""
    z3 = np.zeros(3, int)
    for dtype in (np.complex64, np.complex128):
        v = permanent(np.ones((3, 3), dtype=dtype), rows=z3, cols=z3)
        assert v.shape == () and v.dtype == dtype and v == 1
    v = permanent(np.ones((3, 3)), z3, z3)            # float64 -> complex128 overload
    assert v.dtype == np.complex128
    v = permanent(np.ones((3, 3), dtype=np.float32), z3, z3)  # float32 -> complex64 overload
    assert v.dtype == np.complex64
    lp = permanent_laplace(np.ones((0, 0), dtype=np.complex128), [], [])
    assert lp.shape == (1,) and lp[0] == 1
    lp = permanent_laplace(np.ones((2, 3), dtype=np.complex128), [0, 0], [1, 0, 0])
    assert lp.shape == (1,) and lp[0] == 1            # sum(rows) == 0: length-1 result


def test_argument_errors():
    with pytest.raises(RuntimeError):                 # the reference throws too
        permanent(np.eye(2, dtype=complex), [1, 1], [1, 0])
    with pytest.raises(ValueError):
        permanent(np.eye(2, dtype=complex), [1, 1, 1], [1, 1])
    with pytest.raises(ValueError):
        permanent(np.ones(4, dtype=complex), [1], [1])
    with pytest.raises(TypeError):
        permanent(np.array([["a", "b"], ["c", "d"]]), [1, 1], [1, 1])
    with pytest.raises(TypeError):                    # no safe cast to complex128
        permanent(np.ones((2, 2), dtype=np.clongdouble), [1, 1], [1, 1])


def test_compute_without_a_device_fails_loudly(lib):
    if lib.pq_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_lib.PqPermError) as e:
        permanent(haar(3, 1), [1, 1, 1], [1, 1, 1])
    assert e.value.code == _lib.PQ_ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)
    with pytest.raises(_lib.PqPermError):
        permanent_laplace(haar(3, 1), [1, 1, 0], [1, 1, 1])


def test_finish_scaling_and_error_free_combine():
    from piquasso_b200.distributed import combine, finish
    v = finish([3.0, 0.5, -8.0, 0.0], 4)
    assert v == complex(3.5 / 8, -1.0)
    # partials far larger than their sum: a plain sum of the hi parts returns 0
    quads = [[1e20, 0.0, 1.0, 0.0], [1.0, 1e-20, 1e-17, 0.0], [-1e20, 0.0, -1.0, 0.0]]
    out = combine(quads)
    assert out[0] + out[1] == 1.0 and out[2] + out[3] == 1e-17


def test_pybind_module_surface():
    """Same module name, callables, keyword names and overload order as the
    reference binding (piquasso/_math/permanent.cpp:71-85)."""
    from piquasso_b200.native import permanent as native
    assert native.__name__.endswith("permanent")
    doc = native.permanent.__doc__
    assert doc.index("numpy.complex64") < doc.index("numpy.complex128")
    assert "matrix" in doc and "rows" in doc and "cols" in doc
    z = np.zeros(3, int)
    assert native.permanent(np.ones((3, 3), dtype=np.complex64), rows=z, cols=z).dtype == np.complex64
    assert native.permanent(np.ones((3, 3)), rows=z, cols=z).dtype == np.complex128
    assert native.permanent(np.ones((3, 3), dtype=np.int64), z, z).dtype == np.complex128
    out = native.permanent_laplace(np.ones((0, 0), dtype=complex), [], [])
    assert out.shape == (1,) and out[0] == 1
    with pytest.raises(RuntimeError):
        native.permanent(np.eye(2, dtype=complex), [1, 1], [1, 0])


def test_numpy_choice_equivalences():
    """The lock-step sampler issues Generator.choice's two forms as integers() and
    random() + searchsorted; both must consume the bit stream exactly like choice
    (numpy/random/_generator.pyx) or the samples would differ from the reference's."""
    for seed in range(200):
        a = np.random.default_rng(seed)
        b = np.random.default_rng(seed)
        for m in (25, 24, 7, 3, 2, 1, 13, 100):
            assert a.choice(m) == b.integers(0, m)
        p = np.random.default_rng(seed + 1000).random(50)
        p /= p.sum()
        x = a.choice(np.arange(50), p=p)
        cdf = p.cumsum()
        cdf /= cdf[-1]
        u = b.random()
        assert x == cdf.searchsorted(u, side="right") == (cdf <= u).sum()
        assert a.random() == b.random()


def test_connector_factory_reports_missing_piquasso():
    """piquasso itself is not a dependency: the plugin connector is built on demand."""
    import importlib.util
    from piquasso_b200.connector import make_connector
    if importlib.util.find_spec("piquasso") is None:
        with pytest.raises(ImportError, match="piquasso is not importable"):
            make_connector()
    else:
        conn = make_connector()
        assert hasattr(conn, "permanent") and hasattr(conn, "permanent_laplace")


def test_plain_c_program_links_against_the_abi(tmp_path):
    """include/pqperm.h + libpqperm.so from a C translation unit (gcc -std=c99)."""
    import subprocess
    from piquasso_b200 import _lib as L
    exe = tmp_path / "c_abi_smoke"
    libdir = os.path.dirname(L.LIB_PATH)
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic",
                    "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c_abi_smoke.c"), "-o", str(exe),
                    "-L", libdir, "-lpqperm", "-Wl,-rpath," + libdir], check=True)
    proc = subprocess.run([str(exe)], capture_output=True, text=True)
    assert proc.returncode == 0, (proc.returncode, proc.stdout, proc.stderr)
    assert "c abi ok" in proc.stdout


# ---- sampler host logic with the oracle standing in for the GPU call ---------

def test_lockstep_sampler_host_logic_against_reference_goldens():
    """generate_samples' vectorised host bookkeeping (per-shot RNG order, input
    growth, normalisation, draw) on the reference's seeded goldens
    (tests/_simulators/passive/test_measurements.py:237-561), the pmf rows coming
    from the oracle instead of pq_sampler_pmf_c128."""
    from conftest import golden_matrix, load_golden, oracle_pmf_rows
    from piquasso_b200.sampling import generate_samples
    for case in load_golden("sampler.json"):
        if len(case["input"]) > 8 or case["shots"] > 200:
            continue  # keep the CPU suite short; the GPU suite replays all of them
        it = iter(list(case["rejects"]))
        got = generate_samples(case["input"], case["shots"],
                               golden_matrix(case["interferometer"]), case["seed_sequence"],
                               reject_condition=lambda: next(it), pmf_rows=oracle_pmf_rows)
        assert [list(s) for s in got] == case["samples"], case["source"]


def test_sampler_variants_host_logic_against_reference_goldens():
    """Post-selected / partially distinguishable / lossy per-shot algorithms
    (piquasso/_simulators/passive/sampling.py:110-146, 239-529) as shot coroutines:
    identical samples to the reference run with the same seeds
    (tests/golden/sampler_variants.json, made by make_golden.py section 3b)."""
    from conftest import load_golden, oracle_pmf_rows, run_sampler_variant
    cases = load_golden("sampler_variants.json")
    assert len(cases) >= 8
    for case in cases:
        got = run_sampler_variant(case, pmf_rows=oracle_pmf_rows)
        assert [list(s) for s in got] == case["samples"], case["label"]


def test_postselection_gives_up_after_max_trials():
    from conftest import haar, oracle_pmf_rows
    from piquasso_b200.sampling import generate_samples
    from piquasso_b200.shot_engine import InvalidSimulation
    u = haar(4, 4)
    with pytest.raises(InvalidSimulation):
        # three photons can never leave four in mode 0
        generate_samples([1, 1, 1, 0], 2, u, 3, postselect_data=((0,), (4,), 5),
                         pmf_rows=oracle_pmf_rows)


def test_expanded_interferometer_embeds_the_lossy_matrix_isometrically():
    from conftest import haar
    from piquasso_b200.shot_engine import expanded_interferometer
    lossy = haar(5, 1) @ np.diag(np.sqrt([0.9, 0.2, 0.5, 0.7, 0.6])) @ haar(5, 2)
    big = expanded_interferometer(lossy)
    assert big.shape == (10, 10)
    assert np.allclose(big[:5, :5], lossy)
    # the photons enter in the first d modes only: those columns must be orthonormal
    # (the reference's [[S, C], [C, S]] middle factor is not unitary as a whole)
    left = big[:, :5]
    assert np.allclose(left.conj().T @ left, np.eye(5), atol=1e-12)


def test_shot_streams_replay_numpy_generators():
    """piquasso_b200.shot_rng.ShotStreams derives random() and integers(0, high)
    from the raw PCG64 stream exactly as numpy's Generator does, for any
    interleaving (32-bit half buffering), for high == 1 (no draw), through the
    Lemire rejection loop (large bounds) and past the pre-drawn chunk."""
    from piquasso_b200.shot_rng import ShotStreams
    nshots, seed0 = 300, 1234
    streams = ShotStreams(seed0, 5, 5 + nshots, draws_per_shot=6)
    gens = [np.random.default_rng(seed0 + idx) for idx in range(5, 5 + nshots)]
    plan = np.random.default_rng(0)
    everyone = np.arange(nshots)
    for step in range(60):
        live = everyone if step % 3 else np.flatnonzero(plan.random(nshots) < 0.7)
        kind = plan.integers(0, 4)
        if kind == 0:
            want = np.array([gens[s].random() for s in live])
            assert np.array_equal(streams.random(live), want)
        elif kind == 1:   # the sampler's bounds
            high = plan.integers(1, 27, size=live.size)
            want = np.array([gens[s].integers(0, h) for s, h in zip(live, high)])
            assert np.array_equal(streams.integers(live, high), want)
        elif kind == 2:   # bounds where Lemire's rejection actually triggers
            high = plan.choice([3 << 30, (1 << 32) - 1, (1 << 31) + 12345, 1 << 32, 1],
                               size=live.size)
            want = np.array([gens[s].integers(0, h) for s, h in zip(live, high)])
            assert np.array_equal(streams.integers(live, high), want)
        else:             # choice(n) is integers(0, n)
            want = np.array([gens[s].choice(13) for s in live])
            assert np.array_equal(streams.integers(live, 13), want)


def test_sampler_level_dropin_adapters():
    """piquasso_b200.integration: the adapters keep the reference's signatures
    (sampling.py:33-42, 110-117), recognise its `lambda: False`, and install()
    swaps / restores both bindings.  (Whole-simulator equality with the reference
    is asserted where the reference is importable: tests/golden/make_golden.py 3c.)"""
    from types import SimpleNamespace
    from conftest import golden_matrix, load_golden, oracle_pmf_rows
    from piquasso_b200 import integration

    assert integration._never_rejects(None)
    assert integration._never_rejects(lambda: False)
    shared = np.random.default_rng(0)
    assert not integration._never_rejects(lambda: shared.uniform() > 0.5)
    assert not integration._never_rejects(lambda: True)

    def stock_generate_samples(*a, **k):
        raise AssertionError("the stock sampler must not run while patched")

    fake_sampling = SimpleNamespace(generate_samples=stock_generate_samples,
                                    generate_lossy_samples=stock_generate_samples)
    fake_steps = SimpleNamespace(generate_samples=stock_generate_samples, unrelated=1)
    case = next(c for c in load_golden("sampler_variants.json")
                if c["label"] == "postselect one mode")
    config = SimpleNamespace(seed_sequence=case["seed_sequence"], use_dask=False)
    postselect = (tuple(case["postselect_modes"]), tuple(case["postselect_photons"]),
                  case["max_trials"])
    with integration.install(modules=[fake_sampling, fake_steps], pmf_rows=oracle_pmf_rows):
        got = fake_steps.generate_samples(
            input=np.array(case["input"]), shots=case["shots"],
            calculate_permanent_laplace=None,
            interferometer=golden_matrix(case["interferometer"]),
            reject_condition=lambda: False, postselect_data=postselect,
            uniform_particle_overlap=None, config=config)
        assert fake_sampling.generate_lossy_samples is not stock_generate_samples
    assert [list(s) for s in got] == case["samples"]
    assert fake_steps.generate_samples is stock_generate_samples
    assert fake_sampling.generate_lossy_samples is stock_generate_samples
    with pytest.raises(ImportError):
        integration.install(modules=[SimpleNamespace()])


def test_plan_expands_multiplicity_columns_without_touching_the_enumeration(lib):
    """Columns of multiplicity c are handed to the kernels as c unit columns
    (cols_padded covers sum(c)), the reported algorithmic figures and the
    offset -> Gray-digit map stay those of the caller's problem."""
    rows = np.array([2, 0, 1, 3, 1])
    cols = np.array([3, 0, 2, 1, 1])
    p = plan.plan(rows, cols)
    assert p["active_cols"] == 4 and p["cols_padded"] == 8  # 7 unit columns, padded to 8
    assert p["flops_per_term"] == 2 * 4 + 6 * 7 + 2
    _, _, idx_max = oracle.gray_of_offset(rows, 0)
    assert p["idx_max"] == idx_max
    for off in range(0, idx_max, 5):
        want, _, _ = oracle.gray_of_offset(rows, off)
        assert list(plan.gray_of_offset(rows, off)) == list(want)
    # too many photons to expand: the multiplicity columns stay as they are
    wide = plan.plan(np.array([40, 40]), np.array([2] * 40))
    assert wide["active_cols"] == 40 and wide["cols_padded"] == 40


def test_sampler_batch_bounds():
    from piquasso_b200.sampling import _batch_bounds
    assert _batch_bounds(10, None, 1) == [(0, 10)]
    assert _batch_bounds(0, None, 1) == []
    assert _batch_bounds(10, 4, 1) == [(0, 4), (4, 8), (8, 10)]
    assert _batch_bounds(500, None, 2) == [(0, 500)]  # too few shots to split
    assert _batch_bounds(10001, None, 2) == [(0, 5000), (5000, 10001)]  # two halves
    bounds = _batch_bounds(10000, None, 3)
    assert bounds[0][0] == 0 and bounds[-1][1] == 10000 and len(bounds) == 4
    assert all(a[1] == b[0] for a, b in zip(bounds, bounds[1:]))


def test_sampler_edge_cases_on_the_host():
    """Empty inputs, zero shots, everything rejected, every variant: shapes and
    photon counts come out right (pmf rows from the oracle)."""
    from conftest import haar, oracle_pmf_rows
    from piquasso_b200.sampling import generate_lossy_samples, generate_samples
    u = haar(4, 4)
    kw = dict(pmf_rows=oracle_pmf_rows)
    assert generate_samples([0, 0, 0, 0], 3, u, 1, **kw) == [(0, 0, 0, 0)] * 3
    assert generate_samples([1, 0, 1, 0], 0, u, 1, **kw) == []
    assert generate_samples([1, 0, 1, 0], 2, u, 1, reject_condition=lambda: True, **kw) == \
        [(0, 0, 0, 0)] * 2
    for smp in generate_samples([1, 0, 1, 0], 5, u, 1, uniform_particle_overlap=0.5, **kw):
        assert len(smp) == 4 and sum(smp) == 2
    assert generate_samples([0, 0, 0, 0], 2, u, 1, uniform_particle_overlap=0.5, **kw) == \
        [(0, 0, 0, 0)] * 2
    for smp in generate_samples([1, 0, 1, 0], 5, u, 1, postselect_data=((1,), (0,), 100), **kw):
        assert len(smp) == 3 and sum(smp) == 2  # the post-selected mode is removed
    for smp in generate_lossy_samples([1, 0, 1, 0], 5, 0.7 * u, 1, **kw):
        assert len(smp) == 4 and sum(smp) <= 2
    assert len(generate_samples([3, 0, 0, 0], 2, u, 1, devices=[0], **kw)) == 2


def test_bulk_pcg64_streams_equal_numpy_bit_generators(lib):
    """pq_pcg64_streams restates numpy's SeedSequence + PCG64 seeding on the host: raw
    outputs identical to np.random.PCG64(seed).random_raw for seeds on both sides of
    the 32-bit word boundary and up to 2^64 (the per-shot generators of the reference's
    sampler, piquasso/_simulators/passive/sampling.py:149-194)."""
    import ctypes
    for seed0, n, draws in ((0, 5, 3), (123, 64, 17), (2 ** 32 - 40, 80, 5), (2 ** 47 + 1, 9, 60),
                            (2 ** 64 - 12, 11, 4)):
        out = np.empty((n, draws), dtype=np.uint64)
        rc = lib.pq_pcg64_streams(ctypes.c_uint64(seed0), n, draws,
                                  out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)))
        assert rc == 0
        want = np.array([np.random.PCG64(seed0 + i).random_raw(draws) for i in range(n)])
        assert np.array_equal(out, want), seed0
    # seeds the library does not restate fall back to real numpy generators
    from piquasso_b200.shot_rng import ShotStreams
    big = ShotStreams(2 ** 70, 0, 3, 4)
    want = np.random.default_rng(2 ** 70 + 1).random()
    assert big.random(np.array([1]))[0] == want
    # a stream that runs out of its first chunk is regenerated, longer, identically
    s = ShotStreams(7, 0, 4, 4)
    rng = np.random.default_rng(7 + 2)
    for _ in range(20):
        assert s.random(np.array([2]))[0] == rng.random()


def test_batch_plan_flavours_and_term_counts():
    """Planner of pq_perm_batch_c128 (no GPU): the hypercube flavour needs unit columns
    (after the column multiplicities were written out) and three rows of multiplicity 1
    after the split; it always keeps those three digits low (segments are multiples of 8
    terms); segments tile the term space exactly; column multiplicities that do not fit
    32 unit columns keep the general flavour."""
    from piquasso_b200 import plan as pqplan
    rng = np.random.default_rng(11)
    seen = set()
    for trial in range(300):
        d = int(rng.integers(3, 40))
        n = int(rng.integers(1, 45))
        rows = rng.multinomial(n, np.ones(d) / d)
        k = int(rng.integers(1, d + 1))
        cols = np.zeros(d, dtype=int)
        cols[rng.choice(d, k, replace=False)] = rng.multinomial(n, np.ones(k) / k)
        nprob = int(rng.choice([1, 10, 1000, 100000]))
        p = pqplan.batch_plan(rows, cols, nprob)
        # reference term count: prod(r_i + 1) after the split (src/permanent.cpp:131-142)
        r = rows[rows > 0].copy()
        r[np.argmin(r)] -= 1
        idx_max = int(np.prod(r + 1, dtype=object))
        assert p["idx_max"] == idx_max and p["seg_len"] * p["nseg"] == idx_max
        nc, m = int(np.count_nonzero(cols)), int(cols.sum())
        unit_after = m <= 32
        ones = int(np.count_nonzero(r == 1))
        width = m if unit_after else nc
        want_hyper = unit_after and ones >= 3 and 4 <= width <= 32
        assert (p["kernel"] == 4) == want_hyper, (rows, cols, p)
        assert p["active_cols"] == width
        if p["kernel"] == 4:
            assert p["seg_len"] % 8 == 0 and p["low_digits"] >= 3
        if nc <= 32:
            assert p["cols_padded"] == width       # one lane per segment, no padding
        seen.add((p["kernel"], unit_after))
    assert {(3, True), (4, True), (3, False)} <= seen
    # early-outs and errors as for the single permanent (src/permanent.cpp:97-108)
    assert pqplan.batch_plan([0, 0], [0, 0])["trivial"] == 1
    with pytest.raises(RuntimeError):
        pqplan.batch_plan([1, 0], [1, 1])


def test_factorial_table_equals_scipy():
    from scipy.special import factorial
    from piquasso_b200.sampling import _factorials
    occ = np.random.default_rng(1).integers(0, 40, size=(64, 9))
    assert np.array_equal(_factorials(occ), factorial(occ))
    big = np.array([0, 170, 171, 200])
    assert np.array_equal(_factorials(big), factorial(big))
