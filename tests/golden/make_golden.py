"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

What it does
------------
* imports the reference's Python package from ``/root/reference`` with its one
  missing dependency stubbed (``blackbird``), its unrelated native modules
  stubbed (torontonian, pfaffian), and ``piquasso._math.permanent`` bound to the
  UNMODIFIED reference C++ compiled by ``oracle/build.py``
  (``oracle/_ref/libpqref.so``);
* runs the reference's OWN tests for the path -- every parameter-less test of
  ``tests/_math/test_permanent.py`` and the seeded sampler / detection
  probability goldens of ``tests/_simulators/passive`` -- which assert their
  own golden values, and records every ``permanent`` / ``permanent_laplace`` /
  ``generate_samples`` call made with its inputs and its output;
* adds seeded Haar-random cases (inputs + reference output) for sizes the
  reference finishes in seconds.

Nothing here is read at test time except the JSON it writes.
"""

from __future__ import annotations

import importlib.util
import inspect
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REFERENCE = "/root/reference"
sys.path.insert(0, REPO)
sys.path.insert(0, REFERENCE)

import oracle  # noqa: E402

CALLS = []  # (kind, matrix, rows, cols, value)


def _install_stubs():
    bb = types.ModuleType("blackbird")

    class BlackbirdProgram:  # piquasso/api/program.py:18 only needs the name
        pass

    bb.BlackbirdProgram = BlackbirdProgram
    bb.load = lambda *a, **k: None
    bb.loads = lambda *a, **k: None
    sys.modules["blackbird"] = bb

    perm = types.ModuleType("piquasso._math.permanent")

    def permanent(matrix, rows, cols):
        m = np.asarray(matrix)
        if m.dtype == np.complex64:
            v = oracle.ref_permanent_c64(m, rows, cols)
            out = np.array(np.complex64(v))
        else:
            v = oracle.ref_permanent(m, rows, cols)
            out = np.array(np.complex128(v))
        CALLS.append(("permanent", np.array(m), np.array(rows), np.array(cols), complex(v)))
        return out

    def permanent_laplace(matrix, rows, cols):
        v = oracle.ref_permanent_laplace(matrix, rows, cols)
        CALLS.append(("laplace", np.array(matrix), np.array(rows), np.array(cols), v.copy()))
        return v

    perm.permanent = permanent
    perm.permanent_laplace = permanent_laplace
    sys.modules["piquasso._math.permanent"] = perm

    for name in ("piquasso._math.torontonian", "piquasso._math.pfaffian"):
        mod = types.ModuleType(name)

        def _missing(attr, _name=name):
            def f(*a, **k):
                raise RuntimeError("stubbed native module %s.%s" % (_name, attr))
            return f

        mod.__getattr__ = _missing
        sys.modules[name] = mod


def _load_test_module(relpath):
    path = os.path.join(REFERENCE, relpath)
    spec = importlib.util.spec_from_file_location("ref_" + os.path.basename(relpath)[:-3], path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _c(z):
    z = complex(z)
    return [z.real, z.imag]


def _mat(m):
    m = np.asarray(m)
    mc = m.astype(np.complex128)
    return {"shape": list(m.shape), "dtype": str(m.dtype),
            "re": mc.real.reshape(-1).tolist(), "im": mc.imag.reshape(-1).tolist()}


def _is_monkey(fn):
    return any(getattr(m, "name", "") == "monkey" for m in getattr(fn, "pytestmark", []))


def main():
    if not oracle.ref_available():
        raise SystemExit("oracle/_ref could not be built: no reference sources")
    _install_stubs()
    import piquasso as pq  # noqa: F401  (the reference package)
    from piquasso._simulators.passive import sampling as ref_sampling
    from piquasso._simulators.passive import simulation_steps as ref_steps

    # ---- 1. the reference's own permanent tests ------------------------------
    perm_cases = []
    tmod = _load_test_module("tests/_math/test_permanent.py")
    for name, fn in sorted(inspect.getmembers(tmod, inspect.isfunction)):
        if not name.startswith("test_") or _is_monkey(fn):
            continue
        if inspect.signature(fn).parameters:
            continue
        start = len(CALLS)
        fn()  # asserts the reference's golden value against oracle/_ref
        for kind, m, r, c, v in CALLS[start:]:
            assert kind == "permanent"
            perm_cases.append({"source": "tests/_math/test_permanent.py::" + name,
                               "matrix": _mat(m), "rows": r.tolist(), "cols": c.tolist(),
                               "value": _c(v)})
    print("reference permanent tests replayed:", len(perm_cases), "calls")

    # ---- 2. detection-probability goldens (connector.permanent callers) ------
    for rel, names in (
        ("tests/_simulators/passive/test_preparations.py", None),
        ("tests/_simulators/passive/test_state.py", None),
    ):
        try:
            mod = _load_test_module(rel)
        except Exception as exc:  # a test module needing an absent extra
            print("skip", rel, type(exc).__name__, exc)
            continue
        for name, fn in sorted(inspect.getmembers(mod, inspect.isfunction)):
            if not name.startswith("test_") or _is_monkey(fn):
                continue
            if inspect.signature(fn).parameters:
                continue
            start = len(CALLS)
            try:
                fn()
            except Exception as exc:
                print("  (not replayed) %s::%s: %s" % (rel, name, type(exc).__name__))
                del CALLS[start:]
                continue
            for kind, m, r, c, v in CALLS[start:]:
                if kind != "permanent" or m.size > 400:
                    continue
                perm_cases.append({"source": rel + "::" + name, "matrix": _mat(m),
                                   "rows": r.tolist(), "cols": c.tolist(), "value": _c(v)})
    print("permanent cases incl. probability goldens:", len(perm_cases))

    # ---- 3. seeded sampler goldens ---------------------------------------------
    sampler_cases = []
    laplace_cases = []
    orig_generate = ref_sampling.generate_samples

    def recording_generate(input, shots, calculate_permanent_laplace, interferometer,
                           reject_condition, postselect_data, uniform_particle_overlap,
                           config):
        rejects = []

        def logged_reject():
            r = bool(reject_condition())
            rejects.append(r)
            return r

        start = len(CALLS)
        samples = orig_generate(input, shots, calculate_permanent_laplace, interferometer,
                                logged_reject, postselect_data, uniform_particle_overlap,
                                config)
        if len(postselect_data[0]) == 0 and uniform_particle_overlap is None:
            sampler_cases.append({
                "input": [int(x) for x in input], "shots": int(shots),
                "interferometer": _mat(interferometer),
                "seed_sequence": int(config.seed_sequence),
                "rejects": rejects,
                "samples": [[int(x) for x in s] for s in samples],
                "source": CURRENT[0],
            })
            for kind, m, r, c, v in CALLS[start:][:40]:
                if kind == "laplace":
                    laplace_cases.append({"source": CURRENT[0], "matrix": _mat(m),
                                          "rows": r.tolist(), "cols": c.tolist(),
                                          "value": [_c(z) for z in v]})
        return samples

    ref_sampling.generate_samples = recording_generate
    ref_steps.generate_samples = recording_generate
    CURRENT = [""]
    mmod = _load_test_module("tests/_simulators/passive/test_measurements.py")
    for name, args in (
        ("test_boson_sampling_seeded", ()),
        ("test_boson_sampling_seeded_use_dask", (False,)),
        ("test_LossyInterferometer_boson_sampling_seeded", ()),
        ("test_LossyInterferometer_boson_sampling_uniform_losses", ()),
        ("test_uniform_loss", ()),
        ("test_general_loss", ()),
        ("test_mach_zehnder", ()),
        ("test_fourier", ()),
    ):
        CURRENT[0] = "tests/_simulators/passive/test_measurements.py::" + name
        getattr(mmod, name)(*args)  # asserts the reference's golden samples
    ref_sampling.generate_samples = orig_generate
    ref_steps.generate_samples = orig_generate
    print("sampler goldens:", len(sampler_cases), "laplace calls kept:", len(laplace_cases))


    # ---- 3b. the other per-shot algorithms of the sampler ----------
    # post-selection, uniform particle overlap, both, non-uniform losses: the
    # reference's generate_samples / generate_lossy_samples called directly with
    # seeded per-shot generators; reject conditions that draw from a shared
    # generator are described by (seed, transmission) so that a test can rebuild
    # them.
    from types import SimpleNamespace
    from scipy.stats import unitary_group as _ug
    variant_cases = []

    def run_variant(label, input, shots, U, seed, postselect=((), (), 1000), overlap=None,
                    loss=None, lossy_dilation=False):
        config = SimpleNamespace(seed_sequence=seed, use_dask=False)
        if loss is None:
            reject = lambda: False  # noqa: E731
        else:
            shared = np.random.default_rng(loss[0])
            reject = lambda: shared.uniform() > loss[1]  # noqa: E731
        if lossy_dilation:
            samples = ref_sampling.generate_lossy_samples(
                np.array(input), shots, oracle.ref_permanent_laplace, U, postselect, config)
        else:
            samples = orig_generate(np.array(input), shots, oracle.ref_permanent_laplace, U,
                                    reject, postselect, overlap, config)
        variant_cases.append({
            "label": label, "input": [int(x) for x in input], "shots": shots,
            "interferometer": _mat(U), "seed_sequence": seed,
            "postselect_modes": [int(x) for x in postselect[0]],
            "postselect_photons": [int(x) for x in postselect[1]],
            "max_trials": int(postselect[2]), "overlap": overlap,
            "loss": list(loss) if loss else None, "lossy_dilation": lossy_dilation,
            "samples": [[int(x) for x in smp] for smp in samples]})

    U6 = _ug.rvs(6, random_state=606)
    U5 = _ug.rvs(5, random_state=505)
    lossy5 = U5 @ np.diag(np.sqrt([0.9, 0.5, 0.7, 0.95, 0.3])) @ _ug.rvs(5, random_state=506)
    run_variant("postselect one mode", [1, 1, 1, 1, 0, 0], 40, U6, 11,
                postselect=((2,), (1,), 1000))
    run_variant("postselect two modes, bunched input", [2, 0, 1, 1, 0, 0], 40, U6, 12,
                postselect=((0, 4), (1, 0), 1000))
    run_variant("uniform overlap", [1, 1, 2, 0, 1, 0], 40, U6, 13, overlap=0.6)
    run_variant("uniform overlap + uniform loss", [1, 1, 2, 0, 1, 0], 40, 0.8 ** 0.5 * U6, 14,
                overlap=0.35, loss=(99, 0.8))
    # (with an exactly unitary matrix the reference's loss weight 1 - sum|u|^2 comes
    # out as -1e-16 and numpy rejects the weights: the combination needs losses)
    run_variant("postselect + uniform overlap + uniform loss", [1, 1, 1, 1, 0, 0], 30,
                0.9 ** 0.5 * U6, 15, postselect=((1,), (1,), 1000), overlap=0.5, loss=(5, 0.9))
    run_variant("postselect + uniform loss (sequential reject order)", [1, 1, 1, 1, 0, 0], 30,
                0.9 ** 0.5 * U6, 16, postselect=((3,), (1,), 1000), loss=(7, 0.9))
    run_variant("non-uniform losses (2d dilation)", [1, 1, 1, 0, 0], 40, lossy5, 17,
                lossy_dilation=True)
    run_variant("non-uniform losses + postselect", [1, 1, 1, 0, 0], 30, lossy5, 18,
                postselect=((0,), (1,), 1000), lossy_dilation=True)
    print("sampler variant goldens:", len(variant_cases))


    # ---- 3c. sampler-level drop-in (piquasso_b200.integration) --------------------
    # End-to-end check made HERE, where the reference is importable: whole
    # PassiveSimulator programs (ideal, uniformly lossy, post-selected, non-uniformly
    # lossy) give the same Result.samples with the reference's own sampler and with
    # ours patched in (pmf rows from the oracle: the host logic is what is checked).
    sys.path.insert(0, os.path.join(REPO, "tests"))
    from conftest import oracle_pmf_rows
    from piquasso_b200 import integration

    def programs():
        U = _ug.rvs(5, random_state=55)
        with pq.Program() as ideal:
            pq.Q(all) | pq.NumberState([2, 1, 1, 0, 1])
            pq.Q(all) | pq.Interferometer(U)
            pq.Q(all) | pq.ParticleNumberMeasurement()
        with pq.Program() as uniform_loss:
            pq.Q(all) | pq.NumberState([1, 1, 1, 1, 0])
            pq.Q(all) | pq.Interferometer(U)
            for i in range(5):
                pq.Q(i) | pq.Loss(transmissivity=0.9)
            pq.Q(all) | pq.ParticleNumberMeasurement()
        with pq.Program() as general_loss:
            pq.Q(all) | pq.NumberState([1, 1, 1, 0, 0])
            pq.Q(all) | pq.Interferometer(U)
            pq.Q(0) | pq.Loss(transmissivity=0.4)
            pq.Q(1) | pq.Loss(transmissivity=0.5)
            pq.Q(all) | pq.ParticleNumberMeasurement()
        with pq.Program() as postselected:
            pq.Q(all) | pq.NumberState([1, 1, 1, 1, 0])
            pq.Q(all) | pq.Interferometer(U)
            pq.Q(2) | pq.PostSelectPhotons(photon_counts=(1,))
            pq.Q(all) | pq.ParticleNumberMeasurement()
        return {"ideal": ideal, "uniform loss": uniform_loss, "general loss": general_loss,
                "postselected": postselected}

    def run_all():
        out = {}
        for label, program in programs().items():
            simulator = pq.PassiveSimulator(d=5, config=pq.Config(seed_sequence=77))
            out[label] = simulator.execute(program, shots=25).samples
        return out

    stock = run_all()
    with integration.install(pmf_rows=oracle_pmf_rows):
        patched = run_all()
    for label in stock:
        assert stock[label] == patched[label], ("sampler drop-in differs", label)
    print("sampler-level drop-in: identical Result.samples for", sorted(stock))

    # ---- 4. seeded Haar cases against the compiled reference ---------------------
    from scipy.stats import unitary_group
    rng = np.random.default_rng(2024)
    haar_cases = []
    for n in (2, 3, 4, 6, 8, 10, 12, 14, 16, 18, 20):
        U = unitary_group.rvs(n, random_state=n)
        ones = np.ones(n, dtype=int)
        haar_cases.append({"source": "haar n=%d seed=%d all-ones" % (n, n), "matrix": _mat(U),
                           "rows": ones.tolist(), "cols": ones.tolist(),
                           "value": _c(oracle.ref_permanent(U, ones, ones))})
    for trial in range(40):
        d = int(rng.integers(2, 11))
        nph = int(rng.integers(1, 13))
        rows = rng.multinomial(nph, np.ones(d) / d)
        cols = rng.multinomial(nph, np.ones(d) / d)
        U = unitary_group.rvs(d, random_state=1000 + trial)
        haar_cases.append({"source": "haar d=%d seed=%d multiplicities" % (d, 1000 + trial),
                           "matrix": _mat(U), "rows": rows.tolist(), "cols": cols.tolist(),
                           "value": _c(oracle.ref_permanent(U, rows, cols))})
    for trial in range(6):  # rectangular d1 x d2 (tests/_math/test_permanent.py:304-320 shape)
        d1, d2 = int(rng.integers(2, 7)), int(rng.integers(2, 7))
        nph = int(rng.integers(1, 8))
        rows = rng.multinomial(nph, np.ones(d1) / d1)
        cols = rng.multinomial(nph, np.ones(d2) / d2)
        A = rng.normal(size=(d1, d2)) + 1j * rng.normal(size=(d1, d2))
        haar_cases.append({"source": "rectangular %dx%d trial %d" % (d1, d2, trial),
                           "matrix": _mat(A), "rows": rows.tolist(), "cols": cols.tolist(),
                           "value": _c(oracle.ref_permanent(A, rows, cols))})
    # config 3 of BASELINE.json: 60 modes, 24 photons, unfiltered d x d call
    U60 = unitary_group.rvs(60, random_state=60)
    r60 = np.random.default_rng(3)
    for label, out_occ, in_occ in (
        ("multinomial", r60.multinomial(24, np.ones(60) / 60), r60.multinomial(24, np.ones(60) / 60)),
    ):
        haar_cases.append({"source": "cfg3 60 modes 24 photons " + label, "matrix": _mat(U60),
                           "rows": [int(x) for x in out_occ], "cols": [int(x) for x in in_occ],
                           "value": _c(oracle.ref_permanent(U60, out_occ, in_occ))})
    for trial in range(30):
        d = int(rng.integers(2, 10))
        k = int(rng.integers(1, 10))
        rows = rng.multinomial(k - 1, np.ones(d) / d) if k > 1 else np.zeros(d, dtype=int)
        cols = rng.multinomial(k, np.ones(d) / d)
        U = unitary_group.rvs(d, random_state=2000 + trial)
        if trial % 2 == 0:  # sampler style (zero-filtered, sampling.py:711-734)
            U = U[np.ix_(rows > 0, cols > 0)]
            rows = rows[rows > 0]
            cols = cols[cols > 0]
        v = oracle.ref_permanent_laplace(U, rows, cols)
        laplace_cases.append({"source": "haar laplace trial %d" % trial, "matrix": _mat(U),
                              "rows": [int(x) for x in rows], "cols": [int(x) for x in cols],
                              "value": [_c(z) for z in v]})

    # ---- 5. Gray-code traces from the reference counter ---------------------------
    gray_cases = []
    for limits in ([2, 2, 2, 2], [3, 2, 3], [1, 2, 2, 2, 2], [4, 1, 3, 2], [2, 5, 1, 3, 2, 2]):
        total = int(np.prod(limits))
        for off in sorted({0, 1, total // 3, total // 2, total - 1}):
            g0, trace = oracle.ref_gray_trace(limits, off, min(total - 1 - off, 64))
            gray_cases.append({"limits": limits, "offset": off, "gray0": g0.tolist(),
                               "trace": [list(t) for t in trace]})

    out = {
        "permanent_reference_tests.json": perm_cases,
        "permanent_haar.json": haar_cases,
        "laplace.json": laplace_cases,
        "sampler.json": sampler_cases,
        "sampler_variants.json": variant_cases,
        "gray.json": gray_cases,
    }
    for fname, payload in out.items():
        with open(os.path.join(HERE, fname), "w") as fh:
            json.dump(payload, fh)
        print("wrote", fname, len(payload), "cases,",
              os.path.getsize(os.path.join(HERE, fname)) // 1024, "KiB")


if __name__ == "__main__":
    main()
