"""Test configuration.

* ``-m "not gpu"``: oracle vs the golden vectors, host logic (planning, Gray
  enumeration, overload resolution, error paths), the C ABI's symbol table and
  the two-rank gloo path.  No test in this set launches a kernel.
* ``-m gpu``: the parity tests proper; every one goes through libpqperm.so.

/root/reference is never read by a test: the fixtures in tests/golden/ were
produced from it by tests/golden/make_golden.py.
"""

import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as fh:
        return json.load(fh)


def golden_matrix(entry):
    shape = entry["shape"]
    m = np.array(entry["re"], dtype=np.float64) + 1j * np.array(entry["im"], dtype=np.float64)
    m = m.reshape(shape).astype(np.complex128)
    if entry.get("dtype") == "complex64":
        m = m.astype(np.complex64)
    elif entry.get("dtype", "").startswith("float"):
        m = m.real.astype(entry["dtype"])
    return m


def golden_complex(v):
    return complex(v[0], v[1])


def haar(n, seed):
    from scipy.stats import unitary_group
    return unitary_group.rvs(n, random_state=seed) if n > 1 else np.array([[np.exp(0.3j)]])


def relerr(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def oracle_pmf_rows(interferometer, out_occ, in_occ):
    """Unnormalised pmf rows of the sampler's photon step computed with the
    ORACLE's permanent_laplace (the reference's _calculate_pmf,
    piquasso/_simulators/passive/sampling.py:723-749): the CPU stand-in for
    piquasso_b200.sampling.sampler_pmf when only the host logic is under test."""
    import oracle
    u = np.asarray(interferometer, dtype=np.complex128)
    rows = []
    for out, inp in zip(np.atleast_2d(out_occ), np.atleast_2d(in_occ)):
        inz, onz = inp > 0, out > 0
        part = oracle.permanent_laplace(u[np.ix_(onz, inz)], out[onz], inp[inz])
        amp = np.zeros(u.shape[0], dtype=np.complex128)
        for j, col in enumerate(np.flatnonzero(inz)[: len(part)]):
            amp += inp[col] * part[j] * u[:, col]
        rows.append(np.abs(amp) ** 2)
    return np.array(rows)


def run_sampler_variant(case, pmf_rows=None):
    """Replay one case of tests/golden/sampler_variants.json through
    piquasso_b200.sampling (pmf_rows=None: the CUDA path)."""
    from piquasso_b200 import sampling
    u = golden_matrix(case["interferometer"])
    postselect = (tuple(case["postselect_modes"]), tuple(case["postselect_photons"]),
                  case["max_trials"])
    if case["lossy_dilation"]:
        return sampling.generate_lossy_samples(case["input"], case["shots"], u,
                                               case["seed_sequence"],
                                               postselect_data=postselect, pmf_rows=pmf_rows)
    reject = None
    if case["loss"]:
        shared = np.random.default_rng(case["loss"][0])
        transmission = case["loss"][1]
        reject = lambda: shared.uniform() > transmission  # noqa: E731
    return sampling.generate_samples(case["input"], case["shots"], u, case["seed_sequence"],
                                     reject_condition=reject, postselect_data=postselect,
                                     uniform_particle_overlap=case["overlap"],
                                     pmf_rows=pmf_rows)


def _ensure_built():
    """The .so files are git-ignored build products; a checkout that has not run
    __graft_entry__.build() yet gets them built once (nvcc / gcc, no GPU needed)."""
    from piquasso_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from piquasso_b200 import build
        build.build()
        build.build_pybind()


_ensure_built()


@pytest.fixture(scope="session")
def lib():
    from piquasso_b200 import _lib
    return _lib.load()
