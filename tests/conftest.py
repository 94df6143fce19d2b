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
