"""57..64 active columns: generic walk (NC = 60 / 64 instantiations) vs the lane-split batch
walk in permanent-only mode (dev helper)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from piquasso_b200 import _lib
from piquasso_b200._math.permanent import permanent
from piquasso_b200.sampling import permanent_batch
lib = _lib.load()
rng = np.random.default_rng(5)
for rows in ([20, 18, 18], [21, 21, 22], [16, 16, 16, 12], [16, 16, 16, 16]):
    rows = np.array(rows, np.int32); n = int(rows.sum())
    a = (rng.normal(size=(len(rows), n)) + 1j * rng.normal(size=(len(rows), n))) / 3
    cols = np.ones(n, np.int32)
    for label, fn in (("generic", lambda: complex(permanent(a, rows, cols))),
                      ("batch  ", lambda: complex(permanent_batch(a, rows[None, :], cols[None, :])[0]))):
        for _ in range(2): v = fn()
        ks = []
        for _ in range(5):
            v = fn(); ks.append(lib.pq_last_kernel_ms(0))
        print(n, rows.tolist(), label, "kernel %.3f ms" % np.median(ks), v, flush=True)
