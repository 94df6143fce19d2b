"""Config 3 through the two walks that can compute it (dev helper): the generic n-ary walk
(pq_perm_c128) and the lane-split batch walk in permanent-only mode (pq_perm_batch_c128)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.stats import unitary_group
import bench_secondary
from piquasso_b200 import _lib
from piquasso_b200._math.permanent import permanent
from piquasso_b200.sampling import permanent_batch
lib = _lib.load()
U60 = unitary_group.rvs(60, random_state=60)
for name, (rows, cols) in bench_secondary.cfg3_cases().items():
    rows = rows.astype(np.int32); cols = cols.astype(np.int32)
    for label, fn in (("generic walk", lambda: complex(permanent(U60, rows, cols))),
                      ("batch walk  ", lambda: complex(permanent_batch(U60, rows[None, :], cols[None, :])[0]))):
        for _ in range(3): v = fn()
        ts = []; ks = []
        for _ in range(30):
            t = time.perf_counter(); v = fn(); ts.append(time.perf_counter() - t); ks.append(lib.pq_last_kernel_ms(0))
        print("%-18s %s wall %.1f us kernel %.1f us value %.12e%+.12ej" % (name, label, np.median(ts) * 1e6, np.median(ks) * 1e3, v.real, v.imag), flush=True)
