"""Config 3 through the batched-permanent entry with ONE problem (dev helper)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.stats import unitary_group
from piquasso_b200 import _lib
from piquasso_b200._math.permanent import permanent
from piquasso_b200.sampling import permanent_batch
lib = _lib.load()
U60 = unitary_group.rvs(60, random_state=60)
r60 = np.random.default_rng(3)
cases = {"multinomial": (r60.multinomial(24, np.ones(60) / 60), r60.multinomial(24, np.ones(60) / 60))}
hard = np.array([1] * 16 + [2] * 4 + [0] * 40); heavy = np.array([2] * 12 + [0] * 48)
cases["hard_16ones_4twos"] = (hard, hard); cases["heavy_12twos"] = (heavy, heavy)
for name, (r, c) in cases.items():
    r = r.astype(np.int32); c = c.astype(np.int32)
    v = complex(permanent(U60, r, c))
    ts = []
    for _ in range(30):
        t = time.perf_counter(); permanent(U60, r, c); ts.append(time.perf_counter() - t)
    k1 = lib.pq_last_kernel_ms(0)
    vb = complex(permanent_batch(U60, r[None, :], c[None, :])[0])
    tb = []
    for _ in range(30):
        t = time.perf_counter(); permanent_batch(U60, r[None, :], c[None, :]); tb.append(time.perf_counter() - t)
    k2 = lib.pq_last_kernel_ms(0)
    print("%-20s single: wall %.1f us kernel %.1f us | batch of one: wall %.1f us kernel %.1f us | rel diff %.1e"
          % (name, 1e6 * np.median(ts), 1e3 * k1, 1e6 * np.median(tb), 1e3 * k2, abs(v - vb) / abs(v)), flush=True)
