"""Config 3 (60 modes / 24 photons, multiplicities) and neighbours: GPU wall / kernel time only."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.stats import unitary_group
from piquasso_b200 import _lib
from piquasso_b200._math.permanent import permanent
lib = _lib.load()
U60 = unitary_group.rvs(60, random_state=60)
r60 = np.random.default_rng(3)
cases = {"multinomial": (r60.multinomial(24, np.ones(60) / 60), r60.multinomial(24, np.ones(60) / 60))}
hard = np.array([1] * 16 + [2] * 4 + [0] * 40); heavy = np.array([2] * 12 + [0] * 48)
cases["hard_16ones_4twos"] = (hard, hard); cases["heavy_12twos"] = (heavy, heavy)
ones20 = np.array([1] * 20 + [0] * 40); twos = np.array([2] * 10 + [0] * 50)
cases["rows_unit_cols_2x10"] = (ones20, twos); cases["rows_2x10_cols_unit"] = (twos, ones20)
for name, (r, c) in cases.items():
    r = r.astype(np.int32); c = c.astype(np.int32)
    v = complex(permanent(U60, r, c))
    ts = []
    for _ in range(50):
        t = time.perf_counter(); permanent(U60, r, c); ts.append(time.perf_counter() - t)
    print("%-22s wall %.4f ms  kernel %.4f ms  value %.6e%+.6ej" % (name, 1e3 * np.median(ts), lib.pq_last_kernel_ms(0), v.real, v.imag), flush=True)
