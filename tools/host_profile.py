"""PQ_HOST_PROFILE=1 python tools/host_profile.py [n]: host-side split of small permanents
through the pybind11 module (dev helper)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "piquasso_b200", "native"))
import numpy as np
from scipy.stats import unitary_group
import permanent as pyb
from piquasso_b200 import _lib
lib = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
u = unitary_group.rvs(n, random_state=n); ones = np.ones(n, np.int32)
for _ in range(50):
    pyb.permanent(u, ones, ones)
ts = []
for _ in range(2000):
    t = time.perf_counter(); pyb.permanent(u, ones, ones); ts.append(time.perf_counter() - t)
print("n=%d pybind wall: median %.2f us, min %.2f us; kernel %.2f us" % (n, np.median(ts) * 1e6, np.min(ts) * 1e6, lib.pq_last_kernel_ms(0) * 1e3))
lib.pq_set_timing(0)
ts = []
for _ in range(2000):
    t = time.perf_counter(); pyb.permanent(u, ones, ones); ts.append(time.perf_counter() - t)
print("n=%d pybind wall, timing events off: median %.2f us, min %.2f us" % (n, np.median(ts) * 1e6, np.min(ts) * 1e6))
lib.pq_set_timing(1)
