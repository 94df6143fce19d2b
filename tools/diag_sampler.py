import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from scipy.stats import unitary_group
from piquasso_b200 import _lib, sampling
from piquasso_b200._math.permanent import permanent
mode = sys.argv[1]
U = unitary_group.rvs(100, random_state=100); inp = np.array([1] * 25 + [0] * 75)
if mode in ("omp", "both"):
    import oracle
    V = unitary_group.rvs(24, random_state=24); ones = np.ones(24, np.int32)
    oracle.ref_permanent(V, ones, ones)
if mode in ("perm", "both"):
    V = unitary_group.rvs(24, random_state=24); ones = np.ones(24, np.int32)
    for _ in range(20): permanent(V, ones, ones)
sampling.generate_samples(inp, 2, U, 123)
for rep in range(2):
    sampling.TIMERS.clear()
    t = time.perf_counter(); sampling.generate_samples(inp, 10000, U, 123); dt = time.perf_counter() - t
    print(mode, os.environ.get("PQ_PLAN_THREADS"), "rep", rep, "%.2f s" % dt, {k.strip()[:12]: round(v, 3) for k, v in sampling.TIMERS.items()}, flush=True)
