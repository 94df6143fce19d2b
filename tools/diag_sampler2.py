"""One fresh process: warm-up (2 shots), then the 10^4-shot config-4 run with every timer
(dev helper for the sporadic slow first call): python tools/diag_sampler2.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
t00 = time.perf_counter()
import numpy as np
from scipy.stats import unitary_group
from piquasso_b200 import _lib, sampling
d, n, shots = 100, 25, 10000
U = unitary_group.rvs(d, random_state=d)
inp = np.array([1] * n + [0] * (d - n))
lib = _lib.load()
t0 = time.perf_counter(); sampling.generate_samples(inp, 2, U, 123); tw = time.perf_counter() - t0
out = []
for rep in range(2):
    sampling.TIMERS.clear()
    t0 = time.perf_counter(); s = sampling.generate_samples(inp, shots, U, 123); dt = time.perf_counter() - t0
    tm = sampling.TIMERS
    out.append("run%d %.3f s [gen %.3f grow %.3f plan %.3f dev %.3f kern %.3f | growth %.3f stage %.3f enq %.3f wait %.3f]" % (
        rep, dt, tm.get("host: per-shot generators", 0), tm.get("host: grow input", 0),
        tm.get("  of which planning (host threads)", 0), tm.get("  of which device phase", 0),
        tm.get("  of which GPU kernels (CUDA events)", 0), tm.get("    device phase: scratch growth", 0),
        tm.get("    device phase: staging descriptors", 0), tm.get("    device phase: enqueue", 0),
        tm.get("    device phase: waiting for the stream", 0)))
print("import+setup %.2f s, warm-up %.3f s; " % (t0 - t00, tw) + " ; ".join(out), flush=True)
