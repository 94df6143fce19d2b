import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from scipy.stats import unitary_group
from piquasso_b200 import _lib, sampling
lib = _lib.load()
u60 = unitary_group.rvs(60, random_state=60)
inp = np.array([1] * 20 + [0] * 40, dtype=np.int32)
outs = np.random.default_rng(7).multinomial(20, np.ones(60) / 60, size=2000).astype(np.int32)
sampling.detection_probabilities(u60, inp, outs)
ts = []
for _ in range(7):
    t = time.perf_counter(); p = sampling.detection_probabilities(u60, inp, outs); ts.append(time.perf_counter() - t)
print("f2: 2000 outputs x 20 photons: wall %.2f ms (min %.2f), kernel %.2f ms" % (1e3 * np.median(ts), 1e3 * min(ts), lib.pq_last_kernel_ms(0)))
