"""Run one permanent on the GPU: python tools/run_one.py N VARIANT [REPS] (dev helper)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.stats import unitary_group
from piquasso_b200 import _lib
from piquasso_b200._math.permanent import permanent

n = int(sys.argv[1]); variant = int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
lib = _lib.load()
lib.pq_set_kernel_choice(variant)
U = unitary_group.rvs(n, random_state=n)
ones = np.ones(n, dtype=np.int32)
for _ in range(reps):
    v = complex(permanent(U, ones, ones))
    ms = lib.pq_last_kernel_ms(0)
    print(n, variant, v, "%.3f ms" % ms, "%.2f Gterms/s" % (2.0 ** (n - 1) / ms / 1e6), flush=True)
