"""python tools/run_laplace.py K [BATCH]: one permanent_laplace of the sampler shape (dev helper)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.stats import unitary_group
from piquasso_b200 import _lib
from piquasso_b200._math.permanent import permanent_laplace
from piquasso_b200.sampling import permanent_laplace_batch
k = int(sys.argv[1]); batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1
lib = _lib.load()
a = np.ascontiguousarray(unitary_group.rvs(30, random_state=k)[: k - 1, :k]); r = np.ones(k - 1, np.int32); c = np.ones(k, np.int32)
for _ in range(2):
    if batch == 1:
        permanent_laplace(a, r, c)
    else:
        permanent_laplace_batch([a] * batch, [r] * batch, [c] * batch)
    print(k, batch, "%.3f ms" % lib.pq_last_kernel_ms(0), flush=True)
