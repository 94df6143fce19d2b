import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np
from scipy.stats import unitary_group
from piquasso_b200 import _lib, sampling
from piquasso_b200._math.permanent import permanent_laplace
import oracle
lib = _lib.load()
for k in range(2, 14):
    a = np.ascontiguousarray(unitary_group.rvs(30, random_state=k)[: k - 1, :k]); r = np.ones(k - 1, np.int32); c = np.ones(k, np.int32)
    try:
        got = permanent_laplace(a, r, c)
        want = oracle.permanent_laplace(a, r, c, precision=1)
        print(k, 'batch ok', np.max(np.abs(got-want)/np.abs(want)))
    except Exception as e:
        print(k, 'batch FAIL', e)
U = unitary_group.rvs(12, random_state=3)
for k in range(1, 8):
    inp = np.array([1]*k + [0]*(12-k))
    try:
        s = sampling.generate_samples(inp, 3, U, 5)
        print(k, 'sampler ok', s[0])
    except Exception as e:
        print(k, 'sampler FAIL', e)
