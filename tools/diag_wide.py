"""Development probe: wide (S=32) and S=4 lane-split walks on n-ary problems vs the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from piquasso_b200._math.permanent import permanent, permanent_laplace
from piquasso_b200.sampling import permanent_batch
rng = np.random.default_rng(1)
def rel(a, b): return abs(a - b) / abs(b)
for rows in ([65], [33, 32], [60, 5], [40, 35], [2, 63], [1, 64], [30, 30, 5], [33, 31], [40, 24], [20, 20, 24], [3, 61], [10, 9], [12, 8, 4]):
    rows = np.array(rows); n = int(rows.sum())
    a = (rng.normal(size=(len(rows), n)) + 1j * rng.normal(size=(len(rows), n))) / 2
    cols = np.ones(n, int)
    want = oracle.permanent(a, rows, cols, precision=1)
    ref = rel(oracle.permanent(a, rows, cols), want)
    got1 = complex(permanent(a, rows, cols))
    got2 = complex(permanent_batch(a, rows[None, :], cols)[0])
    print(rows.tolist(), "n=%d" % n, "ref_err %.1e  permanent %.2e  batch(lane-split) %.2e" % (ref, rel(got1, want), rel(got2, want)), flush=True)
for nc, rows in ((70, [23, 23, 23]), (60, [20, 20, 19]), (66, [65])):
    rows = np.array(rows)
    a = (rng.normal(size=(len(rows), nc)) + 1j * rng.normal(size=(len(rows), nc))) / 3
    cols = np.ones(nc, int)
    want = oracle.permanent_laplace(a, rows, cols, precision=1)
    ref = np.max(np.abs(oracle.permanent_laplace(a, rows, cols) - want) / np.abs(want))
    got = permanent_laplace(a, rows, cols)
    print("laplace", nc, rows.tolist(), "ref_err %.1e  gpu %.2e" % (ref, np.max(np.abs(got - want) / np.abs(want))), flush=True)
