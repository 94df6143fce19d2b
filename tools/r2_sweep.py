"""Segment-length sweep of the two permanent walks on small Haar problems (dev helper):
python tools/r2_sweep.py -> kernel microseconds per (n, kernel choice, segment length)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.stats import unitary_group
from piquasso_b200 import _lib, plan as pqplan
from piquasso_b200._math.permanent import permanent
lib = _lib.load()
for n in (16, 18, 20, 22, 24, 25, 26):
    u = unitary_group.rvs(n, random_state=n); ones = np.ones(n, np.int32)
    for choice in (1, 2):
        row = []
        for hint in (8, 16, 32, 64, 128, 256, 512, 1024):
            lib.pq_set_kernel_choice(choice); lib.pq_set_seg_len_hint(hint)
            W = pqplan.plan(ones, ones)["seg_len"]
            best = 1e9
            for _ in range(12):
                permanent(u, ones, ones)
                best = min(best, lib.pq_last_kernel_ms(0) * 1e3)
            row.append("W%d:%.1f" % (W, best))
        print(n, "choice", choice, " ".join(row), flush=True)
for n in (28, 30, 32):
    u = unitary_group.rvs(n, random_state=n); ones = np.ones(n, np.int32)
    row = []
    for hint in (256, 512, 1024, 2048, 4096, 8192, 16384):
        lib.pq_set_kernel_choice(2); lib.pq_set_seg_len_hint(hint)
        W = pqplan.plan(ones, ones)["seg_len"]
        best = 1e9
        for _ in range(4):
            permanent(u, ones, ones)
            best = min(best, lib.pq_last_kernel_ms(0) * 1e3)
        row.append("W%d:%.1f" % (W, best))
    print(n, "binary", " ".join(row), flush=True)
lib.pq_set_kernel_choice(0); lib.pq_set_seg_len_hint(0)
for n in (16, 18, 20, 22, 24, 25, 26, 28, 30, 32):
    ones = np.ones(n, np.int32); u = unitary_group.rvs(n, random_state=n)
    p = pqplan.plan(ones, ones)
    best = 1e9
    for _ in range(6):
        permanent(u, ones, ones)
        best = min(best, lib.pq_last_kernel_ms(0) * 1e3)
    print(n, "auto: kernel", p["kernel"], "W", p["seg_len"], "%.1f us" % best, "%.2f TFLOP/s" % (2.0 ** (n - 1) * (8 * n + 2) / best / 1e6), flush=True)
