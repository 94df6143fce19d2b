"""Throughput of pq_perm_batch_c128 (detection-probability batches):
python tools/batch_probe.py -> kernel ms and algorithmic TFLOP/s per batch shape."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.stats import unitary_group
from piquasso_b200 import _lib
from piquasso_b200.sampling import permanent_batch
from piquasso_b200._math.permanent import permanent

lib = _lib.load()
peak = lib.pq_fp64_peak_tflops(0, 20000) if hasattr(lib, "pq_fp64_peak_tflops") else 36.7
m = 60
U = unitary_group.rvs(m, random_state=m)
rng = np.random.default_rng(7)
shapes = [(8, 50000), (12, 20000), (16, 10000), (20, 4000), (24, 400)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in a.split(":")) for a in sys.argv[1:]]
for n, B in shapes:
    inp = np.array([1] * n + [0] * (m - n), np.int32)
    outs = rng.multinomial(n, np.ones(m) / m, size=B).astype(np.int32)
    flops = 0.0; terms = 0.0
    for r in outs:
        nz = r[r > 0].copy()
        nz[np.argmin(nz)] -= 1
        t = float(np.prod(nz + 1.0))
        terms += t; flops += t * (2 * n + 6 * n + 2)
    best = 1e30; wall = 1e30
    for _ in range(4):
        t0 = time.perf_counter(); v = permanent_batch(U, outs, inp); wall = min(wall, time.perf_counter() - t0)
        best = min(best, lib.pq_last_kernel_ms(0))
    err = max(abs(v[i] - complex(permanent(U, outs[i], inp))) / max(abs(v[i]), 1e-300) for i in range(0, B, max(1, B // 16)))
    # general flavour: the same photons through input modes with multiplicities
    inp2 = np.zeros(m, np.int32); inp2[: n // 2] = 2; inp2[n // 2] = n - 2 * (n // 2)
    v2 = permanent_batch(U, outs[:64], inp2)
    err2 = max(abs(v2[i] - complex(permanent(U, outs[i], inp2))) / max(abs(v2[i]), 1e-300) for i in range(0, 64, 4))
    print("    max relative difference to single permanent() calls: unit columns %.2e, column multiplicities %.2e" % (err, err2))
    print("n=%2d B=%6d terms %.3e  kernel %.3f ms  wall %.3f ms  %.2f alg TFLOP/s (%.1f%% of %.1f)  %.2f ps/term"
          % (n, B, terms, best, wall * 1e3, flops / best / 1e9, 100 * flops / best / 1e9 / peak, peak,
             best * 1e9 / terms), flush=True)
