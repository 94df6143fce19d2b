"""Development probe 3: time binary-kernel variants (choice = 2 + 10*B + 100*CH)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.stats import unitary_group
from piquasso_b200 import _lib
from piquasso_b200._math.permanent import permanent

lib = _lib.load()
peak = lib.pq_fp64_peak_tflops(0, 1 << 17)
print("fp64 peak", peak, flush=True)
variants = [int(v) for v in sys.argv[2].split(",")]
for n in [int(x) for x in sys.argv[1].split(",")]:
    U = unitary_group.rvs(n, random_state=n); ones = np.ones(n, dtype=np.int32)
    ref = None
    for v in variants:
        lib.pq_set_kernel_choice(v)
        permanent(U, ones, ones)
        best = 1e30
        for _ in range(2):
            val = complex(permanent(U, ones, ones)); best = min(best, lib.pq_last_kernel_ms(0))
        if ref is None: ref = val
        terms = 2.0 ** (n - 1)
        tf = (8 * n + 2) * terms / (best * 1e-3) / 1e12
        print(f"n={n} variant={v}: {best:.3f} ms {terms/best/1e6:.2f} Gterms/s {tf:.2f} TF ({tf/peak*100:.1f}% of DFMA peak) dev={abs(val-ref)/abs(ref):.1e}", flush=True)
