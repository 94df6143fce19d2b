"""Development probe 4: accuracy of the kernel variants (Haar vs long-double oracle,
rank-1 vs closed form) and their speed."""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.stats import unitary_group
import oracle
from piquasso_b200 import _lib
from piquasso_b200._math.permanent import permanent

lib = _lib.load()
peak = lib.pq_fp64_peak_tflops(0, 1 << 17)
print("fp64 peak", peak, flush=True)
variants = [1, 22, 32, 42]

def run(U, ones, v):
    lib.pq_set_kernel_choice(v)
    try:
        val = complex(permanent(U, ones, ones))
    except Exception as e:
        return None, None
    return val, lib.pq_last_kernel_ms(0)

for n in [int(x) for x in sys.argv[1].split(",")]:
    U = unitary_group.rvs(n, random_state=n); ones = np.ones(n, dtype=np.int32)
    t = time.time(); want = oracle.permanent(U, ones, ones, precision=1, njobs=512); dt = time.time() - t
    dbl = oracle.permanent(U, ones, ones, precision=0, njobs=64)
    print(f"haar n={n}: oracle long double {dt:.1f}s; oracle double(64 jobs) relerr={abs(dbl-want)/abs(want):.1e}", flush=True)
    for v in variants:
        val, ms = run(U, ones, v)
        if val is None: continue
        print(f"   variant={v}: relerr={abs(val-want)/abs(want):.2e}  {ms:.3f} ms", flush=True)
for n in [int(x) for x in sys.argv[2].split(",")]:
    rng = np.random.default_rng(0)
    u = np.exp(2j * np.pi * rng.random(n)); w = np.exp(2j * np.pi * rng.random(n))
    A = np.outer(u, w); ones = np.ones(n, dtype=np.int32)
    exact = math.factorial(n) * np.prod(u) * np.prod(w)
    for v in variants:
        val, ms = run(A, ones, v)
        if val is None: continue
        terms = 2.0 ** (n - 1); tf = (8 * n + 2) * terms / (ms * 1e-3) / 1e12
        print(f"rank1 n={n} variant={v}: relerr={abs(val-exact)/abs(exact):.2e}  {ms:.2f} ms {tf:.2f} TF ({tf/peak*100:.1f}%)", flush=True)
    J = np.ones((n, n), dtype=complex)
    val, ms = run(J, ones, 0)
    print(f"all-ones n={n} auto: relerr={abs(val-math.factorial(n))/math.factorial(n):.2e}", flush=True)
