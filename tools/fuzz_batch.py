"""Differential fuzz of pq_perm_batch_c128 (one-lane / general / hypercube / wide flavours, column
expansion, short segments, host-thread planning) against single permanent() calls, which run
through the independent perm_walk_* kernels.  python tools/fuzz_batch.py [CASES]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.stats import unitary_group
from piquasso_b200._math.permanent import permanent
from piquasso_b200.sampling import permanent_batch

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 120
rng = np.random.default_rng(2026)
worst = 0.0
worst_case = None
nprob = 0
for case in range(cases):
    m = int(rng.integers(3, 41))
    U = unitary_group.rvs(m, random_state=int(rng.integers(1, 10**6)))
    ph = int(rng.integers(1, min(26, 3 * m) + 1))
    # inputs: single photons, or multiplicities concentrated on a few modes
    if rng.random() < 0.5 and ph <= m:
        inp = np.zeros(m, np.int32); inp[rng.choice(m, ph, replace=False)] = 1
    else:
        k = int(rng.integers(1, m + 1)); inp = np.zeros(m, np.int32)
        inp[rng.choice(m, k, replace=False)] = rng.multinomial(ph, np.ones(k) / k)
    B = int(rng.integers(1, 600)) if case % 7 else int(rng.integers(600, 3000))
    k = int(rng.integers(1, m + 1))
    sup = rng.choice(m, k, replace=False)
    outs = np.zeros((B, m), np.int32)
    outs[:, sup] = rng.multinomial(ph, np.ones(k) / k, size=B)
    # keep the single-call side affordable: at most ~2^22 terms per problem
    terms = np.prod(outs.astype(np.float64) + 1.0, axis=1)
    keep = terms <= 2.0 ** 22
    outs = outs[keep][:400]
    if len(outs) == 0:
        continue
    got = permanent_batch(U, outs, inp)
    idx = rng.choice(len(outs), min(len(outs), 6), replace=False)
    for b in idx:
        want = complex(permanent(U, outs[b], inp))
        scale = max(abs(want), 1e-300)
        err = abs(got[b] - want) / scale
        # both sides are double walks of an ill-conditioned sum: compare to the size of the terms
        if err > 1e-7:
            print("MISMATCH case", case, "m", m, "photons", ph, "b", b, got[b], want, err,
                  "rows", outs[b][outs[b] > 0], "cols", inp[inp > 0], flush=True)
        if err > worst:
            worst, worst_case = err, (U, outs[b].copy(), inp.copy(), got[b], want)
        nprob += 1
print("fuzz_batch: %d cases, %d problems compared, worst relative difference %.2e" % (cases, nprob, worst))
if worst_case is not None:
    # who is off in the worst case?  both against the long-double oracle
    import oracle
    U_, r_, c_, gb, gs = worst_case
    ref = oracle.permanent(U_, r_, c_, precision=1)
    print("  worst case rows %s cols %s: batch %.2e, single %.2e from the long-double oracle"
          % (r_[r_ > 0], c_[c_ > 0], abs(gb - ref) / abs(ref), abs(gs - ref) / abs(ref)))

# one batch larger than the chunk the library plans and uploads at a time (2^17 problems)
m, ph, B = 16, 5, 300000
U = unitary_group.rvs(m, random_state=16)
inp = np.zeros(m, np.int32); inp[:ph] = 1
outs = rng.multinomial(ph, np.ones(m) / m, size=B).astype(np.int32)
got = permanent_batch(U, outs, inp)
w = 0.0
for b in list(rng.choice(B, 40, replace=False)) + [0, 131071, 131072, 262143, 262144, B - 1]:
    want = complex(permanent(U, outs[b], inp))
    w = max(w, abs(got[b] - want) / max(abs(want), 1e-300))
assert w < 1e-9, w
print("chunked batch of %d problems: worst relative difference %.2e" % (B, w))
