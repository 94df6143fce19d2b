"""Development probe 2 (run under gpurun): n-ary fix, dynamic scheduling timings, Laplace."""
import ctypes, json, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.stats import unitary_group
import oracle
from piquasso_b200 import _lib
from piquasso_b200._math.permanent import permanent, permanent_laplace

lib = _lib.load()
peak = lib.pq_fp64_peak_tflops(0, 1 << 17)
print("fp64 peak TFLOP/s", peak, flush=True)

def relerr(a, b):
    return abs(a - b) / max(abs(b), 1e-300)

bad = 0
rng = np.random.default_rng(7)
for trial in range(60):
    d = int(rng.integers(2, 9))
    nph = int(rng.integers(1, 10))
    rows = rng.multinomial(nph, np.ones(d) / d).astype(np.int32)
    cols = rng.multinomial(nph, np.ones(d) / d).astype(np.int32)
    U = unitary_group.rvs(d, random_state=trial + 100)
    want = oracle.permanent(U, rows, cols, precision=1)
    for hint in [0, 1, 6, 64]:
        lib.pq_set_seg_len_hint(hint)
        got = complex(permanent(U, rows, cols))
        e = relerr(got, want)
        ok = e < 1e-10 or abs(got - want) < 1e-13
        bad += (not ok)
        if not ok:
            print(f"BAD nary d={d} rows={rows.tolist()} cols={cols.tolist()} hint={hint} relerr={e:.2e}", flush=True)
lib.pq_set_seg_len_hint(0)
print("nary bad", bad, flush=True)

# Laplace
lbad = 0
for trial in range(80):
    d = int(rng.integers(2, 10))
    k = int(rng.integers(1, 9))
    rows = rng.multinomial(k - 1, np.ones(d) / d).astype(np.int32) if k > 1 else np.zeros(d, dtype=np.int32)
    cols = rng.multinomial(k, np.ones(d) / d).astype(np.int32)
    U = unitary_group.rvs(d, random_state=trial + 500)
    if trial % 3 == 0:  # sampler style: filtered
        U = U[np.ix_(rows > 0, cols > 0)]; rows = rows[rows > 0]; cols = cols[cols > 0]
    want = oracle.permanent_laplace(U, rows, cols, precision=1)
    got = permanent_laplace(U, rows, cols)
    ok = got.shape == want.shape and np.allclose(got, want, rtol=1e-10, atol=1e-13)
    lbad += (not ok)
    if not ok or trial < 4:
        print(f"laplace shape={U.shape} rows={rows.tolist()} cols={cols.tolist()} ok={ok}\n   got={got}\n  want={want}", flush=True)
# bigger Laplace: sampler shape, k columns unit, k-1 rows unit
for k in [10, 13, 16, 20, 24, 25]:
    U = unitary_group.rvs(max(k, 30), random_state=k)[: k - 1, :k]
    rows = np.ones(k - 1, dtype=np.int32); cols = np.ones(k, dtype=np.int32)
    t = time.perf_counter(); got = permanent_laplace(U, rows, cols); dt = time.perf_counter() - t
    ms = lib.pq_last_kernel_ms(0)
    if k <= 20:
        want = oracle.permanent_laplace(U, rows, cols, njobs=64)
        ok = np.allclose(got, want, rtol=1e-9, atol=1e-14)
    else:
        # identity: laplace[l] == permanent with column l removed
        l = 3
        want_l = complex(permanent(np.delete(U, l, axis=1), rows, np.ones(k - 1, dtype=np.int32)))
        ok = abs(got[l] - want_l) <= 1e-9 * abs(want_l)
    lbad += (not ok)
    terms = 2.0 ** (k - 2)
    print(f"laplace k={k}: ok={ok} wall={dt*1e3:.2f} ms kernel={ms:.3f} ms  {terms/ms/1e6:.2f} Gterms/s  {terms*22*k/ms/1e9:.2f} TFLOP/s(22k)", flush=True)
# batch: 2000 problems of step k=12 with collisions
B = 2000
mats, rws, cls = [], [], []
for b in range(B):
    k = 12
    U = unitary_group.rvs(16, random_state=b % 50)
    out = rng.multinomial(k - 1, np.ones(16) / 16)
    keep = out > 0
    mats.append(np.ascontiguousarray(U[keep][:, :k])); rws.append(out[keep].astype(np.int32)); cls.append(np.ones(k, dtype=np.int32))
from piquasso_b200.sampling import permanent_laplace_batch
t = time.perf_counter(); res = permanent_laplace_batch(mats, rws, cls); dt = time.perf_counter() - t
t = time.perf_counter(); res = permanent_laplace_batch(mats, rws, cls); dt2 = time.perf_counter() - t
nb = 0
for b in range(0, B, 97):
    want = oracle.permanent_laplace(mats[b], rws[b], cls[b])
    if not np.allclose(res[b], want, rtol=1e-10, atol=1e-14): nb += 1
lbad += nb
print(f"batch {B} problems: {dt*1e3:.1f} ms first, {dt2*1e3:.1f} ms second, kernel {lib.pq_last_kernel_ms(0):.3f} ms, bad={nb}", flush=True)
print("laplace bad", lbad, flush=True)

def time_perm(n, variant, reps=3):
    U = unitary_group.rvs(n, random_state=n)
    ones = np.ones(n, dtype=np.int32)
    lib.pq_set_kernel_choice(variant)
    permanent(U, ones, ones)
    best = 1e30
    for _ in range(reps):
        v = complex(permanent(U, ones, ones))
        best = min(best, lib.pq_last_kernel_ms(0))
    terms = 2.0 ** (n - 1)
    return v, best, terms / (best * 1e-3), (8 * n + 2) * terms / (best * 1e-3) / 1e12

variants = [1, 212, 122, 222, 132, 232]
for n in [20, 26, 30, 32]:
    for v in variants:
        val, ms, tps, tf = time_perm(n, v)
        print(f"time n={n} variant={v}: {ms:.3f} ms  {tps/1e9:.2f} Gterms/s  {tf:.2f} TFLOP/s ({tf/peak*100:.1f}% of DFMA peak)", flush=True)
for n, v in [(36, 222), (36, 122), (40, 222)] if "--n40" in sys.argv else [(36, 222), (36, 122)]:
    val, ms, tps, tf = time_perm(n, v, reps=1)
    print(f"time n={n} variant={v}: {ms:.1f} ms {tps/1e9:.2f} Gterms/s {tf:.2f} TFLOP/s ({tf/peak*100:.1f}%)  value={val}", flush=True)
