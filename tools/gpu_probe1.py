"""Development probe (run under gpurun): correctness of every kernel variant
against the oracle, DFMA peak, and kernel timings."""
import ctypes, json, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.stats import unitary_group
import oracle
from piquasso_b200 import _lib
from piquasso_b200._math.permanent import permanent

lib = _lib.load()
res = {}
print("devices", lib.pq_device_count(), flush=True)
peak = lib.pq_fp64_peak_tflops(0, 1 << 17)
print("fp64 peak TFLOP/s", peak, flush=True)
res["fp64_peak_tflops"] = peak

def relerr(a, b):
    return abs(a - b) / max(abs(b), 1e-300)

bad = 0
# binary correctness, all variants
variants = [1, 212, 122, 222, 132, 232]
for n in [1, 2, 3, 5, 8, 9, 12, 13, 16, 18]:
    U = unitary_group.rvs(n, random_state=n) if n > 1 else np.array([[0.3 + 0.4j]])
    ones = np.ones(n, dtype=np.int32)
    want = oracle.permanent(U, ones, ones, precision=1)
    for v in variants:
        lib.pq_set_kernel_choice(v)
        got = complex(permanent(U, ones, ones))
        e = relerr(got, want)
        info = _lib.PlanInfo()
        lib.pq_perm_plan(n, n, ones.ctypes.data_as(_lib.c_int32_p), ones.ctypes.data_as(_lib.c_int32_p), ctypes.byref(info))
        flag = "" if e < 1e-11 else "  <<<<<< BAD"
        bad += e >= 1e-11
        print(f"n={n} variant={v} kernel={info.kernel} W={info.seg_len} nseg={info.nseg} relerr={e:.2e}{flag}", flush=True)
lib.pq_set_kernel_choice(0)
# n-ary correctness
rng = np.random.default_rng(7)
for trial in range(40):
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
        flag = "" if e < 1e-10 or abs(got - want) < 1e-13 else "  <<<<<< BAD"
        bad += bool(flag)
        if flag or hint == 0:
            print(f"nary d={d} rows={rows.tolist()} cols={cols.tolist()} hint={hint} relerr={e:.2e} abs={abs(got-want):.2e}{flag}", flush=True)
lib.pq_set_seg_len_hint(0)
print("BAD COUNT", bad, flush=True)
res["bad"] = int(bad)

# timings
def time_perm(n, variant, reps=3):
    U = unitary_group.rvs(n, random_state=n)
    ones = np.ones(n, dtype=np.int32)
    lib.pq_set_kernel_choice(variant)
    permanent(U, ones, ones)
    best = 1e30
    for _ in range(reps):
        t = time.perf_counter(); v = complex(permanent(U, ones, ones)); dt = time.perf_counter() - t
        best = min(best, lib.pq_last_kernel_ms(0))
    terms = 2.0 ** (n - 1)
    fl = (8 * n + 2) * terms
    return v, best, terms / (best * 1e-3), fl / (best * 1e-3) / 1e12

timing = []
for n in [20, 24, 28, 30, 32]:
    for v in variants:
        val, ms, tps, tf = time_perm(n, v)
        print(f"time n={n} variant={v}: {ms:.3f} ms  {tps/1e9:.2f} Gterms/s  {tf:.2f} TFLOP/s ({tf/peak*100:.1f}% of DFMA peak)", flush=True)
        timing.append(dict(n=n, variant=v, ms=ms, gterms=tps / 1e9, tflops=tf))
res["timing"] = timing
# n=36 best variant only
best_v = max([t for t in timing if t["n"] == 32], key=lambda t: t["tflops"])["variant"]
for n in [36, 40 if "--n40" in sys.argv else 34]:
    val, ms, tps, tf = time_perm(n, best_v, reps=1)
    print(f"time n={n} variant={best_v}: {ms:.1f} ms {tps/1e9:.2f} Gterms/s {tf:.2f} TFLOP/s ({tf/peak*100:.1f}%)", flush=True)
    timing.append(dict(n=n, variant=best_v, ms=ms, gterms=tps / 1e9, tflops=tf))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/probe1.json", "w"), indent=1)
