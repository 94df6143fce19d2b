"""All five BASELINE.json configs on one B200, each beside the reference's own C++
(oracle/_ref) on the box's host cores.  Writes gpurun_out/configs.json."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.stats import unitary_group
import oracle
from piquasso_b200 import _lib, sampling
from piquasso_b200._math.permanent import permanent

lib = _lib.load()
out = {"cores": os.cpu_count()}

def timed(fn, reps):
    fn()
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); v = fn(); ts.append(time.perf_counter() - t)
    return v, float(np.median(ts))

# config 1: n=20 Haar (scripts/permanent_benchmark.py shape)
for name, n, reps, ref_reps in (("cfg1_n20", 20, 50, 5), ("n24", 24, 20, 3), ("cfg2_n30", 30, 10, 1)):
    U = unitary_group.rvs(n, random_state=n); ones = np.ones(n, dtype=np.int32)
    v, t = timed(lambda: complex(permanent(U, ones, ones)), reps)
    kms = lib.pq_last_kernel_ms(0)
    rv, rt = timed(lambda: oracle.ref_permanent(U, ones, ones), ref_reps) if n <= 30 else (None, None)
    out[name] = {"gpu_wall_ms": t * 1e3, "gpu_kernel_ms": kms, "ref_cpu_ms": rt * 1e3, "speedup": rt / t,
                 "relerr_vs_ref": abs(v - rv) / abs(rv), "terms": 2 ** (n - 1)}
    print(name, out[name], flush=True)

# config 1b: the script's literal shape: d=2..20, every multiplicity 2 is infeasible beyond d~10 for the
# reference (3^d terms); d=10
d = 10
rng = np.random.default_rng(1)
A = rng.random((d, d)) + 1j * rng.random((d, d)); A = A + A.T
twos = 2 * np.ones(d, dtype=np.int32)
v, t = timed(lambda: complex(permanent(A, twos, twos)), 20)
rv, rt = timed(lambda: oracle.ref_permanent(A, twos, twos), 3)
out["cfg1b_d10_mult2"] = {"gpu_wall_ms": t * 1e3, "ref_cpu_ms": rt * 1e3, "speedup": rt / t, "relerr_vs_ref": abs(v - rv) / abs(rv)}
print("cfg1b", out["cfg1b_d10_mult2"], flush=True)

# config 3: 60-mode interferometer, 24 photons, unfiltered d x d call
U60 = unitary_group.rvs(60, random_state=60)
r3 = np.random.default_rng(3)
cases = {"multinomial": (r3.multinomial(24, np.ones(60) / 60), r3.multinomial(24, np.ones(60) / 60)),
         "hard_16ones_4twos": (np.array([1] * 16 + [2] * 4 + [0] * 40), np.array([0] * 30 + [1] * 16 + [2] * 4 + [0] * 10)),
         "heavy_12twos": (np.array([2] * 12 + [0] * 48), np.array([0] * 20 + [2] * 12 + [0] * 28))}
for name, (rows, cols) in cases.items():
    rows = rows.astype(np.int32); cols = cols.astype(np.int32)
    v, t = timed(lambda: complex(permanent(U60, rows, cols)), 10)
    kms = lib.pq_last_kernel_ms(0)
    from piquasso_b200 import plan
    p = plan.plan(rows, cols)
    if p["idx_max"] <= 2 ** 30:
        rv, rt = timed(lambda: oracle.ref_permanent(U60, rows, cols), 1)
        out["cfg3_" + name] = {"idx_max": p["idx_max"], "gpu_wall_ms": t * 1e3, "gpu_kernel_ms": kms, "ref_cpu_ms": rt * 1e3,
                               "speedup": rt / t, "relerr_vs_ref": abs(v - rv) / abs(rv)}
    else:
        out["cfg3_" + name] = {"idx_max": p["idx_max"], "gpu_wall_ms": t * 1e3, "gpu_kernel_ms": kms}
    print("cfg3", name, out["cfg3_" + name], flush=True)

# config 4: Clifford-Clifford 100 modes / 25 photons
shots = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
U = unitary_group.rvs(100, random_state=100)
inp = np.array([1] * 25 + [0] * 75)
sampling.generate_samples(inp, 2, U, 123)
sampling.TIMERS.clear()
t = time.perf_counter(); samples = sampling.generate_samples(inp, shots, U, 123); dt = time.perf_counter() - t
out["cfg4_sampler"] = {"shots": shots, "seconds": dt, "ms_per_shot": dt / shots * 1e3, "timers": dict(sampling.TIMERS),
                       "extrapolated_10k_shots_s": dt / shots * 1e4}
print("cfg4", out["cfg4_sampler"], flush=True)
# reference: one shot of the same algorithm on the compiled reference permanent_laplace
def ref_shot(seed):
    d = 100; n = 25
    rng = np.random.default_rng(seed)
    sample = np.zeros(d, dtype=int); cur = np.zeros(d, dtype=int); shrink = np.repeat(np.arange(d), inp)
    tl = 0.0
    for _ in range(n):
        ri = rng.choice(len(shrink)); cur[shrink[ri]] += 1; shrink = np.delete(shrink, ri)
        nz = cur > 0; oz = sample > 0
        t0 = time.perf_counter()
        part = oracle.ref_permanent_laplace(np.ascontiguousarray(U[np.ix_(oz, nz)]), sample[oz], cur[nz])
        tl += time.perf_counter() - t0
        amp = U[:, nz] @ (cur[nz] * part)
        pmf = np.abs(amp) ** 2; pmf /= pmf.sum()
        sample[rng.choice(np.arange(d), p=pmf)] += 1
    return tuple(int(x) for x in sample), tl
t = time.perf_counter(); s0, tl = ref_shot(123); rt = time.perf_counter() - t
out["cfg4_sampler"].update({"ref_cpu_s_per_shot": rt, "ref_cpu_s_in_permanent_laplace": tl,
                            "ref_first_shot_identical": s0 == samples[0], "speedup_per_shot": rt / (dt / shots)})
print("cfg4 ref", rt, tl, s0 == samples[0], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/configs.json", "w"), indent=1, default=float)
