"""Round-2 GPU probe (dev helper): python tools/r2_probe.py [sections...]
sections: arbiter small cfg3 laplace.  Writes gpurun_out/r2_probe.json."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.stats import unitary_group
from piquasso_b200 import _lib, arbiter, plan as pqplan
from piquasso_b200._math.permanent import permanent, permanent_laplace
from piquasso_b200.sampling import permanent_laplace_batch

lib = _lib.load()
sections = sys.argv[1:] or ["arbiter", "small", "cfg3", "laplace"]
out = {}
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def timed(fn, reps):
    fn()
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); v = fn(); ts.append(time.perf_counter() - t)
    return v, float(np.median(ts)), float(np.min(ts))


if "arbiter" in sections:
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "arbiter.json")))
    res = []
    for e in gold["haar"]:
        n = e["n"]
        u = unitary_group.rvs(n, random_state=n); ones = np.ones(n, np.int32)
        t = time.perf_counter(); hi, lo = arbiter.permanent_dd(u, ones, ones); dt = time.perf_counter() - t
        ghi, glo = complex(*e["hi"]), complex(*e["lo"])
        d_arb = abs((hi - ghi) + (lo - glo)) / abs(ghi)
        v = complex(permanent(u, ones, ones))
        d_gpu = arbiter.relerr_vs(v, ghi, glo)
        res.append({"n": n, "precision": e["precision"], "arbiter_vs_cpu": d_arb, "gpu_vs_cpu": d_gpu,
                    "gpu_vs_arbiter": arbiter.relerr_vs(v, hi, lo), "arbiter_s": dt,
                    "arbiter_kernel_ms": lib.pq_last_kernel_ms(0)})
        print("arbiter", res[-1], flush=True)
    for e in gold["nary"]:
        rows, cols = np.array(e["rows"], np.int32), np.array(e["cols"], np.int32)
        a = (np.array(e["re"]) + 1j * np.array(e["im"])).reshape(len(rows), len(cols))
        hi, lo = arbiter.permanent_dd(a, rows, cols)
        ghi, glo = complex(*e["hi"]), complex(*e["lo"])
        v = complex(permanent(a, rows, cols))
        res.append({"rows": e["rows"], "arbiter_vs_cpu": abs((hi - ghi) + (lo - glo)) / abs(ghi),
                    "gpu_vs_cpu": arbiter.relerr_vs(v, ghi, glo)})
        print("arbiter nary", res[-1], flush=True)
    for n in (34, 36):
        u = unitary_group.rvs(n, random_state=n); ones = np.ones(n, np.int32)
        t = time.perf_counter(); hi, lo = arbiter.permanent_dd(u, ones, ones); dt = time.perf_counter() - t
        v = complex(permanent(u, ones, ones))
        res.append({"n": n, "gpu_vs_arbiter": arbiter.relerr_vs(v, hi, lo), "arbiter_s": dt,
                    "arbiter": [hi.real, hi.imag, lo.real, lo.imag], "gpu": [v.real, v.imag],
                    "gpu_kernel_ms": lib.pq_last_kernel_ms(0)})
        print("arbiter", res[-1], flush=True)
    out["arbiter"] = res

if "small" in sections:
    sys.path.insert(0, os.path.join(ROOT, "piquasso_b200", "native"))
    import permanent as pyb  # the pybind11 drop-in module
    res = []
    for n in (8, 10, 12, 14, 16, 18, 20, 22, 24, 26, 28):
        u = unitary_group.rvs(n, random_state=n); ones = np.ones(n, np.int32)
        row = {"n": n}
        for choice in (1, 2):
            lib.pq_set_kernel_choice(choice)
            reps = 200 if n <= 22 else 30
            v, med, mn = timed(lambda: pyb.permanent(u, ones, ones), reps)
            row["choice%d" % choice] = {"wall_us_median": med * 1e6, "wall_us_min": mn * 1e6,
                                        "kernel_us": lib.pq_last_kernel_ms(0) * 1e3,
                                        "plan": pqplan.plan(ones, ones)["seg_len"], "value": [complex(v).real, complex(v).imag]}
        lib.pq_set_kernel_choice(0)
        v, med, mn = timed(lambda: pyb.permanent(u, ones, ones), 200 if n <= 22 else 30)
        kus = lib.pq_last_kernel_ms(0) * 1e3
        p = pqplan.plan(ones, ones)
        row["auto"] = {"wall_us_median": med * 1e6, "wall_us_min": mn * 1e6, "kernel_us": kus,
                       "kernel": p["kernel"], "seg_len": p["seg_len"],
                       "tflops": 2.0 ** (n - 1) * (8 * n + 2) / (kus * 1e-6) / 1e12}
        _, medc, _ = timed(lambda: permanent(u, ones, ones), 50)
        row["ctypes_wall_us_median"] = medc * 1e6
        res.append(row)
        print("small", row, flush=True)
    out["small"] = res

if "cfg3" in sections:
    U60 = unitary_group.rvs(60, random_state=60)
    r3 = np.random.default_rng(3)
    cases = {"multinomial": (r3.multinomial(24, np.ones(60) / 60), r3.multinomial(24, np.ones(60) / 60)),
             "hard_16ones_4twos": (np.array([1] * 16 + [2] * 4 + [0] * 40), np.array([0] * 30 + [1] * 16 + [2] * 4 + [0] * 10)),
             "heavy_12twos": (np.array([2] * 12 + [0] * 48), np.array([0] * 20 + [2] * 12 + [0] * 28))}
    res = {}
    for name, (rows, cols) in cases.items():
        rows = rows.astype(np.int32); cols = cols.astype(np.int32)
        rr = []
        for hint in (0, 8, 16, 32, 64, 128, 256, 512, 1024):
            lib.pq_set_seg_len_hint(hint)
            p = pqplan.plan(rows, cols)
            v, med, mn = timed(lambda: complex(permanent(U60, rows, cols)), 20)
            kus = lib.pq_last_kernel_ms(0) * 1e3
            rr.append({"hint": hint, "seg_len": p["seg_len"], "idx_max": p["idx_max"], "wall_us": med * 1e6,
                       "kernel_us": kus, "tflops": p["idx_max"] * p["flops_per_term"] / (kus * 1e-6) / 1e12})
            print("cfg3", name, rr[-1], flush=True)
        lib.pq_set_seg_len_hint(0)
        res[name] = rr
    out["cfg3"] = res

if "laplace" in sections:
    res = []
    for k, batch in ((22, 2000), (23, 2000), (24, 2000), (25, 2000), (25, 1), (20, 2000), (16, 5000), (12, 10000)):
        a = np.ascontiguousarray(unitary_group.rvs(30, random_state=k)[: k - 1, :k]); r = np.ones(k - 1, np.int32); c = np.ones(k, np.int32)
        best = 1e30
        for _ in range(3):
            if batch == 1:
                permanent_laplace(a, r, c)
            else:
                permanent_laplace_batch([a] * batch, [r] * batch, [c] * batch)
            best = min(best, lib.pq_last_kernel_ms(0))
        terms = batch * 2.0 ** (k - 2)
        res.append({"k": k, "batch": batch, "kernel_ms": best, "ps_per_term": best * 1e9 / terms,
                    "alg_tflops": terms * 22 * k / (best * 1e-3) / 1e12})
        print("laplace", res[-1], flush=True)
    out["laplace"] = res

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2_probe.json"), "w"), indent=1, default=float)
