"""The other BASELINE.json configs, measured in the same ``bench.py`` run as the
headline (``secondary`` object of its JSON line).

configs[0]  n=20 Haar permanent through the pybind11 drop-in module (the shape of
            scripts/permanent_benchmark.py:49-65 of the reference)
configs[1]  n=30 Haar permanent (2^29 terms)
configs[2]  60-mode interferometer, 24 photons with occupation multiplicities,
            called unfiltered like piquasso/_simulators/passive/utils.py:131-138
configs[3]  Clifford-Clifford sampling, 100 modes / 25 photons, 10^4 shots
            (piquasso/_simulators/passive/sampling.py:149-236); at N > 1 the shots
            are sharded over the ranks

Every entry carries wall time through the public API with HOST buffers, the
kernels' own time (CUDA events), the algorithmic FP64 rate against the DFMA peak
measured in this run, and -- rank 0, N = 1 -- the unmodified reference C++
(oracle/_ref) on the SAME input on this box's host cores, i.e. a same-workload
ratio.  Nothing here reads /root/reference.
"""

from __future__ import annotations

import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))


def _haar(n, seed):
    from scipy.stats import unitary_group
    return np.ascontiguousarray(unitary_group.rvs(n, random_state=seed), dtype=np.complex128)


def _timed(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    v = None
    for _ in range(reps):
        t0 = time.perf_counter()
        v = fn()
        ts.append(time.perf_counter() - t0)
    return v, float(np.median(ts)), float(np.min(ts))


def _relerr(a, b):
    return abs(a - b) / abs(b)


def cfg3_cases():
    """The three occupation patterns SURVEY.md 8(d) names for configs[2]."""
    r3 = np.random.default_rng(3)
    return {
        "multinomial": (r3.multinomial(24, np.ones(60) / 60), r3.multinomial(24, np.ones(60) / 60)),
        "hard_16ones_4twos": (np.array([1] * 16 + [2] * 4 + [0] * 40),
                              np.array([0] * 30 + [1] * 16 + [2] * 4 + [0] * 10)),
        "heavy_12twos": (np.array([2] * 12 + [0] * 48), np.array([0] * 20 + [2] * 12 + [0] * 28)),
    }


def reference_shot(U, inp, seed, laplace):
    """One shot of the reference's algorithm (sampling.py:208-236, 723-753) on the
    given permanent_laplace; returns (sample, seconds inside permanent_laplace)."""
    d = len(inp)
    rng = np.random.default_rng(seed)
    sample = np.zeros(d, dtype=int)
    cur = np.zeros(d, dtype=int)
    shrink = np.repeat(np.arange(d), inp)
    spent = 0.0
    for _ in range(int(np.sum(inp))):
        ri = rng.choice(len(shrink))
        cur[shrink[ri]] += 1
        shrink = np.delete(shrink, ri)
        nz, oz = cur > 0, sample > 0
        t0 = time.perf_counter()
        part = laplace(np.ascontiguousarray(U[np.ix_(oz, nz)]), sample[oz], cur[nz])
        spent += time.perf_counter() - t0
        idx = np.arange(d)[nz]
        pmf = np.empty(d)
        for m in range(d):
            p = 0.0
            for j in range(len(part)):
                p += cur[idx[j]] * part[j] * U[m, idx[j]]
            pmf[m] = np.abs(p) ** 2
        pmf = pmf / pmf.sum()
        sample[rng.choice(np.arange(d), p=pmf)] += 1
    return tuple(int(x) for x in sample), spent


def run(lib, peak_tflops, rank, world, local_rank, with_reference, sampler_shots=10000):
    """Returns the ``secondary`` dict (rank 0) or None (other ranks)."""
    from piquasso_b200 import _lib, plan as pqplan, sampling
    from piquasso_b200._math.permanent import permanent as perm_ctypes

    out = {}
    oracle = None
    if with_reference:
        import oracle as _oracle
        oracle = _oracle if _oracle.ref_available() else None
    dev = local_rank

    def kernel_ms():
        return float(lib.pq_last_kernel_ms(dev))

    def frac(flops, ms):
        if not (ms > 0 and peak_tflops > 0):
            return None
        return flops / (ms * 1e-3) / 1e12 / peak_tflops

    if rank == 0:
        # a failure in one of these single-rank sections must not take the headline line (or the
        # collective sampler section below) with it: it is reported in the record instead
        try:
            _lib.check(lib.pq_set_devices((ctypes.c_int32 * 1)(dev), 1))
            # ---- configs[0]: n = 20 through the pybind11 module -----------------------
            native = os.path.join(ROOT, "piquasso_b200", "native")
            if native not in sys.path:
                sys.path.insert(0, native)
            try:
                import permanent as pyb
                entry = pyb.permanent
                binding = "pybind11 module piquasso_b200/native/permanent (drop-in for piquasso._math.permanent)"
            except ImportError:
                entry = perm_ctypes
                binding = "ctypes mirror piquasso_b200._math.permanent (pybind11 module not built)"
            for name, n, reps, ref_reps in (("cfg1_n20", 20, 300, 5), ("n24", 24, 50, 3),
                                            ("cfg2_n30", 30, 5, 1)):
                u = _haar(n, n)
                ones = np.ones(n, dtype=np.int32)
                v, med, mn = _timed(lambda: entry(u, ones, ones), reps)
                kms = kernel_ms()
                # the same call with the library's per-launch CUDA-event pair switched off
                # (pq_set_timing(0): what a latency-sensitive caller would configure)
                lib.pq_set_timing(0)
                _, med_off, _ = _timed(lambda: entry(u, ones, ones), reps)
                lib.pq_set_timing(1)
                terms = 2 ** (n - 1)
                flops = terms * (8 * n + 2)
                p = pqplan.plan(ones, ones)
                e = {"n": n, "terms": terms, "wall_ms": med * 1e3, "wall_ms_min": mn * 1e3,
                     "wall_ms_timing_off": med_off * 1e3,
                     "kernel_ms": kms, "terms_per_s": terms / med, "binding": binding,
                     "plan": {"kernel": p["kernel"], "seg_len": p["seg_len"]},
                     "roofline": {"bound": "fp64", "achieved": flops / (kms * 1e-3) / 1e12 if kms > 0 else None,
                                  "peak": peak_tflops, "unit": "TFLOP/s", "frac": frac(flops, kms)},
                     "value": [complex(v).real, complex(v).imag]}
                if oracle is not None:
                    rv, rmed, _ = _timed(lambda: oracle.ref_permanent(u, ones, ones), ref_reps,
                                         warm=1 if n <= 24 else 0)
                    e["reference"] = {"kind": "reference", "cores": os.cpu_count(), "wall_ms": rmed * 1e3,
                                      "same_config": True, "speedup_wall": rmed / med,
                                      "relerr_gpu_vs_reference": _relerr(complex(v), rv)}
                if n == 30:
                    try:
                        gold = json.load(open(os.path.join(ROOT, "tests", "golden", "arbiter.json")))
                        g = [x for x in gold["haar"] if x["n"] == 30 and x["precision"] == 1][0]
                        hi, lo = complex(*g["hi"]), complex(*g["lo"])
                        e["relerr_gpu_vs_long_double_arbiter"] = abs((complex(v) - hi) - lo) / abs(hi)
                        if "reference_cpp" in g:
                            e["relerr_reference_vs_long_double_arbiter"] = abs(
                                (complex(*g["reference_cpp"]) - hi) - lo) / abs(hi)
                    except (OSError, IndexError, KeyError):
                        pass
                out[name] = e
            # ---- configs[2]: 60 modes / 24 photons, unfiltered d x d call -------------
            u60 = _haar(60, 60)
            cfg3 = {}
            for name, (rows, cols) in cfg3_cases().items():
                rows, cols = rows.astype(np.int32), cols.astype(np.int32)
                p = pqplan.plan(rows, cols)
                v, med, mn = _timed(lambda: entry(u60, rows, cols), 50)
                kms = kernel_ms()
                lib.pq_set_timing(0)
                _, med_off, _ = _timed(lambda: entry(u60, rows, cols), 50)
                lib.pq_set_timing(1)
                flops = p["idx_max"] * p["flops_per_term"]
                e = {"idx_max": p["idx_max"], "flops_per_term": p["flops_per_term"],
                     "wall_ms": med * 1e3, "wall_ms_min": mn * 1e3,
                     "wall_ms_timing_off": med_off * 1e3, "kernel_ms": kms,
                     "plan": {"kernel": p["kernel"], "seg_len": p["seg_len"]},
                     "roofline": {"bound": "fp64", "achieved": flops / (kms * 1e-3) / 1e12 if kms > 0 else None,
                                  "peak": peak_tflops, "unit": "TFLOP/s", "frac": frac(flops, kms)}}
                if oracle is not None and p["idx_max"] <= 2 ** 30:
                    rv, rmed, _ = _timed(lambda: oracle.ref_permanent(u60, rows, cols), 3, warm=1)
                    e["reference"] = {"kind": "reference", "cores": os.cpu_count(), "wall_ms": rmed * 1e3,
                                      "same_config": True, "speedup_wall": rmed / med,
                                      "relerr_gpu_vs_reference": _relerr(complex(v), rv)}
                cfg3[name] = e
            out["cfg3_60modes_24photons"] = cfg3
            # ---- a detection-probability batch (one interferometer, many
            # output occupations in one call; the reference loops connector.permanent,
            # passive/utils.py:131-138) -- 20 single-photon inputs in 60 modes
            nph, nout = 20, 2000
            inp20 = np.array([1] * nph + [0] * (60 - nph), dtype=np.int32)
            outs = np.random.default_rng(7).multinomial(nph, np.ones(60) / 60, size=nout).astype(np.int32)
            terms = 0.0
            for r in outs:
                nz = r[r > 0].astype(np.float64)
                nz[np.argmin(nz)] -= 1.0
                terms += float(np.prod(nz + 1.0))
            flops = terms * (8 * nph + 2)
            pv, med, mn = _timed(lambda: sampling.detection_probabilities(u60, inp20, outs), 5, warm=1)
            kms = kernel_ms()
            e = {"modes": 60, "photons": nph, "outputs": nout, "gray_code_terms": terms,
                 "wall_ms": med * 1e3, "wall_ms_min": mn * 1e3, "kernel_ms": kms,
                 "entry": "piquasso_b200.sampling.detection_probabilities -> pq_perm_batch_c128 (host buffers)",
                 "roofline": {"bound": "fp64", "achieved": flops / (kms * 1e-3) / 1e12 if kms > 0 else None,
                              "peak": peak_tflops, "unit": "TFLOP/s", "frac": frac(flops, kms)}}
            if oracle is not None:
                from scipy.special import factorial
                nref = 8
                t0 = time.perf_counter()
                ref = np.array([oracle.ref_permanent(u60, outs[b], inp20) for b in range(nref)])
                rt = (time.perf_counter() - t0) / nref
                pref = np.abs(ref) ** 2 / np.prod(factorial(outs[:nref]), axis=1)
                e["reference"] = {"kind": "reference", "cores": os.cpu_count(),
                                  "ms_per_output": rt * 1e3, "sample": "%d of the %d outputs" % (nref, nout),
                                  "same_config": True, "speedup_per_output": rt / (med / nout),
                                  "max_relerr_gpu_vs_reference": float(np.max(np.abs(pv[:nref] - pref) / pref))}
            out["f2_detection_probability_batch"] = e
        except Exception as exc:  # noqa: BLE001
            out["error"] = "%s: %s" % (type(exc).__name__, exc)

    # ---- configs[3]: the sampler, shots sharded over the ranks ---------------------
    import torch
    import torch.distributed as dist
    from piquasso_b200.distributed import generate_samples_sharded

    u100 = _haar(100, 100)
    inp = np.array([1] * 25 + [0] * 75)
    generate_samples_sharded(inp, 2 * world, u100, 123, device_index=dev)  # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    sampling.TIMERS.clear()
    lib.pq_sampler_work_reset()
    t0 = time.perf_counter()
    samples = generate_samples_sharded(inp, sampler_shots, u100, 123, device_index=dev,
                                       as_array=True)
    dt = time.perf_counter() - t0
    work = (ctypes.c_double * 2)()
    lib.pq_sampler_work(work)
    ksec = sampling.TIMERS.get("  of which GPU kernels (CUDA events)", 0.0)
    t = torch.tensor([dt, ksec, work[0], work[1]], dtype=torch.float64, device="cuda:%d" % dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dt, ksec = float(tmax[0]), float(tmax[1])
        terms, flops = float(tsum[2]), float(tsum[3])
    else:
        terms, flops = float(work[0]), float(work[1])
    if rank != 0:
        return None
    e = {"shots": sampler_shots, "modes": 100, "photons": 25, "n_gpus": world,
         "seconds": dt, "ms_per_shot": dt / sampler_shots * 1e3,
         "kernel_seconds_max_rank": ksec, "gray_code_terms": terms,
         "algorithmic_flops": flops,
         "sharding": "shots split contiguously over the ranks, per-shot seeds; one all_gather_object "
                     "of the finished samples, no data-path collective",
         "roofline": {"bound": "fp64", "unit": "TFLOP/s", "peak": peak_tflops,
                      "achieved": flops / world / ksec / 1e12 if ksec > 0 else None,
                      "frac": flops / world / ksec / 1e12 / peak_tflops if ksec > 0 and peak_tflops > 0 else None,
                      "note": "per GPU: 22k flops per Gray-code term of a k-column Laplace problem "
                              "(SURVEY.md 8d) over the slowest rank's kernel seconds"},
         "output": "samples returned as one (shots, modes) int32 array on every rank",
         "first_sample": [int(x) for x in samples[0]]}
    if oracle is not None:
        t0 = time.perf_counter()
        s0, tl = reference_shot(u100, inp, 123, oracle.ref_permanent_laplace)
        rt = time.perf_counter() - t0
        e["reference"] = {"kind": "reference", "cores": os.cpu_count(), "seconds_per_shot": rt,
                          "seconds_per_shot_in_permanent_laplace": tl, "same_config": True,
                          "first_shot_identical": s0 == tuple(int(x) for x in samples[0]),
                          "speedup_per_shot": rt / (dt / sampler_shots)}
    out["cfg4_sampler_100modes_25photons"] = e
    return out
