"""Config 4 (BASELINE.json): Clifford-Clifford boson sampling, 100 modes / 25 photons.
python tools/sampler_bench.py SHOTS [MODES PHOTONS]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.stats import unitary_group
from piquasso_b200 import _lib, sampling

pos = [a for a in sys.argv[1:] if not a.startswith("--")]
shots = int(pos[0]); d = int(pos[1]) if len(pos) > 1 else 100; n = int(pos[2]) if len(pos) > 2 else 25
U = unitary_group.rvs(d, random_state=d)
inp = np.array([1] * n + [0] * (d - n))
lib = _lib.load()
kw = {}
for a in sys.argv:
    if a.startswith("--fractions="):   # batch sizes handed to the overlap threads
        sampling._OVERLAP_FRACTIONS = tuple(float(x) for x in a.split("=")[1].split(","))
    if a.startswith("--overlap="):     # worker threads (1 = single batch, no overlap)
        kw["overlap"] = int(a.split("=")[1])
    if a.startswith("--devices="):     # single process, shots sharded over these GPUs
        kw["devices"] = [int(x) for x in a.split("=")[1].split(",")]
sampling.generate_samples(inp[:], 2 * len(kw.get('devices', [0])), U, 123, **kw)  # warm-up
sampling.TIMERS.clear()
t = time.perf_counter()
samples = sampling.generate_samples(inp, shots, U, 123, **kw)
dt = time.perf_counter() - t
print(f"{shots} shots, {d} modes, {n} photons: {dt:.2f} s  ({dt/shots*1e3:.2f} ms/shot)", flush=True)
for k, v in sampling.TIMERS.items():
    print(f"   {k}: {v:.3f} s")
print("first samples:", samples[:2])
if "--account" in sys.argv:
    # Gray-code terms actually walked: replay the shots' occupations on the host
    import numpy as np
    gpu_s = sampling.TIMERS.get("  of which GPU kernels (CUDA events)", 0.0)
    terms = 0.0; flops = 0.0
    for smp_i, smp in enumerate(samples[: min(shots, 500)]):
        # the order in which photons were placed is not recorded; the term count of
        # step k only needs the multiset of occupied output modes after k-1 photons,
        # which we re-draw here with the same generator the sampler used
        rng = np.random.default_rng(123 + smp_i)
        out = np.zeros(d, dtype=int); cur = np.zeros(d, dtype=int); shrink = np.repeat(np.arange(d), inp)
        U_ = U
        for k in range(1, n + 1):
            ri = rng.integers(0, len(shrink)); cur[shrink[ri]] += 1; shrink = np.delete(shrink, ri)
            rows = out[out > 0].copy()
            if rows.size:
                rows[np.argmin(rows)] -= 1
                t = float(np.prod(rows + 1))
                terms += t; flops += t * 22 * k
            pmf = sampling.sampler_pmf(U_, out[None, :], cur[None, :])[0]
            p = pmf / np.cumsum(pmf)[-1]; cdf = np.cumsum(p); cdf /= cdf[-1]
            out[(cdf <= rng.random()).sum()] += 1
        assert tuple(out) == smp
    scale = shots / min(shots, 500)
    print(f"GPU kernel time {gpu_s:.3f} s -> {terms*scale/gpu_s/1e9:.2f} G terms/s, "
          f"{flops*scale/gpu_s/1e12:.2f} algorithmic TFLOP/s")
    print(f"Gray-code terms walked (extrapolated from {min(shots,500)} shots): {terms*scale:.3e}; "
          f"algorithmic flops (22k per term, SURVEY 8d): {flops*scale:.3e}")

if "--check" in sys.argv:
    import oracle
    # reference-style sequential sampler on the oracle's permanent_laplace for the first shots
    def ref_sample(seed):
        rng = np.random.default_rng(seed)
        sample = np.zeros(d, dtype=int); cur = np.zeros(d, dtype=int); shrink = np.repeat(np.arange(d), inp)
        for _ in range(n):
            ri = rng.choice(len(shrink)); cur[shrink[ri]] += 1; shrink = np.delete(shrink, ri)
            nz = cur > 0; oz = sample > 0
            part = oracle.ref_permanent_laplace(U[np.ix_(oz, nz)], sample[oz], cur[nz])
            idx = np.arange(d)[nz]
            pmf = np.empty(d)
            for m in range(d):
                p = 0.0
                for j in range(len(part)):
                    p += cur[idx[j]] * part[j] * U[m, idx[j]]
                pmf[m] = np.abs(p) ** 2
            pmf = pmf / pmf.sum()
            sample[rng.choice(np.arange(d), p=pmf)] += 1
        return tuple(int(x) for x in sample)
    ncheck = int(os.environ.get("NCHECK", "3"))
    t = time.perf_counter()
    ok = all(ref_sample(123 + i) == samples[i] for i in range(ncheck))
    print(f"first {ncheck} shots identical to the reference algorithm on the compiled reference: {ok}  ({(time.perf_counter()-t)/ncheck:.2f} s/shot on CPU)")
