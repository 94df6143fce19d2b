"""One-off differential check (build container only: needs /root/reference): random small
configurations of every per-shot sampler algorithm, the reference's generate_samples /
generate_lossy_samples against piquasso_b200's (pmf rows from the oracle on both sides'
permanent_laplace).  python tools/fuzz_sampler_variants.py [N]"""
import os, sys
from types import SimpleNamespace
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, "/root/reference")
import numpy as np
from scipy.stats import unitary_group
import oracle
import make_golden
make_golden._install_stubs()
import piquasso  # noqa: F401  (the reference)
from piquasso._simulators.passive import sampling as ref
from piquasso.api.exceptions import InvalidSimulation as RefInvalid
from conftest import oracle_pmf_rows
from piquasso_b200 import sampling as mine
from piquasso_b200.shot_engine import InvalidSimulation

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(2026)
bad = 0
for trial in range(N):
    d = int(rng.integers(2, 6)); n = int(rng.integers(1, 5))
    inp = rng.multinomial(n, np.ones(d) / d)
    t = float(rng.uniform(0.5, 1.0))
    kind = int(rng.integers(0, 6))
    U = unitary_group.rvs(d, random_state=int(rng.integers(1 << 30))) if d > 1 else np.eye(1, dtype=complex)
    seed = int(rng.integers(1 << 20)); shots = int(rng.integers(1, 7))
    post = ((), (), 200); overlap = None; loss = None; dilation = False
    if kind in (1, 3, 5):
        k = int(rng.integers(1, min(d, 3)))
        modes = tuple(int(x) for x in sorted(rng.choice(d, size=k, replace=False)))
        post = (modes, tuple(int(x) for x in rng.integers(0, 2, size=k)), 200)
    if kind in (2, 3):
        overlap = float(rng.uniform(0.1, 0.9))
    if kind in (3, 4) or (kind == 2 and rng.random() < 0.5):
        loss = (int(rng.integers(1 << 20)), t); U = np.sqrt(t) * U
    if kind == 5:
        dilation = True
        U = U @ np.diag(np.sqrt(rng.uniform(0.3, 0.95, size=d))) @ unitary_group.rvs(d, random_state=trial + 1)
    config = SimpleNamespace(seed_sequence=seed, use_dask=False)
    def reject_pair():
        if loss is None:
            return (lambda: False), None
        a = np.random.default_rng(loss[0]); b = np.random.default_rng(loss[0])
        return (lambda: a.uniform() > loss[1]), (lambda: b.uniform() > loss[1])
    r_ref, r_mine = reject_pair()
    try:
        if dilation:
            want = ref.generate_lossy_samples(np.array(inp), shots, oracle.ref_permanent_laplace, U, post, config)
        else:
            want = ref.generate_samples(np.array(inp), shots, oracle.ref_permanent_laplace, U, r_ref, post, overlap, config)
        want = [tuple(int(x) for x in s) for s in want]
    except RefInvalid:
        want = "invalid"
    except ValueError as exc:   # e.g. the reference's own negative-probability failure
        want = "valueerror"
    try:
        if dilation:
            got = mine.generate_lossy_samples(inp, shots, U, seed, postselect_data=post, pmf_rows=oracle_pmf_rows)
        else:
            got = mine.generate_samples(inp, shots, U, seed, reject_condition=r_mine, postselect_data=post,
                                        uniform_particle_overlap=overlap, pmf_rows=oracle_pmf_rows)
    except InvalidSimulation:
        got = "invalid"
    except ValueError:
        got = "valueerror"
    if got != want:
        bad += 1
        print("MISMATCH trial", trial, dict(kind=kind, d=d, inp=inp.tolist(), post=post, overlap=overlap, loss=loss, dilation=dilation))
        print("  want", want); print("  got ", got)
print("checked", N, "configurations; mismatches:", bad)
