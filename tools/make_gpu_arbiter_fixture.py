"""Haar n > 32 fixtures from the GPU's double-double arbiter (B200 box):

    python tools/make_gpu_arbiter_fixture.py 34 36 38 40

writes gpurun_out/arbiter_gpu.json (copy to tests/golden/arbiter_gpu.json).  The
arbiter kernel is pinned against CPU binary128 / long double up to n = 32
(tests/test_gpu_envelope.py); beyond that no CPU arbiter is affordable and the
reference itself is wrong (src/n_aryGrayCodeCounter.hpp:179), so the arbiter's
value is recorded once here (n = 40: ~6 minutes on one B200) and the production
walks are held against it.  Each entry also records the production result of the
same run and its relative error."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.stats import unitary_group
from piquasso_b200 import _lib, arbiter
from piquasso_b200._math.permanent import permanent

lib = _lib.load()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.path.join(ROOT, "gpurun_out", "arbiter_gpu.json")
os.makedirs(os.path.dirname(path), exist_ok=True)
doc = {"generator": "scipy.stats.unitary_group.rvs(n, random_state=n), rows = cols = 1",
       "arbiter": "pq_perm_arbiter_c128 (double-double, csrc/pqperm_arbiter.cu)", "haar": []}
for n in [int(x) for x in sys.argv[1:]] or [34, 36]:
    u = unitary_group.rvs(n, random_state=n)
    ones = np.ones(n, np.int32)
    t0 = time.perf_counter()
    hi, lo = arbiter.permanent_dd(u, ones, ones)
    dt = time.perf_counter() - t0
    v = complex(permanent(u, ones, ones))
    kms = lib.pq_last_kernel_ms(0)
    e = {"n": n, "seed": n, "hi": [hi.real, hi.imag], "lo": [lo.real, lo.imag],
         "arbiter_seconds": dt, "production": [v.real, v.imag], "production_kernel_ms": kms,
         "production_relerr": arbiter.relerr_vs(v, hi, lo)}
    doc["haar"].append(e)
    print(e, flush=True)
    json.dump(doc, open(path, "w"), indent=1)
