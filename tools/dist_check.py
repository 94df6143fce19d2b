"""Run under torchrun on N GPUs: the distributed permanent equals the single-GPU one."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as dist
from scipy.stats import unitary_group
from piquasso_b200._math.permanent import permanent
from piquasso_b200.distributed import permanent_allgather

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for n in (12, 22, 28):
    U = unitary_group.rvs(n, random_state=n); ones = np.ones(n, dtype=np.int32)
    got = complex(permanent_allgather(U, ones, ones, device_index=local))
    from piquasso_b200 import _lib
    _lib.load().pq_set_devices((__import__("ctypes").c_int32 * 1)(local), 1)
    want = complex(permanent(U, ones, ones))
    e = abs(got - want) / abs(want)
    ok &= e < 1e-12
    if rank == 0: print(f"n={n} world={world} relerr vs single GPU {e:.2e}", flush=True)
rng = np.random.default_rng(3)
rows = rng.multinomial(14, np.ones(10) / 10); cols = rng.multinomial(14, np.ones(10) / 10)
U = unitary_group.rvs(10, random_state=1)
got = complex(permanent_allgather(U, rows, cols, device_index=local)); want = complex(permanent(U, rows, cols))
ok &= abs(got - want) <= 1e-12 * abs(want)
n = 34
u = np.exp(2j * np.pi * rng.random(n)); w = np.exp(2j * np.pi * rng.random(n))
exact = math.factorial(n) * np.prod(u) * np.prod(w)
got = complex(permanent_allgather(np.outer(u, w), np.ones(n, np.int32), np.ones(n, np.int32), device_index=local))
e = abs(got - exact) / abs(exact); ok &= e < 1e-10
if rank == 0: print(f"rank-1 n=34 world={world} relerr vs closed form {e:.2e}; ALL OK = {ok}", flush=True)
# one large permanent_laplace, term space split over the ranks
import time
from piquasso_b200._math.permanent import permanent_laplace
from piquasso_b200.distributed import permanent_laplace_allgather
for k in (16, 29):
    A = np.ascontiguousarray(unitary_group.rvs(k + 3, random_state=k)[: k - 1, :k])
    r = np.ones(k - 1, np.int32); c = np.ones(k, np.int32)
    permanent_laplace_allgather(A, r, c, device_index=local)
    dist.barrier(); t = time.perf_counter()
    got = permanent_laplace_allgather(A, r, c, device_index=local)
    dist.barrier(); dt = time.perf_counter() - t
    t = time.perf_counter(); want = permanent_laplace(A, r, c); dt1 = time.perf_counter() - t
    e = float(np.max(np.abs(got - want) / np.abs(want))); ok &= e < 1e-11
    if rank == 0: print(f"laplace k={k} world={world}: {dt*1e3:.2f} ms vs {dt1*1e3:.2f} ms on one GPU, max relerr {e:.2e}", flush=True)
# latency of the exchange: results left on the device vs staged through the host
for k in (12, 16, 20, 24):
    A = np.ascontiguousarray(unitary_group.rvs(k + 3, random_state=k)[: k - 1, :k])
    r = np.ones(k - 1, np.int32); c = np.ones(k, np.int32)
    res = {}
    for mode in ("device", "host"):
        if mode == "host":
            os.environ["PQ_LAPLACE_ALLGATHER_HOST"] = "1"
        else:
            os.environ.pop("PQ_LAPLACE_ALLGATHER_HOST", None)
        for _ in range(5):
            permanent_laplace_allgather(A, r, c, device_index=local)
        ts = []
        for _ in range(30):
            dist.barrier(); torch.cuda.synchronize(); t = time.perf_counter()
            permanent_laplace_allgather(A, r, c, device_index=local)
            ts.append(time.perf_counter() - t)
        res[mode] = float(np.median(ts)) * 1e3
    os.environ.pop("PQ_LAPLACE_ALLGATHER_HOST", None)
    if rank == 0: print(f"laplace k={k} world={world}: device path {res['device']:.3f} ms, host-staged {res['host']:.3f} ms", flush=True)
# zero-multiplicity column (gets the full product) and the reference's early-out
A = np.ascontiguousarray(unitary_group.rvs(12, random_state=5)[:9, :11])
r = np.ones(9, np.int32); c = np.array([1, 1, 0, 1, 1, 1, 1, 1, 1, 1, 1], np.int32)
got = permanent_laplace_allgather(A, r, c, device_index=local); want = permanent_laplace(A, r, c)
ok &= bool(np.allclose(got, want, rtol=1e-11, atol=0))
got = permanent_laplace_allgather(A, np.zeros(9, np.int32), c, device_index=local)
ok &= got.shape == (1,) and got[0] == 1
if rank == 0: print(f"ALL OK (incl. laplace) = {ok}", flush=True)
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok else 1)
