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
if rank == 0: print(f"ALL OK (incl. laplace) = {ok}", flush=True)
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok else 1)
