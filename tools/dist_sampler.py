"""Run under torchrun: config 4 (100 modes / 25 photons) with the shots sharded over the ranks."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as dist
from scipy.stats import unitary_group
from piquasso_b200 import _lib
from piquasso_b200.distributed import generate_samples_sharded

shots = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
os.environ["NCCL_DEBUG"] = "WARN"
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import ctypes
_lib.load().pq_set_devices((ctypes.c_int32 * 1)(local), 1)
d, n = 100, 25
U = unitary_group.rvs(d, random_state=d)
inp = np.array([1] * n + [0] * (d - n))
generate_samples_sharded(inp, 2 * world, U, 999)  # warm-up
dist.barrier(); torch.cuda.synchronize()
t = time.perf_counter()
samples = generate_samples_sharded(inp, shots, U, 123)
torch.cuda.synchronize(); dist.barrier()
dt = time.perf_counter() - t
if rank == 0:
    print(f"{shots} shots on {world} GPUs: {dt:.3f} s ({dt/shots*1e3:.3f} ms/shot); first sample {samples[0][:12]}... n_samples={len(samples)}", flush=True)
dist.destroy_process_group()
