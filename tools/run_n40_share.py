"""python tools/run_n40_share.py [NPARTS]: one launch of the headline instantiation
(perm_walk_binary_pm<40,3,64>) over 1/NPARTS of the n=40 term space -- the same
kernel, grid and per-thread work pattern as the full permanent, short enough for
`ncu --set full` (which replays the kernel ~40 times).  Dev helper."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.stats import unitary_group
from piquasso_b200 import _lib

nparts = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lib = _lib.load()
n = 40
a = np.ascontiguousarray(unitary_group.rvs(n, random_state=n), dtype=np.complex128)
ones = np.ones(n, dtype=np.int32)
import torch
out = torch.zeros(4, dtype=torch.float64, device="cuda:0")
status = ctypes.c_int(0)
triv = np.zeros(2)
for rep in range(2):
    _lib.check(lib.pq_perm_partial_c128(
        a.ctypes.data_as(_lib.c_double_p), n, n, ones.ctypes.data_as(_lib.c_int32_p),
        ones.ctypes.data_as(_lib.c_int32_p), 0, nparts, 0, None, ctypes.c_void_p(out.data_ptr()),
        ctypes.byref(status), triv.ctypes.data_as(_lib.c_double_p)))
    print("share 1/%d: %.3f ms, partial %s" % (nparts, lib.pq_last_kernel_ms(0), out.cpu().numpy()), flush=True)
