import time, numpy as np, sys, os
sys.path.insert(0, os.getcwd())
from scipy.stats import unitary_group
from piquasso_b200 import _lib
lib=_lib.load()
m=60; U=np.ascontiguousarray(unitary_group.rvs(m, random_state=m)); rng=np.random.default_rng(7)
for n,B in ((8,50000),(12,20000)):
    inp=np.array([1]*n+[0]*(m-n),np.int32); outs=np.ascontiguousarray(rng.multinomial(n,np.ones(m)/m,size=B).astype(np.int32))
    cb=np.ascontiguousarray(np.broadcast_to(inp,(B,m))); out=np.zeros(B,np.complex128)
    best=1e9
    for rep in range(5):
        t=time.perf_counter()
        rc=lib.pq_perm_batch_c128(U.ctypes.data_as(_lib.c_double_p), m, m, B, outs.ctypes.data_as(_lib.c_int32_p), cb.ctypes.data_as(_lib.c_int32_p), out.ctypes.data_as(_lib.c_double_p))
        best=min(best,time.perf_counter()-t)
    print("threads", os.environ.get("PQ_PLAN_THREADS"), n, B, "wall %.2f ms kernel %.3f ms"%(best*1e3, lib.pq_last_kernel_ms(0)), rc, flush=True)
