"""Phase timeline of the generic walk on small problems (experiment build only):
   tools/build_variant.sh trace "-DPQ_TRACE=1"
   PQ_LIB_PATH=variants/libpqperm_trace.so python tools/trace_small.py
Thread 0 of every CTA stamps %globaltimer at: 0 entry, 1 matrix staged, 2 tables built,
3 first seed done, 4 walk done, 5 after finish_grid.  Printed: per-phase durations over
the CTAs (min / median / max, us) and the span first entry -> last exit."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.stats import unitary_group
from piquasso_b200 import _lib, plan as pqplan
from piquasso_b200._math.permanent import permanent

lib = _lib.load()
raw = ctypes.CDLL(_lib.LIB_PATH)
raw.pq_debug_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_int]

U60 = unitary_group.rvs(60, random_state=60)
r60 = np.random.default_rng(3)
cases = {}
for n in (8, 16, 20, 22, 24):
    cases["haar%d" % n] = (unitary_group.rvs(n, random_state=n), np.ones(n, np.int32), np.ones(n, np.int32))
cases["cfg3_multinomial"] = (U60, r60.multinomial(24, np.ones(60) / 60).astype(np.int32),
                             r60.multinomial(24, np.ones(60) / 60).astype(np.int32))
hard = np.array([1] * 16 + [2] * 4 + [0] * 40, np.int32)
heavy = np.array([2] * 12 + [0] * 48, np.int32)
cases["cfg3_hard"] = (U60, hard, hard)
cases["cfg3_heavy"] = (U60, heavy, heavy)

lib.pq_set_kernel_choice(1)  # the generic walk is the instrumented one
for name, (a, r, c) in cases.items():
    for _ in range(20):
        permanent(a, r, c)
    ts = []
    for _ in range(30):
        t = time.perf_counter(); permanent(a, r, c); ts.append(time.perf_counter() - t)
    kms = lib.pq_last_kernel_ms(0)
    buf = np.zeros(8 * 8192, np.uint64)
    rc = raw.pq_debug_trace_read(buf.ctypes.data, buf.size)
    assert rc == 0, rc
    info = pqplan.plan_info(r, c) if hasattr(pqplan, "plan_info") else None
    t = buf.reshape(8192, 8).astype(np.int64)
    t0min = None
    # CTAs of THIS launch: entry stamp within 1 ms of the newest entry stamp
    newest = t[:, 0].max()
    live = (t[:, 0] > newest - 1_000_000) & (t[:, 5] >= t[:, 0])
    t = t[live]
    t0min = t[:, 0].min()
    span = (t[:, 5].max() - t0min) / 1e3
    print("%-18s wall %.1f us  kernel(events) %.1f us  CTAs %d  span %.2f us  plan %s"
          % (name, 1e6 * np.median(ts), 1e3 * kms, len(t), span, info), flush=True)
    labels = ["entry skew", "stage matrix", "tables", "first seed", "walk (rest)", "finish"]
    rel = [(t[:, 0] - t0min)] + [t[:, k + 1] - t[:, k] for k in range(5)]
    for lab, d in zip(labels, rel):
        d = d / 1e3
        print("    %-14s min %7.2f  med %7.2f  max %7.2f us" % (lab, d.min(), np.median(d), d.max()))
    last = t[np.argmax(t[:, 5])]
    print("    last CTA: entry +%.2f, walk done +%.2f, exit +%.2f us; latest walk-done of any CTA +%.2f"
          % ((last[0] - t0min) / 1e3, (last[4] - t0min) / 1e3, (last[5] - t0min) / 1e3,
             (t[:, 4].max() - t0min) / 1e3))
