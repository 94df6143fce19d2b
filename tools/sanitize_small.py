"""Small invocations of every kernel family, for compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.stats import unitary_group
import oracle
from piquasso_b200 import _lib
from piquasso_b200._math.permanent import permanent, permanent_laplace
from piquasso_b200.sampling import permanent_batch, sampler_pmf, sampler_draw, grad_perm, permanent_laplace_batch

lib = _lib.load()
rng = np.random.default_rng(0)
def chk(a, b, what):
    assert abs(a - b) <= 1e-9 * max(1.0, abs(b)), (what, a, b)
for n, choice in ((9, 1), (12, 22), (12, 32), (13, 42), (10, 0)):
    U = unitary_group.rvs(n, random_state=n); ones = np.ones(n, np.int32)
    lib.pq_set_kernel_choice(choice)
    chk(complex(permanent(U, ones, ones)), oracle.permanent(U, ones, ones), ("perm", n, choice))
lib.pq_set_kernel_choice(0)
U = unitary_group.rvs(6, random_state=6)
rows = np.array([1, 1, 0, 3, 2, 2]); cols = np.array([2, 1, 3, 0, 1, 2])
for hint in (0, 1, 7):
    lib.pq_set_seg_len_hint(hint)
    chk(complex(permanent(U, rows, cols)), oracle.permanent(U, rows, cols), ("nary", hint))
lib.pq_set_seg_len_hint(0)
for k in (3, 7, 10, 14, 27):
    a = (rng.normal(size=(4, k)) + 1j * rng.normal(size=(4, k))) / 2
    r = rng.multinomial(k - 1, np.ones(4) / 4); c = np.ones(k, int)
    got = permanent_laplace(a, r, c); want = oracle.permanent_laplace(a, r, c)
    assert np.allclose(got, want, rtol=1e-8), ("laplace", k)
U = unitary_group.rvs(8, random_state=8)
outs = rng.multinomial(3, np.ones(8) / 8, size=20); ins = rng.multinomial(4, np.ones(8) / 8, size=20)
sampler_pmf(U, outs, ins)
permanent_batch(U, rng.multinomial(3, np.ones(8) / 8, size=30), rng.multinomial(3, np.ones(8) / 8, size=30))
grad_perm(U[:4, :4], [1, 2, 0, 1], [1, 1, 1, 1])
print("sanitize_small ok, launches:", lib.pq_launch_count())

# device-side draw (pq_sampler_draw_c128), incl. the multi-threaded planner
Ud = unitary_group.rvs(7, random_state=7)
kk = rng.integers(1, 6, size=2500)
ins_ = np.array([rng.multinomial(k, np.ones(7) / 7) for k in kk])
outs_ = np.array([rng.multinomial(k - 1, np.ones(7) / 7) if k > 1 else np.zeros(7, int) for k in kk])
idx = sampler_draw(Ud, outs_, ins_, rng.random(2500))
assert idx.min() >= 0 and idx.max() < 7
print("sampler_draw ok")

# wide problems (warp-per-segment walk, S = 32) and one large Laplace split over 3 parts
from piquasso_b200.distributed import _laplace_device_partial
gen = np.random.default_rng(10)
rows_w = np.array([40, 35]); uw = np.exp(2j * np.pi * gen.random(2)); vw = np.exp(2j * np.pi * gen.random(75))
import math
exact = float(math.factorial(75)) * np.prod(uw ** rows_w) * np.prod(vw)
got = complex(permanent(np.outer(uw, vw), rows_w, np.ones(75, int)))
assert abs(got - exact) <= 1e-9 * abs(exact), ("wide", got, exact)
aw = (rng.normal(size=(3, 70)) + 1j * rng.normal(size=(3, 70))) / 3
lw = permanent_laplace(aw, [23, 23, 23], np.ones(70, int))
assert lw.shape == (70,) and np.all(np.isfinite(lw))
a16 = np.ascontiguousarray(unitary_group.rvs(18, random_state=3)[:15, :16])
whole = permanent_laplace(a16, np.ones(15, int), np.ones(16, int))
parts = sum(_laplace_device_partial(a16, np.ones(15, np.int32), np.ones(16, np.int32), g, 3) for g in range(3))
assert np.allclose(parts, whole, rtol=1e-12)
print("wide + laplace split ok")

# round 2: the double-double arbiter, walks with more than one CTA (last-CTA reduction),
# the Laplace accumulation modes (zero-multiplicity column -> full product; column
# multiplicities -> chunked passes) and a sampler step
from piquasso_b200 import arbiter
for n in (10, 13):
    U = unitary_group.rvs(n, random_state=n); ones = np.ones(n, np.int32)
    hi, lo = arbiter.permanent_dd(U, ones, ones)
    chk(hi, oracle.permanent(U, ones, ones, precision=1), ("arbiter", n))
    chk(complex(permanent(U, ones, ones)), hi, ("walk after arbiter", n))
hi, lo = arbiter.permanent_dd(unitary_group.rvs(6, random_state=6), rows, cols)
chk(hi, oracle.permanent(unitary_group.rvs(6, random_state=6), rows, cols, precision=1), "arbiter n-ary")
for n, choice in ((15, 1), (16, 32), (17, 22)):
    U = unitary_group.rvs(n, random_state=n); ones = np.ones(n, np.int32)
    lib.pq_set_kernel_choice(choice)
    chk(complex(permanent(U, ones, ones)), oracle.permanent(U, ones, ones), ("multi-CTA", n, choice))
lib.pq_set_kernel_choice(0)
a5 = (rng.normal(size=(3, 6)) + 1j * rng.normal(size=(3, 6))) / 2
for c5 in ([1, 0, 2, 1, 0, 1], [2, 2, 0, 0, 0, 1], [1, 1, 1, 1, 1, 0]):
    r5 = rng.multinomial(sum(c5) - 1, np.ones(3) / 3)
    got = permanent_laplace(a5, r5, c5); want = oracle.permanent_laplace(a5, r5, c5)
    assert np.allclose(got, want, rtol=1e-8), ("laplace modes", c5)
print("round-2 kernels ok, launches:", lib.pq_launch_count())

# batched permanents: one lane per segment (9..32 columns), general columns, and the
# hypercube flavour (three binary rows low, blocks of 8 terms)
U20 = unitary_group.rvs(20, random_state=20)
for ph in (9, 12):
    inp = np.zeros(20, np.int32); inp[:ph] = 1
    outs_b = rng.multinomial(ph, np.ones(20) / 20, size=12).astype(np.int32)
    got = permanent_batch(U20, outs_b, inp)
    for b in range(0, 12, 3):
        chk(got[b], oracle.permanent(U20, outs_b[b], inp), ("batch hyper/one-lane", ph, b))
    inp2 = np.zeros(20, np.int32); inp2[:3] = (ph - 4, 2, 2)
    got = permanent_batch(U20, outs_b[:4], inp2)
    chk(got[0], oracle.permanent(U20, outs_b[0], inp2), ("batch general", ph))
print("batched permanents ok")
