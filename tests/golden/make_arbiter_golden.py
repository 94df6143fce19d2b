"""Extended-precision fixtures that pin the GPU's double-double arbiter kernel
(pq_perm_arbiter_c128) and the accuracy envelope of the production walks.

    python tests/golden/make_arbiter_golden.py [--quad-max N] [--ld-max N]

Run in the BUILD container (CPU only; minutes to an hour).  Writes
tests/golden/arbiter.json:

* Haar unitaries ``scipy.stats.unitary_group.rvs(n, random_state=n)`` (the
  generator of the reference's fixtures, tests/conftest.py:61-69 there), all
  multiplicities 1: the permanent from oracle/perm_oracle.c in software
  binary128 (n <= --quad-max) and in long double (n <= --ld-max), as (hi, lo)
  double pairs;
* for n = 30 (BASELINE config 2) also the value of the UNMODIFIED reference C++
  (oracle/_ref), so that the reference's own error against the arbiter is on
  record beside ours;
* a few repeated-row/column cases (n-ary digits) in binary128.
"""
import json
import os
import sys
import time

import numpy as np
from scipy.stats import unitary_group

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "arbiter.json")


def arg(name, default):
    for i, a in enumerate(sys.argv):
        if a == name:
            return int(sys.argv[i + 1])
    return default


def pair(z):
    return [float(z.real), float(z.imag)]


def main():
    quad_max = arg("--quad-max", 26)
    ld_max = arg("--ld-max", 32)
    threads = oracle.num_threads()
    doc = {"generator": "scipy.stats.unitary_group.rvs(n, random_state=n), rows = cols = 1",
           "haar": [], "nary": []}
    if os.path.exists(OUT):
        doc = json.load(open(OUT))
    have = {(e["n"], e["precision"]) for e in doc["haar"]}

    def save():
        with open(OUT, "w") as fh:
            json.dump(doc, fh, indent=1)

    for n in range(14, max(quad_max, ld_max) + 1, 2):
        u = unitary_group.rvs(n, random_state=n)
        ones = np.ones(n, dtype=np.int32)
        for precision, limit in ((2, quad_max), (1, ld_max)):
            if n > limit or (n, precision) in have:
                continue
            t0 = time.time()
            hi, lo = oracle.permanent_hilo(u, ones, ones, njobs=8 * threads, precision=precision)
            entry = {"n": n, "seed": n, "precision": precision, "hi": pair(hi), "lo": pair(lo),
                     "seconds": round(time.time() - t0, 2)}
            if n == 30 and oracle.ref_available() and precision == 1:
                entry["reference_cpp"] = pair(oracle.ref_permanent(u, ones, ones))
            doc["haar"].append(entry)
            print(entry, flush=True)
            save()
    if not doc["nary"]:
        rng = np.random.default_rng(7)
        for rows, cols in (([2, 1, 3, 0, 2, 1, 1, 2], [1, 2, 2, 1, 0, 3, 2, 1]),
                           ([1] * 10 + [2] * 4, [2] * 4 + [1] * 10),
                           ([3, 3, 3, 3, 2, 2], [1] * 16)):
            rows, cols = np.array(rows, dtype=np.int32), np.array(cols, dtype=np.int32)
            a = (rng.normal(size=(len(rows), len(cols))) + 1j * rng.normal(size=(len(rows), len(cols)))) / 3
            hi, lo = oracle.permanent_hilo(a, rows, cols, njobs=8 * threads, precision=2)
            doc["nary"].append({"rows": rows.tolist(), "cols": cols.tolist(),
                                "re": a.real.ravel().tolist(), "im": a.imag.ravel().tolist(),
                                "precision": 2, "hi": pair(hi), "lo": pair(lo)})
        save()
    save()


if __name__ == "__main__":
    main()
