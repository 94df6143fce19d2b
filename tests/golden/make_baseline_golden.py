"""Golden fixtures AT THE BASELINE.json SHAPES, produced by the reference itself.

    python tests/golden/make_baseline_golden.py        (build container, ~1 minute)

Writes tests/golden/baseline_shapes.json:

* configs[0]: n=20 Haar (seed 20) permanent -- the compiled reference C++
  (oracle/_ref) and the long-double arbiter;
* configs[2]: the three 60x60 UNFILTERED occupation patterns of SURVEY.md 8(d)
  (zeros included, as piquasso/_simulators/passive/utils.py:131-138 calls it) --
  compiled reference and long-double arbiter;
* configs[3]: the first 3 shots of the 100-mode / 25-photon Clifford-Clifford run
  (Haar seed 100, input [1]*25+[0]*75, seed_sequence 123) from the reference's own
  ``_generate_samples`` (piquasso/_simulators/passive/sampling.py:149-236) driven by
  the compiled reference ``permanent_laplace``;
* the two detection probabilities of the reference's
  tests/_simulators/passive/test_preparations.py:231-282 with the inputs that
  reproduce them through ``connector.permanent``.

configs[1] (n=30) lives in tests/golden/arbiter.json (make_arbiter_golden.py).
Nothing here is read at test time except the JSON it writes.
"""
import json
import os
import sys
from types import SimpleNamespace

import numpy as np
from scipy.stats import unitary_group

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, REPO)

import oracle  # noqa: E402
from make_golden import _install_stubs, _mat  # noqa: E402

sys.path.insert(0, os.path.join(REPO))
import bench_secondary  # noqa: E402  (the occupation patterns of configs[2])


def c(z):
    z = complex(z)
    return [z.real, z.imag]


def main():
    if not oracle.ref_available():
        raise SystemExit("oracle/_ref could not be built: no reference sources")
    _install_stubs()
    import piquasso as pq  # noqa: F401  (the reference package, /root/reference)
    from piquasso._simulators.passive import sampling as ref_sampling

    doc = {}
    u20 = unitary_group.rvs(20, random_state=20)
    ones = np.ones(20, dtype=np.int32)
    doc["cfg1_n20"] = {"n": 20, "seed": 20, "reference_cpp": c(oracle.ref_permanent(u20, ones, ones)),
                       "long_double": c(oracle.permanent(u20, ones, ones, precision=1))}

    u60 = unitary_group.rvs(60, random_state=60)
    cfg3 = {}
    for name, (rows, cols) in bench_secondary.cfg3_cases().items():
        rows, cols = rows.astype(np.int32), cols.astype(np.int32)
        cfg3[name] = {"rows": rows.tolist(), "cols": cols.tolist(),
                      "reference_cpp": c(oracle.ref_permanent(u60, rows, cols)),
                      "long_double": c(oracle.permanent(u60, rows, cols, precision=1))}
        print("cfg3", name, cfg3[name]["reference_cpp"], flush=True)
    doc["cfg3_60modes_24photons"] = {"haar_seed": 60, "cases": cfg3}

    u100 = unitary_group.rvs(100, random_state=100)
    inp = np.array([1] * 25 + [0] * 75)
    config = SimpleNamespace(seed_sequence=123, use_dask=False)
    samples = ref_sampling.generate_samples(inp, 3, oracle.ref_permanent_laplace, u100,
                                            lambda: False, ((), (), 1000), None, config)
    doc["cfg4_sampler_100modes_25photons"] = {
        "haar_seed": 100, "input": inp.tolist(), "seed_sequence": 123, "shots": 3,
        "samples": [[int(x) for x in s] for s in samples]}
    print("cfg4 first shots", doc["cfg4_sampler_100modes_25photons"]["samples"][0][:30], flush=True)

    # detection probabilities: run the reference's two tests and keep what they assert
    import pytest  # noqa: F401
    with pq.Program() as program:
        pq.Q(all) | pq.StateVector([1, 1, 1, 0, 0])
        pq.Q(all) | pq.Interferometer(unitary_group.rvs(5, random_state=42))
    sim = pq.SamplingSimulator(d=5, config=pq.Config(cutoff=4))
    state = sim.execute(program).state
    probe = [([1, 1, 1, 0, 0]), ([0, 2, 0, 1, 0]), ([3, 0, 0, 0, 0]), ([0, 0, 1, 1, 1])]
    doc["detection_probabilities"] = {
        "haar_seed": 42, "input": [1, 1, 1, 0, 0],
        "outputs": probe,
        "values": [float(state.get_particle_detection_probability(np.array(o))) for o in probe],
        "source": "pq.SamplingSimulator(d=5) state.get_particle_detection_probability on the reference "
                  "(piquasso/_simulators/passive/probabilities.py:26-54 -> connector.permanent)"}
    print("detection probabilities", doc["detection_probabilities"]["values"], flush=True)

    with open(os.path.join(HERE, "baseline_shapes.json"), "w") as fh:
        json.dump(doc, fh, indent=1)
    print("wrote baseline_shapes.json")


if __name__ == "__main__":
    main()
