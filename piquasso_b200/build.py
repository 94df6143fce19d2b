"""In-tree build of ``piquasso_b200/libpqperm.so`` (CUDA, sm_100a only).

``python -m piquasso_b200.build`` or ``__graft_entry__.build()``.  nvcc
cross-compiles without a GPU; the resulting ``.so`` is git-ignored but travels
to the GPU box with the gpurun snapshot.  CMakeLists.txt at the repo root
builds the same targets for a piquasso checkout.
"""

from __future__ import annotations

import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libpqperm.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
                     "-Xcompiler", "-Wall", "-Xcompiler", "-Wextra"]
# extra flags for experiments, e.g. PQ_EXTRA_NVCC_FLAGS="-DPQ_SOMETHING=1"
NVCC_FLAGS += os.environ.get("PQ_EXTRA_NVCC_FLAGS", "").split()

# column ranges of the binary constant-bank kernel, one translation unit each
BINARY_PARTS = [(0, 8, 20), (1, 21, 30), (2, 31, 40), (3, 41, 48)]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libpqperm.so cannot be built")


def _units():
    units = [
        ("api", "pqperm_api.cu", []),
        ("api_laplace", "pqperm_api_laplace.cu", []),
        ("plan", "pqperm_plan.cpp", []),
        ("rng", "pqperm_rng.cpp", []),
        ("generic", "pqperm_kernels_generic.cu", []),
        ("permhyper", "pqperm_kernels_permhyper.cu", []),
        # every FMA of the double-double arithmetic is explicit: no contraction
        ("arbiter", "pqperm_arbiter.cu", ["-fmad=false"]),
    ]
    # the batched Laplace walk: unit-column / general flavour x accumulation mode
    for unit in (1, 0):
        for mode in (0, 1, 2):
            units.append(("laplace_u%d_m%d" % (unit, mode), "pqperm_kernels_laplace.cu",
                          ["-DPQ_LAP_UNIT=%d" % unit, "-DPQ_LAP_MODE=%d" % mode]))
        # batched permanents: one lane per segment for 9..20 / 21..32 columns
        for part in (1, 2):
            units.append(("laplace_u%d_m2_p%d" % (unit, part), "pqperm_kernels_laplace.cu",
                          ["-DPQ_LAP_UNIT=%d" % unit, "-DPQ_LAP_MODE=2", "-DPQ_LAP_PART=%d" % part]))
    for part, lo, hi in BINARY_PARTS:
        units.append((
            "binary%d" % part, "pqperm_kernels_binary.cu",
            ["-DPQ_BIN_PART=%d" % part, "-DPQ_BIN_LO=%d" % lo, "-DPQ_BIN_HI=%d" % hi],
        ))
    return units


def _headers_digest() -> str:
    h = hashlib.sha256()
    names = sorted(f for f in os.listdir(CSRC) if f.endswith((".h", ".cuh")))
    for f in names:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    with open(os.path.join(HERE, "..", "include", "pqperm.h"), "rb") as fh:
        h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(unit, digest, force):
    name, src, defs = unit
    src_path = os.path.join(CSRC, src)
    obj = os.path.join(OBJ, name + ".o")
    stamp = obj + ".stamp"
    with open(src_path, "rb") as fh:
        want = hashlib.sha256(fh.read() + digest.encode() + " ".join(defs).encode()).hexdigest()
    if not force and os.path.exists(obj) and os.path.exists(stamp):
        with open(stamp) as fh:
            if fh.read().strip() == want:
                return obj, False
    cmd = [_nvcc()] + NVCC_FLAGS + defs + ["-c", src_path, "-o", obj]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, proc.stdout, proc.stderr))
    with open(stamp, "w") as fh:
        fh.write(want)
    return obj, True


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    digest = _headers_digest()
    units = _units()
    workers = max(1, min(len(units), os.cpu_count() or 1))
    with concurrent.futures.ThreadPoolExecutor(workers) as pool:
        results = list(pool.map(lambda u: _compile(u, digest, force), units))
    objs = [r[0] for r in results]
    rebuilt = any(r[1] for r in results)
    if rebuilt or not os.path.exists(LIB):
        cmd = [_nvcc()] + ARCH + ["-shared", "-o", LIB] + objs
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (proc.stdout, proc.stderr))
    if verbose:
        print("libpqperm:", LIB, "(rebuilt)" if rebuilt else "(up to date)")
    return LIB


def build_pybind(force: bool = False, verbose: bool = False) -> str:
    """The pybind11 drop-in module ``permanent`` (same name and overloads as the
    reference's ``piquasso/_math/permanent*.so``), linked against libpqperm.so
    with an $ORIGIN rpath; lands in ``piquasso_b200/native/``."""
    import sysconfig

    import pybind11

    build(force=force, verbose=verbose)
    out_dir = os.path.join(HERE, "native")
    os.makedirs(out_dir, exist_ok=True)
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    target = os.path.join(out_dir, "permanent" + suffix)
    src = os.path.join(CSRC, "pybind_permanent.cpp")
    hdr = os.path.join(HERE, "..", "include", "pqperm.h")
    if (not force and os.path.exists(target)
            and os.path.getmtime(target) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return target
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++")
    cmd = [gxx, "-O2", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden",
           "-I", pybind11.get_include(), "-I", sysconfig.get_paths()["include"],
           src, "-o", target, "-L", HERE, "-lpqperm", "-Wl,-rpath,$ORIGIN/.."]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("pybind build failed:\n%s\n%s" % (proc.stdout, proc.stderr))
    if verbose:
        print("pybind module:", target)
    return target


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    build_pybind(force="--force" in sys.argv, verbose=True)
