"""Build recipe for the checker libraries (TEST INFRASTRUCTURE, not product).

* ``oracle/libpqoracle.so``  -- our C restatement (oracle/perm_oracle.c), gcc.
* ``oracle/_ref/libpqref.so`` -- the UNMODIFIED reference C++
  (``/root/reference/src/permanent.cpp`` + ``permanent_laplace.cpp``) compiled
  where it lies together with ``oracle/ref_shim.cpp``.  Only attempted when the
  reference tree is present (the build container); on the GPU box the prebuilt
  file that travelled with the snapshot is used.  The reference's own build
  system (CMake/scikit-build) is not run: the path is two translation units
  with header-only dependencies.

``/usr/bin/gcc`` is used explicitly: the image's default ``$CC``
(``/opt/gcc/bin/gcc``) lacks ``libgomp.spec`` and silently loses OpenMP.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_SRC = os.environ.get("PQ_REFERENCE_SRC", "/root/reference/src")

ORACLE_SO = os.path.join(HERE, "libpqoracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_SO = os.path.join(REF_DIR, "libpqref.so")


def _cc(name: str) -> str:
    path = os.path.join("/usr/bin", name)
    return path if os.path.exists(path) else shutil.which(name) or name


def _newer(target: str, *sources: str) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources if os.path.exists(s))


def _run(cmd):
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(
            "command failed: %s\n%s\n%s" % (" ".join(cmd), proc.stdout, proc.stderr)
        )


def build_oracle(force: bool = False) -> str:
    src = os.path.join(HERE, "perm_oracle.c")
    if not force and _newer(ORACLE_SO, src):
        return ORACLE_SO
    _run(
        [_cc("gcc"), "-O2", "-std=c11", "-fPIC", "-shared", "-fopenmp",
         "-Wall", "-Wextra", "-o", ORACLE_SO, src, "-lm"]
    )
    return ORACLE_SO


def reference_sources_present() -> bool:
    return os.path.exists(os.path.join(REFERENCE_SRC, "permanent.cpp"))


def build_ref(force: bool = False) -> str | None:
    """Compile the reference path into oracle/_ref/libpqref.so.

    Returns the path, or None when neither the sources nor a prebuilt file
    exist (then callers must treat the reference arm as unavailable)."""
    shim = os.path.join(HERE, "ref_shim.cpp")
    if not reference_sources_present():
        return REF_SO if os.path.exists(REF_SO) else None
    srcs = [
        os.path.join(REFERENCE_SRC, "permanent.cpp"),
        os.path.join(REFERENCE_SRC, "permanent_laplace.cpp"),
        shim,
    ]
    if not force and _newer(REF_SO, *srcs):
        return REF_SO
    os.makedirs(REF_DIR, exist_ok=True)
    # Flags follow the reference's Release build: C++17, -O3, OpenMP
    # (CMakeLists.txt:9-22, src/CMakeLists.txt:55-59).
    _run(
        [_cc("g++"), "-O3", "-std=c++17", "-fPIC", "-shared", "-fopenmp",
         "-DNDEBUG", "-I", REFERENCE_SRC, "-o", REF_SO] + srcs
    )
    return REF_SO


def main() -> None:
    force = "--force" in sys.argv
    print("oracle:", build_oracle(force))
    print("ref   :", build_ref(force))


if __name__ == "__main__":
    main()
