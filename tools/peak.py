"""python tools/peak.py: the library's measured DFMA peak (roofline denominator)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from piquasso_b200 import _lib
lib = _lib.load()
for it in (1 << 15, 1 << 17, 1 << 17):
    print("pq_fp64_peak_tflops(iters=%d) = %.2f TFLOP/s" % (it, lib.pq_fp64_peak_tflops(0, it)))
