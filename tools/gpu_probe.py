#!/usr/bin/env python
"""Run the tcgen05 hardware probes and print what they say (results recorded in DESIGN.md)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from rover_slam_b200 import FrontEnd

fe = FrontEnd(max_batch=1, max_height=64, max_width=64, max_keypoints=256)
rng = np.random.RandomState(0)
a = rng.randn(136, 64).astype(np.float16).astype(np.float32)
b = rng.randn(64, 64).astype(np.float16).astype(np.float32)
out = fe.debug_probe_shift(a, b)
print("probe 0: row-shifted SWIZZLE_128B K-major A operand; max-abs error vs A[s:s+128] @ B^T")
print("  s   base_offset=0   base_offset=s&7   base_offset=(8-s)&7")
for s in range(9):
    ref = a[s:s + 128].astype(np.float64) @ b.astype(np.float64).T
    errs = [np.abs(out[s, m] - ref).max() for m in range(3)]
    print(f"  {s}   " + "   ".join(f"{e:12.3e}" for e in errs))
