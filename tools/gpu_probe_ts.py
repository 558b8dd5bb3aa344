#!/usr/bin/env python
"""Probe 2: tcgen05.mma with the A operand in TMEM (TS mode): layout check + issue rate."""
import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from rover_slam_b200 import FrontEnd

fe = FrontEnd(max_batch=1, max_height=64, max_width=64, max_keypoints=256)
rng = np.random.RandomState(0)
a = rng.randn(128, 64).astype(np.float16).astype(np.float32)
b = rng.randn(64, 64).astype(np.float16).astype(np.float32)
out = np.zeros(8192 + 8, np.float32)
fe._check(fe.lib.rfe_debug_probe(fe.ctx, 2, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)))
d = out[:8192].reshape(128, 64)
ref = a.astype(np.float64) @ b.astype(np.float64).T
print("probe 2: A from TMEM (thread t = lane t, column j = half2(k=2j, 2j+1)); max-abs error vs A @ B^T: %.3e" % np.abs(d - ref).max())
print("  cycles per MMA: TS N=64 %.1f | TS N=128 %.1f | pair (SS N=128 + TS N=64) %.1f" % tuple(out[8192:8195]))
