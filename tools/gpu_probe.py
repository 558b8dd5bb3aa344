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

import ctypes as C
o = np.zeros(16, np.float32)
fe._check(fe.lib.rfe_debug_probe(fe.ctx, 1, None, None, o.ctypes.data_as(C.c_void_p)))
print("probe 1: cycles per tcgen05.mma (M=128, K=16, SS operands), 512 back-to-back, one CTA")
print("  same A tile      : N=64 %.1f  N=128 %.1f  N=256 %.1f" % tuple(o[0:3]))
print("  alternating A    : N=64 %.1f  N=128 %.1f  N=256 %.1f" % tuple(o[3:6]))
print("  issue-loop only  : N=64 %.1f cycles per instruction issued" % o[6])
print("  A start +128 B   : N=128 %.1f   A start +256 B: N=128 %.1f" % (o[8], o[9]))

o3 = np.zeros(16, np.float32)
fe._check(fe.lib.rfe_debug_probe(fe.ctx, 3, None, None, o3.ctypes.data_as(C.c_void_p)))
print("probe 3: softmax-role limits (640-thread CTA, one SM)")
print("  cycles per 16 KB drained from TMEM (tcgen05.ld.32x32b.x32): 4 warps %.1f  8 warps %.1f  16 warps %.1f" % tuple(o3[0:3]))
print("  16 warps + concurrent N=128 MMAs: %.1f per 16 KB, %.1f cycles per MMA" % (o3[3], o3[4]))
print("  cycles per warp-instruction per scheduler: ex2 %.2f  cvt.f16x2.f32 %.2f  cvt.f32.f16 %.2f  fma.f32x2 %.2f" % tuple(o3[5:9]))

# attention role timing (cycles) on a 2000-keypoint pair
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from oracle import synth
fe2 = FrontEnd(max_batch=2, max_height=64, max_width=64, max_keypoints=2048)
k0, k1, d0, d1, _ = synth.lightglue_inputs(2000, 5)
fe2.debug_read("lg.attn_prof")          # arm
fe2.match(k0, k1, d0, d1, 480, 640)
pr = fe2.debug_read("lg.attn_prof").view(np.uint64)
names = ["wait Q", "pass1 issue loop", "pass2 issue loop", "wait K", "wait free S buffer", "wait V", "wait P (softmax)", "key tiles"]
print("attention MMA-thread cycle breakdown (CTA 0 of the last launch):")
for n, v in zip(names, pr[:8]): print(f"  {n:22s} {int(v)}")
print(f"  pass 1 only: wait K {int(pr[8])}, wait free S buffer {int(pr[9])}")
print(f"  softmax warp 0, pass 2: loop {int(pr[10])}, wait scores {int(pr[11])}, wait free P buffer {int(pr[12])}, tcgen05.ld {int(pr[15])}")
print(f"  softmax warp 0, pass 1: loop {int(pr[13])}, wait scores {int(pr[14])}")

# strip conv (conv1b) MMA-thread breakdown on one 480x640 frame batch of 8
fe3 = FrontEnd(max_batch=8, max_height=480, max_width=640, max_keypoints=4096)
fe3.debug_read("lg.attn_prof")          # arm (shared counter buffer; conv1b uses slots 8..)
imgs = np.stack([synth.frame(s, 480, 640) for s in range(8)])
fe3.extract(imgs, want_desc=False)
pr = fe3.debug_read("lg.attn_prof").view(np.uint64)
print("strip conv1b MMA-thread cycles (CTA 0): total %d, wait TMEM drain %d, wait weights %d, wait rows %d, iterations %d" % tuple(int(v) for v in pr[8:13]))
