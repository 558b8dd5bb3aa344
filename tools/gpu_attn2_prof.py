#!/usr/bin/env python
"""Role counters and per-CTA timeline of ONE persistent attention launch (attn2_kernel) at the bench shape (8 pairs)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import bench
from rover_slam_b200 import FrontEnd

P = int(os.environ.get("RFE_PAIRS", "8"))
B = 2 * P
fe = FrontEnd(max_batch=B, max_height=bench.H, max_width=bench.W, max_keypoints=4096)
frames = torch.from_numpy(bench.make_pairs(P, 1).reshape(B, bench.H, bench.W)).cuda()
sa, sb = list(range(0, B, 2)), list(range(1, B, 2))
fe.extract_device(frames.data_ptr(), bench.H, bench.W, bench.W, B)
for _ in range(2):
    fe.match_slots_batch(sa, sb, bench.H, bench.W, 0.0)
fe.sync()
fe.debug_read("lg.attn_prof")          # arm: the attention launches now run the PROF instantiation
fe.match_slots_batch(sa, sb, bench.H, bench.W, 0.0)
fe.sync()
pr = fe.debug_read("lg.attn_prof").view(np.uint64)
tl = pr[32:].reshape(-1, 3).astype(np.int64)
tl = tl[tl[:, 0] > 0]
t0 = tl[:, 0].min()
start, end, sm, cyc = tl[:, 0] - t0, tl[:, 1] - t0, tl[:, 2] & 0xFFFF, tl[:, 2] >> 16
print(f"CTAs {len(tl)}  SMs {len(np.unique(sm))}  launch span {end.max() / 1e3:.1f} us; CTA cycles p50 {np.median(cyc):.0f} min {cyc.min()} max {cyc.max()}"
      f"; clock {np.median(cyc / np.maximum(end - start, 1)) * 1e3:.0f} MHz")
items, tiles = int(pr[17]), int(pr[7])
print(f"CTA {os.environ.get('RFE_ATTN_PROF_CTA', '0')}: {items} items, {tiles} pass-2 key tiles")
mode = os.environ.get("RFE_ATTN", "3")
if mode == "3":
    rows = (("score issuer: wait Q", 0), ("score issuer: issue loops", 2), ("  wait K", 3), ("  wait group's score buffer", 4),
            ("PV issuer: wait V", 5), ("PV issuer: wait P", 6), ("PV issuer: wait O hand-back", 16),
            ("softmax w0: key-tile loops", 10), ("  wait scores", 11), ("  wait free P", 12), ("  pair-max exchange", 13),
            ("  O corrections (count)", 14), ("softmax w0: merge + epilogue", 15), ("  wait o_full", 18))
else:
    rows = (("score issuer: wait Q", 0), ("score issuer: pass-1 loops", 1), ("score issuer: pass-2 loops", 2), ("  pass 2 wait K", 3),
            ("  pass 2 wait free S", 4), ("  pass 1 wait K", 8), ("  pass 1 wait free S", 9), ("PV issuer: wait V", 5), ("PV issuer: wait P", 6),
            ("PV issuer: wait O hand-back", 16), ("softmax w0: pass-1 loops", 13), ("  wait scores", 14), ("softmax w0: pass-2 loops", 10),
            ("  wait scores", 11), ("  wait free P", 12), ("  phase: tmem ld + release", 19), ("  phase: exp / split math", 20), ("  phase: P stores", 21),
            ("  phase: fence + arrive", 22), ("softmax w0: l exchange + epilogue", 15), ("  wait o_full", 18))
for name, i in rows:
    print(f"  {name:36s} {int(pr[i]):9d}  ({int(pr[i]) / max(items, 1):9.0f} per item)")
