#!/usr/bin/env python
"""Per-CTA timeline of one attention launch at the bench shape (8 pairs): %globaltimer at CTA entry / exit and the SM id.
Prints the launch span, CTA duration statistics per wave position and the idle time between consecutive CTAs of an SM."""
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
dur = end - start
print(f"SM clock while the kernel runs: {np.median(cyc / np.maximum(end - start, 1)) * 1e3:.0f} MHz (clock64 / globaltimer per CTA, median); cycles per CTA p50 {np.median(cyc):.0f}")
print(f"CTAs {len(tl)}  SMs {len(np.unique(sm))}  launch span {end.max() / 1e3:.1f} us")
print(f"CTA duration us: mean {dur.mean() / 1e3:.2f}  min {dur.min() / 1e3:.2f}  p50 {np.median(dur) / 1e3:.2f}  max {dur.max() / 1e3:.2f}")
gaps, per_sm_busy, order_d = [], [], {}
for s in np.unique(sm):
    idx = np.where(sm == s)[0]
    idx = idx[np.argsort(start[idx])]
    per_sm_busy.append(dur[idx].sum())
    for k, i in enumerate(idx):
        order_d.setdefault(k, []).append(dur[i])
    gaps += list(start[idx][1:] - end[idx][:-1])
gaps = np.array(gaps)
print(f"idle between consecutive CTAs on one SM us: mean {gaps.mean() / 1e3:.2f}  p50 {np.median(gaps) / 1e3:.2f}  max {gaps.max() / 1e3:.2f}")
print(f"per-SM busy us: mean {np.mean(per_sm_busy) / 1e3:.1f}  min {np.min(per_sm_busy) / 1e3:.1f}  max {np.max(per_sm_busy) / 1e3:.1f}")
print("mean CTA duration by position on its SM:", " ".join(f"{k}:{np.mean(v) / 1e3:.1f}us(n={len(v)})" for k, v in sorted(order_d.items())))
print(f"first CTA start spread us: {np.sort(start)[min(147, len(start) - 1)] / 1e3:.2f}; last CTA end {end.max() / 1e3:.1f}, earliest last-end {min(end[sm == s].max() for s in np.unique(sm)) / 1e3:.1f}")
names = ["wait Q", "pass1 issue loop", "pass2 issue loop (scores)", "wait K", "wait free S buffer", "wait V", "wait P (softmax)", "key tiles",
         "pass1 wait K", "pass1 wait free S", "softmax w0 pass-2 loop", "softmax w0 wait scores", "softmax w0 wait free P", "softmax w0 pass-1 loop",
         "softmax w0 pass-1 wait scores", "softmax w0 tcgen05.ld"]
print("role counters of CTA", os.environ.get("RFE_ATTN_PROF_CTA", "0"), "(cycles):", ", ".join(f"{n} {int(v)}" for n, v in zip(names, pr[:16])))
