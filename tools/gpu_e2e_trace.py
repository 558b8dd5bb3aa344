#!/usr/bin/env python
"""Where the pipelined host API (rfe_pairs_submit / collect_begin / collect_end) spends host time, per call."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import bench
from rover_slam_b200 import FrontEnd

P = int(os.environ.get("RFE_PAIRS", "8")); B = 2 * P; steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
fe = FrontEnd(max_batch=B, max_height=bench.H, max_width=bench.W, max_keypoints=4096)
sets = [torch.from_numpy(bench.make_pairs(P, 1 + i).reshape(B, bench.H, bench.W)).pin_memory().numpy() for i in range(2)]
def run(n, trace=None):
    fe.pairs_submit(sets[0]); fe.pairs_submit(sets[1])
    for i in range(n):
        t0 = time.perf_counter(); fe.pairs_collect_begin(); t1 = time.perf_counter()
        if i + 2 < n: fe.pairs_submit(sets[i % 2])
        t2 = time.perf_counter(); fe.pairs_collect_end(); t3 = time.perf_counter()
        if trace is not None: trace.append((t1 - t0, t2 - t1, t3 - t2))
run(3); fe.sync()
tr = []; t = time.perf_counter(); run(steps, tr); fe.sync(); dt = time.perf_counter() - t
print(f"e2e {2 * P * steps / dt:.1f} frames/s, {1e3 * dt / steps:.2f} ms per step")
a = np.array(tr) * 1e3
print("ms per call (median): collect_begin %.2f  submit %.2f  collect_end %.2f" % tuple(np.median(a, 0)))
print("first 4 steps:", np.round(a[:4], 2).tolist())
d = torch.from_numpy(np.concatenate(sets)).cuda().reshape(2, B, bench.H, bench.W)
sa, sb = list(range(0, B, 2)), list(range(1, B, 2))
def dev(i):
    fe.extract_device(d[i % 2].data_ptr(), bench.H, bench.W, bench.W, B); fe.match_slots_batch(sa, sb, bench.H, bench.W, 0.0)
for i in range(3): dev(i)
fe.sync(); t = time.perf_counter()
for i in range(steps): dev(i)
fe.sync(); dt = time.perf_counter() - t
print(f"device-resident {2 * P * steps / dt:.1f} frames/s, {1e3 * dt / steps:.2f} ms per step")
