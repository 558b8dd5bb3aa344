#!/usr/bin/env python
"""Per-kernel-tag CUDA-event times of the bench step (8 frames extracted + 4 pairs matched), warm caches.
Usage: python tools/gpu_tagprof.py [steps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import bench
from rover_slam_b200 import FrontEnd

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
arm = len(sys.argv) > 2 and sys.argv[2] == "arm"      # arm the in-kernel role counters (attention runs its PROF build)
P = int(os.environ.get("RFE_PAIRS", "4"))
B = 2 * P
fe = FrontEnd(max_batch=B, max_height=bench.H, max_width=bench.W, max_keypoints=4096)
if arm:
    fe.debug_read("lg.attn_prof")
frames = torch.from_numpy(bench.make_pairs(P, 1).reshape(B, bench.H, bench.W)).cuda()
sa, sb = list(range(0, B, 2)), list(range(1, B, 2))
def step():
    fe.extract_device(frames.data_ptr(), bench.H, bench.W, bench.W, B)
    fe.match_slots_batch(sa, sb, bench.H, bench.W, 0.0)
for _ in range(3):
    step()
fe.sync()
fe.profile(True)
for _ in range(steps):
    step()
fe.sync()
tags = ["sp.conv1a", "sp.conv1b", "sp.conv2a", "sp.conv2b", "sp.conv3a", "sp.conv3b", "sp.conv4a", "sp.conv4b", "sp.convPa", "sp.convDa",
        "sp.convPb_softmax", "sp.convDb_l2norm", "sp.nms", "sp.select", "sp.desc_sample", "lg.prepare",
        "lg.wqkv_rope", "lg.attn_self", "lg.out_proj", "lg.ffn0", "lg.ln_gelu", "lg.ffn3", "lg.to_qk", "lg.to_v", "lg.attn_cross", "lg.to_out",
        "lg.final_proj", "lg.matchability", "lg.sim", "lg.assign"]
tot, _ = fe.profile_read(None)
print(f"{'tag':22s} {'launches/step':>13s} {'us/launch':>10s} {'us/step':>10s} {'share':>7s}")
acc = 0.0
for t in tags:
    ms, n = fe.profile_read(t)
    if n:
        acc += ms
        print(f"{t:22s} {n / steps:13.1f} {1e3 * ms / n:10.1f} {1e3 * ms / steps:10.1f} {100 * ms / tot:6.1f}%")
print(f"{'sum of kernels':22s} {'':13s} {'':10s} {1e3 * tot / steps:10.1f}   (tagged {1e3 * acc / steps:.1f})")
