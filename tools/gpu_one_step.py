#!/usr/bin/env python
"""ONE bench step (16 extracts + 8 matches at 640x480) between cudaProfilerStart / Stop, for `ncu --profile-from-start off`.
  python tools/gpu_one_step.py [pairs]"""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import bench
from rover_slam_b200 import FrontEnd

P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
B = 2 * P
fe = FrontEnd(max_batch=B, max_height=bench.H, max_width=bench.W, max_keypoints=4096)
frames = torch.from_numpy(bench.make_pairs(P, 1).reshape(B, bench.H, bench.W)).cuda()
sa, sb = list(range(0, B, 2)), list(range(1, B, 2))
for _ in range(2):
    fe.extract_device(frames.data_ptr(), bench.H, bench.W, bench.W, B)
    fe.match_slots_batch(sa, sb, bench.H, bench.W, 0.0)
fe.sync()
torch.cuda.cudart().cudaProfilerStart()
fe.extract_device(frames.data_ptr(), bench.H, bench.W, bench.W, B)
fe.match_slots_batch(sa, sb, bench.H, bench.W, 0.0)
fe.sync()
torch.cuda.cudart().cudaProfilerStop()
print("keypoints", [len(fe.read_slot(b, want_desc=False)[0]) for b in range(B)], "matches", [len(fe.read_result(p)[0]) for p in range(P)])
