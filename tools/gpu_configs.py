#!/usr/bin/env python
"""Timing lines for BASELINE configs 1-3 (config 4: tools/lg_sweep.py, config 5: bench.py): host API, every copy inside.
  1. one 640x480 frame, SuperPoint           (rfe_sp_extract_u8, batch 1)
  2. eight 640x480 frames, SuperPoint        (rfe_sp_extract_u8, batch 8)
  3. one 752x480 pair, SuperPoint + LightGlue (rfe_match_pairs_u8: 2 extracts + 1 match, keypoints + matches back)
Each: median of 20 calls after 3 warm-ups, plus the CPU oracle's time for the same input (torch-CPU restatement, all host
threads).  One JSON line per config."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from oracle import lightglue_ref, superpoint_ref, synth
from rover_slam_b200 import FrontEnd

def med(f, n=20, warm=3):
    for _ in range(warm):
        f()
    ts = []
    for _ in range(n):
        t = time.perf_counter(); f(); ts.append(time.perf_counter() - t)
    return 1e3 * float(np.median(ts))

fe = FrontEnd(max_batch=8, max_height=480, max_width=752, max_keypoints=4096)
sp, lg = superpoint_ref.SuperPointRef(), lightglue_ref.LightGlueRef()
torch.set_num_threads(len(os.sched_getaffinity(0)))
img = synth.frame(0, 480, 640)
n1 = len(fe.extract(img)[0][0])
t = time.perf_counter(); sp(img); cpu1 = 1e3 * (time.perf_counter() - t)
print(json.dumps({"config": 1, "what": "one 640x480 frame, SuperPoint, host in / host out", "keypoints": n1, "b200_ms": med(lambda: fe.extract(img)),
                  "cpu_oracle_ms": cpu1, "cpu_threads": torch.get_num_threads()}), flush=True)
imgs = np.stack([synth.frame(s, 480, 640) for s in range(8)])
ms8 = med(lambda: fe.extract(imgs))
print(json.dumps({"config": 2, "what": "eight 640x480 frames, SuperPoint, host in / host out (descriptors included)", "b200_ms": ms8,
                  "frames_per_s": 8e3 / ms8, "cpu_oracle_ms": 8 * cpu1}), flush=True)
ms8n = med(lambda: fe.extract(imgs, want_desc=False))
print(json.dumps({"config": 2, "what": "eight 640x480 frames, SuperPoint, keypoints + scores back only", "b200_ms": ms8n, "frames_per_s": 8e3 / ms8n}), flush=True)
a, b = synth.frame_pair(100, 480, 752)
pair = np.stack([a, b])
kp, res = fe.match_pairs(pair)
t = time.perf_counter()
ka, _, da = sp(a); kb, _, db = sp(b)
lg(lightglue_ref.normalize_keypoints(ka.numpy(), 480, 752), lightglue_ref.normalize_keypoints(kb.numpy(), 480, 752), da, db)
cpu3 = 1e3 * (time.perf_counter() - t)
print(json.dumps({"config": 3, "what": "one 752x480 pair: 2 extracts + 1 match, host frames in, keypoints + matches out", "keypoints": [len(kp[0]), len(kp[1])],
                  "matches": len(res[0][0]), "b200_ms": med(lambda: fe.match_pairs(pair)), "cpu_oracle_ms": cpu3}), flush=True)
