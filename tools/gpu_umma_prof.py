#!/usr/bin/env python
"""Role-level cycle breakdown of one tagged UMMA launch (RFE_PROF_TAG) inside a 4-pair LightGlue batch."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from oracle import synth
from rover_slam_b200 import FrontEnd
fe = FrontEnd(max_batch=8, max_height=480, max_width=640, max_keypoints=4096)
fe.debug_read("lg.attn_prof")
imgs = np.stack([f for s in range(4) for f in synth.frame_pair(s, 480, 640, shift=(5, 3))])
fe.extract_device_from_host(imgs)
fe.match_slots_batch([0, 2, 4, 6], [1, 3, 5, 7], 480, 640)
fe.sync()
pr = fe.debug_read("lg.attn_prof").view(np.uint64)
print("tag", os.environ.get("RFE_PROF_TAG"), ": MMA thread total %d, wait TMEM drain %d, wait TMA %d, tiles %d | epilogue warp total %d, wait accum %d"
      % tuple(int(v) for v in pr[16:22]))
names = ["other", "tmem ld", "bias/scale/rotary", "residual wait+add", "f32 stage+store", "split", "wait staging free", "stage+fence+TMA store"]
ph = [int(v) for v in pr[24:32]]
if any(ph):
    tiles = max(int(pr[19]), 1)
    print("  LINEAR epilogue phases of warp 0 (cycles, all tiles of CTA 0; per chunk = / (2 chunks x tiles)):")
    for n, v in zip(names, ph):
        print(f"    {n:26s} {v:9d}  ({v / (2 * tiles):7.0f} per chunk)")
