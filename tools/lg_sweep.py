#!/usr/bin/env python
"""BASELINE config 4: LightGlue self/cross-attention sweep, N_kpts in {256, 512, 1024, 2048}, 1 x B200.

For every N the SURVEY 8(d) synthetic inputs (random integer keypoints, unit descriptors, permuted + noised second set)
are uploaded once into device slots (rfe_sp_write_slot) and PAIRS copies of the pair are matched per call
(rfe_lg_match_slots_batch), so the timed region holds no host<->device copy of descriptors.  Reported per N:
device time per pair (CUDA events around every kernel, rfe_profile), algorithmic TFLOP/s of the whole matcher and of the
attention kernel alone, and both as a fraction of the measured bf16 peak (split-fp16 runs 3 MMAs per algorithmic MAC, 3.5
for attention, so the fractions are bounded by 1/3 and 1/3.5).  One JSON line per N.
Usage: python tools/lg_sweep.py [pairs_per_call=8] [reps=10]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import bench
from oracle import synth          # input generator only
from rover_slam_b200 import FrontEnd

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
peak, _, peak_src = bench.peaks()
fe = FrontEnd(max_batch=2 * pairs, max_height=480, max_width=640, max_keypoints=2048)
for n in (256, 512, 1024, 2048):
    k0, k1, d0, d1, perm = synth.lightglue_inputs(n, 200 + n)
    for p in range(pairs):
        fe.write_slot(2 * p, k0, d0)
        fe.write_slot(2 * p + 1, k1, d1)
    s0, s1 = list(range(0, 2 * pairs, 2)), list(range(1, 2 * pairs, 2))
    for _ in range(3):
        fe.match_slots_batch(s0, s1, 480, 640, 0.0)
    fe.sync()
    fe.profile(True)
    fe.profile_read(None, reset=True)
    for _ in range(reps):
        fe.match_slots_batch(s0, s1, 480, 640, 0.0)
    fe.sync()
    tot_ms, launches = fe.profile_read(None)
    attn_ms, attn_n = fe.profile_read("lg.attn")
    fe.profile(False)
    m, ms = fe.read_result(0)
    inv = np.empty(n, np.int64)
    inv[perm] = np.arange(n)                # k1[j] = k0[perm[j]]  ->  true partner of query i is inv[i]
    correct = int((inv[m[:, 0]] == m[:, 1]).sum())
    flops = 9 * (4980736 * n + 4096 * n * n) + 263168 * n + 512 * n * n          # SURVEY.md 8(d), per pair
    attn_flops = 9 * 4096 * n * n
    per_pair_ms = tot_ms / reps / pairs
    line = {"config": "BASELINE config 4: LightGlue N sweep", "n_kpts": n, "pairs_per_call": pairs, "reps": reps,
            "ms_per_pair": per_pair_ms, "pairs_per_s": 1e3 / per_pair_ms, "kernel_launches_per_call": launches / reps,
            "matcher_tflops_algorithmic": flops / per_pair_ms / 1e9, "matcher_frac_of_bf16_peak": flops / per_pair_ms / 1e9 / peak,
            "attn_ms_per_pair": attn_ms / reps / pairs, "attn_tflops_algorithmic": attn_flops / (attn_ms / reps / pairs) / 1e9,
            "attn_frac_of_bf16_peak": attn_flops / (attn_ms / reps / pairs) / 1e9 / peak, "attn_share": attn_ms / tot_ms,
            "peak_tflops": peak, "peak_source": peak_src, "matches": int(len(m)), "matches_correct": correct}
    print(json.dumps(line), flush=True)
