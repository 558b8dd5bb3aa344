#!/bin/bash
# compute-sanitizer over the final build: smoke() and a LightGlue batch whose attention launch splits its tail items
mkdir -p gpurun_out
cat > /tmp/split_case.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
from oracle import synth
from rover_slam_b200 import FrontEnd
fe = FrontEnd(max_batch=4, max_height=480, max_width=640, max_keypoints=1280)
k0, k1, d0, d1, perm = synth.lightglue_inputs(1250, 2150)
for p in range(2):
    fe.write_slot(2 * p, k0, d0); fe.write_slot(2 * p + 1, k1, d1)
fe.match_slots_batch([0, 2], [1, 3], 480, 640)
m, s = fe.read_result(0)
print("matches", len(m), "correct", float((perm[m[:, 1]] == m[:, 0]).mean()))
PY
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_memcheck_smoke.txt 2>&1
tail -3 gpurun_out/r02_sanitizer_memcheck_smoke.txt
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python /tmp/split_case.py > gpurun_out/r02_sanitizer_memcheck_split.txt 2>&1
tail -3 gpurun_out/r02_sanitizer_memcheck_split.txt
timeout 1200 compute-sanitizer --tool racecheck --print-limit 5 python /tmp/split_case.py > gpurun_out/r02_sanitizer_racecheck_split.txt 2>&1
tail -3 gpurun_out/r02_sanitizer_racecheck_split.txt
