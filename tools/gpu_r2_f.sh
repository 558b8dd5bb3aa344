#!/bin/bash
# A/B of the attention tail split (RFE_ATTN_SPLIT) + the tests that cover it
mkdir -p gpurun_out
ROVER_FE_LIB=$PWD/rover_slam_b200/librover_fe_dbg.so timeout 300 python -m pytest tests -m gpu -x -q -k "key_split or bench_shape" 2>&1 | tail -4
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "failed under the bounded-wait library: stopping"; exit 0; fi
timeout 400 python -m pytest tests -m gpu -x -q -k "key_split or bench_shape or lightglue or one_to_many" 2>&1 | tail -4
for rep in 1 2; do
  for m in 0 1; do
    RFE_ATTN_SPLIT=$m timeout 200 python bench.py --steps 20 --warmup 5 --cpu-pairs 0 > gpurun_out/r02_split_$m.json 2> gpurun_out/r02_split_$m.err || { tail -3 gpurun_out/r02_split_$m.err; continue; }
    python - $m <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/r02_split_{sys.argv[1]}.json"))
print("split", sys.argv[1], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "attn us/step", d["kernel_us_per_step"]["lg.attn"], "frac", round(d["roofline"]["frac"], 4), "clk", d["clocks"]["sm_mhz"], "launches", d["gpu_launches"])
PY
  done
done
