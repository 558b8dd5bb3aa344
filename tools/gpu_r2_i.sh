#!/bin/bash
# A/B: two score buffers (RFE_ATTN_CFG=4) against three (0)
mkdir -p gpurun_out
K="lightglue or bench_shape or key_split"
ROVER_FE_LIB=$PWD/rover_slam_b200/librover_fe_dbg.so RFE_ATTN_CFG=4 timeout 240 python -m pytest tests -m gpu -x -q -k "$K" 2>&1 | tail -3
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "cfg 4 failed: stopping"; exit 0; fi
RFE_ATTN_CFG=4 RFE_ATTN=2 timeout 120 python tools/gpu_attn2_prof.py > gpurun_out/r02_attn2_cfg4.txt 2>&1; grep -E 'launch span|loops|wait|phase' gpurun_out/r02_attn2_cfg4.txt | head -24
for rep in 1 2; do
  for cfg in 0 4; do
    RFE_ATTN_CFG=$cfg timeout 150 python bench.py --steps 20 --warmup 5 --cpu-pairs 0 > gpurun_out/r02_cfg_$cfg.json 2> gpurun_out/r02_cfg_$cfg.err || { tail -3 gpurun_out/r02_cfg_$cfg.err; continue; }
    python - $cfg <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/r02_cfg_{sys.argv[1]}.json"))
print("cfg", sys.argv[1], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "attn us/step", d["kernel_us_per_step"]["lg.attn"], "frac", round(d["roofline"]["frac"], 4), "clk", d["clocks"]["sm_mhz"])
PY
  done
done
