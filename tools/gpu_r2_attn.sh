#!/bin/bash
# First GPU run of the persistent attention kernel: bounded-wait build first, then parity, then A/B timing.
mkdir -p gpurun_out
echo "== debug-wait build, LightGlue parity subset"
ROVER_FE_LIB=$PWD/rover_slam_b200/librover_fe_dbg.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu \
  -k "lightglue_vs_golden_synth or lightglue_vs_oracle_ragged or batched_matching_equals_single or empty_inputs or cv2dnn_golden" 2>&1 | tail -15
echo "== full GPU tests"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== bench A/B"
RFE_ATTN=3 timeout 300 python bench.py --steps 20 --warmup 5 --cpu-pairs 0 > gpurun_out/r02_bench_attn3.json 2> gpurun_out/r02_bench_attn3.err
RFE_ATTN=2 timeout 300 python bench.py --steps 20 --warmup 5 --cpu-pairs 0 > gpurun_out/r02_bench_attn2.json 2> gpurun_out/r02_bench_attn2.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_attn3.json", "gpurun_out/r02_bench_attn2.json"):
    try:
        d = json.load(open(f))
        print(f, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 3), "attn us/step", d["kernel_us_per_step"].get("lg.attn"),
              "frac", round(d["roofline"]["frac"], 4), "clk", d["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "FAILED", e)
PY
echo "== attn2 role counters"
timeout 300 python tools/gpu_attn2_prof.py > gpurun_out/r02_attn3_prof.txt 2>&1; cat gpurun_out/r02_attn3_prof.txt | tail -25
