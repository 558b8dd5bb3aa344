#!/bin/bash
mkdir -p gpurun_out
echo "== full GPU tests"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== attn2 role counters"
RFE_ATTN=2 timeout 300 python tools/gpu_attn2_prof.py > gpurun_out/r02_attn2_phases.txt 2>&1; cat gpurun_out/r02_attn2_phases.txt | tail -28
echo "== bench"
timeout 300 python bench.py --steps 20 --warmup 5 --cpu-pairs 0 > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_b.json"))
print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 4), "clk", d["clocks"]["sm_mhz"])
print(d["kernel_us_per_step"])
PY
