#!/bin/bash
mkdir -p gpurun_out
echo "== full GPU tests"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== bench N=1"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err
tail -3 gpurun_out/r02_bench_c.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_c.json"))
print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), d["e2e"]["h2d_bytes_per_step"], d["e2e"]["d2h_bytes_per_step"], "stream", round(d["stream"]["value"], 1),
      "ms/step", round(d["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 4), "clk", d["clocks"]["sm_mhz"])
print("latency", d.get("latency"))
print("cpu", d.get("cpu_baseline"))
print(d["kernel_us_per_step"])
PY
