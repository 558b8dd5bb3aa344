#!/bin/bash
# last check of the round: the whole GPU suite with the bounded-wait library, then with the production library, smoke(), bench
mkdir -p gpurun_out
echo "== GPU suite, bounded-wait library"
ROVER_FE_LIB=$PWD/rover_slam_b200/librover_fe_dbg.so timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== GPU suite, production library"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; tail -2 gpurun_out/r02_bench_n1_final.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_n1_final.json"))
print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "stream", round(d["stream"]["value"], 1), "ms/step", round(d["ms_per_step"], 3),
      "frac", round(d["roofline"]["frac"], 4), "clk", d["clocks"]["sm_mhz"], "launches", d["gpu_launches"], "fast", round(d["fast_mode"]["value"], 1))
print(d["kernel_us_per_step"])
PY
