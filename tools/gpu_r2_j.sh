#!/bin/bash
# fast mode: its test, then the default bench line (which now carries the fast_mode block)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "fast_mode or sm_limit or superpoint" 2>&1 | tail -5
timeout 400 python bench.py > gpurun_out/r02_bench_fast.json 2> gpurun_out/r02_bench_fast.err || tail -5 gpurun_out/r02_bench_fast.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_fast.json"))
print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 4))
print("fast", d["fast_mode"])
PY
