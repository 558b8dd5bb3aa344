#!/bin/bash
# final validation of round 2: full GPU suite, smoke(), default bench line, refreshed ncu of the LightGlue tail + launch list
mkdir -p gpurun_out /tmp/ncu
echo "== full GPU tests"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench (default flags)"
timeout 600 python bench.py > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err
tail -2 gpurun_out/r02_bench_n1_final.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_n1_final.json"))
print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "stream", round(d["stream"]["value"], 1), "ms/step", round(d["ms_per_step"], 3),
      "frac", round(d["roofline"]["frac"], 4), "clk", d["clocks"], "launches", d["gpu_launches"])
print("latency", d.get("latency"))
print("cpu", d.get("cpu_baseline"))
print(d["kernel_us_per_step"])
PY
echo "== ncu tail"
N="--set full --clock-control none --profile-from-start off -f"
timeout 600 ncu $N --launch-skip 152 --launch-count 15 -o /tmp/ncu/r02_step_tail python tools/gpu_one_step.py > /tmp/ncu/tail.log 2>&1
python tools/ncu_table.py /tmp/ncu/r02_step_tail.ncu-rep > gpurun_out/r02_step_tail_full.csv 2>/tmp/ncu/table_tail.err
timeout 300 ncu $N -k regex:attn2_combine --launch-count 1 -o /tmp/ncu/r02_combine python tools/gpu_one_step.py > /tmp/ncu/comb.log 2>&1
python tools/ncu_table.py /tmp/ncu/r02_combine.ncu-rep > gpurun_out/r02_attn2_combine_full.csv 2>/tmp/ncu/table_comb.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_one_step.csv \
  python tools/gpu_one_step.py > /tmp/ncu/list.log 2>&1
tail -n 2 /tmp/ncu/tail.log /tmp/ncu/comb.log /tmp/ncu/list.log 2>/dev/null | cat
wc -l gpurun_out/*.csv
