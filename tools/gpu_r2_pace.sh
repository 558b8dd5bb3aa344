#!/bin/bash
# conv64_strip_kernel: what the MMA issuer's clock reads change.  librover_fe.so = without them (default since this A/B),
# librover_fe_pace.so = compiled in (make pace).  Timing A/B (two alternations) and one --set full capture of conv1b with each library.
mkdir -p gpurun_out /tmp/ncu
timeout 300 python -m pytest tests -m gpu -x -q -k "fast_mode or superpoint_vs_golden" 2>&1 | tail -3
PL=$PWD/rover_slam_b200/librover_fe_pace.so
for rep in 1 2; do
  for m in pace nopace; do
    if [ $m = pace ]; then export ROVER_FE_LIB=$PL; else unset ROVER_FE_LIB; fi
    timeout 200 python bench.py --steps 10 --warmup 3 --cpu-pairs 0 > /tmp/pace_$m.json 2> /tmp/pace_$m.err || { tail -3 /tmp/pace_$m.err; continue; }
    python - $m <<'PY'
import json, sys
d = json.load(open(f"/tmp/pace_{sys.argv[1]}.json"))
k = d["kernel_us_per_step"]
print(sys.argv[1], "conv1b", k["sp.conv1b"], "conv2", k["sp.conv2"], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "clk", d["clocks"]["sm_mhz"],
      "fast", round(d["fast_mode"]["value"], 1), round(d["fast_mode"]["keypoint_overlap_mean"], 4), round(d["fast_mode"]["match_overlap_mean"], 4))
PY
  done
done
for m in pace nopace; do
  if [ $m = pace ]; then export ROVER_FE_LIB=$PL; else unset ROVER_FE_LIB; fi
  timeout 400 ncu --set full --clock-control none --profile-from-start off -f -k regex:conv64_strip --launch-count 1 \
    -o /tmp/ncu/strip_$m python tools/gpu_one_step.py > /tmp/ncu/$m.log 2>&1
  ncu -i /tmp/ncu/strip_$m.ncu-rep --page raw --csv > /tmp/ncu/strip_$m.csv 2>/dev/null
done
unset ROVER_FE_LIB
python - <<'PY' | tee gpurun_out/r02_strip_pace_ncu.txt
import csv
want = ["gpu__time_duration.sum", "sm__cycles_active.avg", "launch__registers_per_thread", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__pipe_tma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
rows = {}
for m in ("pace", "nopace"):
    r = list(csv.reader(open(f"/tmp/ncu/strip_{m}.csv")))
    rows[m] = dict(zip(r[0], r[2])); units = dict(zip(r[0], r[1]))
print(f"{'metric':75s} {'unit':>10s} {'with clock reads':>18s} {'without':>14s}")
for k in want:
    if k in rows["pace"]:
        print(f"{k:75s} {units[k]:>10s} {rows['pace'][k]:>18s} {rows['nopace'][k]:>14s}")
PY
