#!/bin/bash
# refresh of the per-kernel ncu evidence for the FINAL build: SuperPoint (16 launches), LightGlue layer 0 (14), the attention
# kernel with source correlation, the launch list.  Summaries are produced on the box (tools/ncu_table.py).
mkdir -p gpurun_out /tmp/ncu
N="--set full --clock-control none --profile-from-start off -f"
timeout 600 ncu $N --launch-count 16 -o /tmp/ncu/r02_step_sp python tools/gpu_one_step.py > /tmp/ncu/sp.log 2>&1
timeout 600 ncu $N --launch-skip 16 --launch-count 14 -o /tmp/ncu/r02_step_lg0 python tools/gpu_one_step.py > /tmp/ncu/lg0.log 2>&1
timeout 400 ncu $N --import-source on -k regex:attn2 --launch-count 1 -o /tmp/ncu/r02_attn2_full python tools/gpu_one_step.py > /tmp/ncu/attn.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_one_step.csv \
  python tools/gpu_one_step.py > /tmp/ncu/list.log 2>&1
for f in sp lg0; do python tools/ncu_table.py /tmp/ncu/r02_step_$f.ncu-rep > gpurun_out/r02_step_${f}_full.csv 2>/tmp/ncu/table_$f.err; done
python tools/ncu_table.py /tmp/ncu/r02_attn2_full.ncu-rep > gpurun_out/r02_attn2_full.csv
cp /tmp/ncu/r02_attn2_full.ncu-rep gpurun_out/
for f in sp lg0 attn list; do echo "--- $f"; tail -n 1 /tmp/ncu/$f.log; done
cat gpurun_out/r02_attn2_full.csv
wc -l gpurun_out/r02_step_sp_full.csv gpurun_out/r02_step_lg0_full.csv gpurun_out/r02_launches_one_step.csv
