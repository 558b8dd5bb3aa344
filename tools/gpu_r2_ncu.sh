#!/bin/bash
# ncu evidence of one bench step (16 extracts + 8 matches): --set full for every distinct kernel instantiation.
#   sp   : SuperPoint (16 launches)
#   lg0  : lg.prepare + layer 0 of LightGlue (14 launches)
#   tail : final_proj, matchability, the 8 similarity GEMMs, the 5 assignment kernels
#   attn : the attention kernel once more with source correlation (the roofline kernel of bench.py)
# The per-kernel summaries (tools/ncu_table.py) are produced ON THE BOX; raw reports come back only while they fit the
# 64 MiB transfer limit (attn first).
mkdir -p gpurun_out /tmp/ncu
N="--set full --clock-control none --profile-from-start off -f"
timeout 900 ncu $N --launch-count 16 -o /tmp/ncu/r02_step_sp python tools/gpu_one_step.py > /tmp/ncu/sp.log 2>&1
timeout 900 ncu $N --launch-skip 16 --launch-count 14 -o /tmp/ncu/r02_step_lg0 python tools/gpu_one_step.py > /tmp/ncu/lg0.log 2>&1
timeout 900 ncu $N --launch-skip 134 --launch-count 15 -o /tmp/ncu/r02_step_tail python tools/gpu_one_step.py > /tmp/ncu/tail.log 2>&1
timeout 600 ncu $N --import-source on -k regex:attn --launch-count 1 -o /tmp/ncu/r02_attn2_full python tools/gpu_one_step.py > /tmp/ncu/attn.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_one_step.csv \
  python tools/gpu_one_step.py > /tmp/ncu/list.log 2>&1
for f in sp lg0 tail; do python tools/ncu_table.py /tmp/ncu/r02_step_$f.ncu-rep > gpurun_out/r02_step_${f}_full.csv 2>/tmp/ncu/table_$f.err; done
python tools/ncu_table.py /tmp/ncu/r02_attn2_full.ncu-rep > gpurun_out/r02_attn2_full.csv
ls -la /tmp/ncu/*.ncu-rep
budget=$((56 * 1024 * 1024))
for f in r02_attn2_full r02_step_lg0 r02_step_sp r02_step_tail; do
  sz=$(stat -c %s /tmp/ncu/$f.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 0 ] && [ "$sz" -lt "$budget" ]; then cp /tmp/ncu/$f.ncu-rep gpurun_out/; budget=$((budget - sz)); echo "copied $f ($sz bytes)"; else echo "left $f behind ($sz bytes)"; fi
done
for f in sp lg0 tail attn list; do echo "--- $f"; tail -n 2 /tmp/ncu/$f.log; done
wc -l gpurun_out/*.csv
