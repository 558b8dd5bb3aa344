#!/bin/bash
# ncu evidence of one bench step (16 extracts + 8 matches): --set full for every distinct kernel instantiation.
#   head : SuperPoint (16 launches) + lg.prepare + layer 0 of LightGlue (13 launches)
#   tail : final_proj, matchability, the 8 similarity GEMMs, the 5 assignment kernels
#   attn : the attention kernel once more with source correlation (the roofline kernel of bench.py)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --profile-from-start off --launch-count 30 -f -o gpurun_out/r02_step_head \
  python tools/gpu_one_step.py > gpurun_out/r02_ncu_head.log 2>&1
timeout 900 ncu --set full --clock-control none --profile-from-start off --launch-skip 134 --launch-count 15 -f -o gpurun_out/r02_step_tail \
  python tools/gpu_one_step.py > gpurun_out/r02_ncu_tail.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn --launch-count 1 -f -o gpurun_out/r02_attn2_full \
  python tools/gpu_one_step.py > gpurun_out/r02_ncu_attn.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_one_step.csv \
  python tools/gpu_one_step.py > gpurun_out/r02_ncu_list.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/r02_ncu_head.log gpurun_out/r02_ncu_tail.log gpurun_out/r02_ncu_attn.log
