#!/bin/bash
# BASELINE.md section 4 numbers: configs 1-3 (tools/gpu_configs.py), config 4 (tools/lg_sweep.py), config 5 at N=1 (bench.py)
mkdir -p gpurun_out
echo "== configs 1-3"
timeout 900 python tools/gpu_configs.py > gpurun_out/r02_configs123.jsonl 2> gpurun_out/r02_configs123.err
cat gpurun_out/r02_configs123.jsonl; tail -3 gpurun_out/r02_configs123.err
echo "== config 4"
timeout 600 python tools/lg_sweep.py 8 10 > gpurun_out/r02_config4_lg_sweep.jsonl 2> gpurun_out/r02_config4.err
cat gpurun_out/r02_config4_lg_sweep.jsonl; tail -3 gpurun_out/r02_config4.err
bash tools/gpu_r2_c.sh
