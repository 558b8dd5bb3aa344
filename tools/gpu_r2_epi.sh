#!/bin/bash
# LINEAR-epilogue phase counters (make epiprof) for the K = 256 / 512 linears, then correctness + the bench breakdown
mkdir -p gpurun_out
for tag in lg.to_qk lg.ffn0 lg.ffn3; do
  ROVER_FE_LIB=$PWD/rover_slam_b200/librover_fe_epiprof.so RFE_PROF_TAG=$tag timeout 120 python tools/gpu_umma_prof.py 2>&1 | tail -10
done | tee gpurun_out/r02_linear_epilogue_phases_after.txt
timeout 400 python -m pytest tests -m gpu -x -q -k "lightglue or bench_shape or one_to_many or config3 or match" 2>&1 | tail -3
for rep in 1 2; do
timeout 200 python bench.py --steps 20 --warmup 5 --cpu-pairs 0 > gpurun_out/r02_bench_epi.json 2> gpurun_out/r02_bench_epi.err || tail -3 gpurun_out/r02_bench_epi.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_epi.json"))
k = d["kernel_us_per_step"]
print("value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 4), "clk", d["clocks"]["sm_mhz"],
      {t: k[t] for t in ["lg.wqkv", "lg.out_proj", "lg.ffn0", "lg.ffn3", "lg.to_qk", "lg.to_v", "lg.to_out", "sp.convDb", "sp.convPb", "lg.sim"]})
PY
done
