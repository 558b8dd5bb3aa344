#!/bin/bash
# A/B of the FFN GEMM tile width (128 x 128 double-buffered vs 128 x 256 single-buffered accumulators)
mkdir -p gpurun_out
echo "== parity with 128x256 FFN tiles"
RFE_FFN0_BN=256 RFE_FFN3_BN=256 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "lightglue or end_to_end or bench_shape" 2>&1 | tail -4
for cfg in "128 128" "256 128" "128 256" "256 256"; do
  set -- $cfg
  RFE_FFN0_BN=$1 RFE_FFN3_BN=$2 timeout 300 python bench.py --steps 10 --warmup 3 --cpu-pairs 0 > gpurun_out/r02_ffn_$1_$2.json 2> gpurun_out/r02_ffn_$1_$2.err
  python - "$1" "$2" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/r02_ffn_{sys.argv[1]}_{sys.argv[2]}.json"))
k = d["kernel_us_per_step"]
print("ffn0 bn", sys.argv[1], "ffn3 bn", sys.argv[2], "| ffn0", k["lg.ffn0"], "ffn3", k["lg.ffn3"], "| ms/step", round(d["ms_per_step"], 3), "clk", d["clocks"]["sm_mhz"])
PY
done
echo "== fused conv1a (setmaxnreg, four producer warps)"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_conv1a" 2>&1 | tail -3
RFE_FUSE_CONV1A=1 timeout 300 python bench.py --steps 10 --warmup 3 --cpu-pairs 0 > gpurun_out/r02_fuse1a.json 2> gpurun_out/r02_fuse1a.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_fuse1a.json"))
k = d["kernel_us_per_step"]
print("fused: conv1a", k.get("sp.conv1a"), "conv1b", k["sp.conv1b"], "| ms/step", round(d["ms_per_step"], 3), "clk", d["clocks"]["sm_mhz"])
PY
