#!/bin/bash
# 2-GPU A/B: SMs left to NCCL during the config-5 stream (RFE_SM_RESERVE) and NCCL's CTA cap
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29537 \
    bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_n2_$name.json 2> gpurun_out/r02_n2_$name.err || tail -3 gpurun_out/r02_n2_$name.err
  python - $name <<'PY'
import json, sys
d = json.loads([l for l in open(f"gpurun_out/r02_n2_{sys.argv[1]}.json") if l.startswith("{")][-1])
print(sys.argv[1], "value", round(d["value"], 1), "stream", round(d["stream"]["value"], 1), "ratio", round(d["stream"]["value"] / d["value"], 3), "clk", d["clocks"]["sm_mhz"])
PY
}
run base RFE_SM_RESERVE=0 NCCL_MAX_CTAS=32 NCCL_MAX_P2P_NCHANNELS=32
run ctas4 RFE_SM_RESERVE=0
run res4 RFE_SM_RESERVE=4
run res8 RFE_SM_RESERVE=8 NCCL_MAX_CTAS=8 NCCL_MAX_P2P_NCHANNELS=8
