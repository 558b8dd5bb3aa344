#!/bin/bash
# 2-GPU run of bench.py exactly as the driver launches it (torchrun, NCCL): the config-5 stream with scatter / gather inside the timed region
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
tail -5 gpurun_out/r02_bench_n2.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02_bench_n2.json") if l.startswith("{")][-1])
print("N=2 value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "stream", d["stream"]["value"], "coll bytes/step", d["stream"]["collective_bytes_per_step"],
      "ms/step", round(d["ms_per_step"], 3), "clk", d["clocks"]["sm_mhz"], "counts", d["stream"]["match_counts_last_step"])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -2
