#!/bin/bash
# N-GPU run of bench.py exactly as the driver launches it (torchrun, NCCL); N = first argument (default 8)
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29527 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
tail -3 gpurun_out/r02_bench_n$N.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
d = json.loads([l for l in open(f"gpurun_out/r02_bench_n{n}.json") if l.startswith("{")][-1])
print("N=" + n, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "stream", d["stream"]["value"], "coll bytes/step", d["stream"]["collective_bytes_per_step"],
      "ms/step", round(d["ms_per_step"], 3), "clk", d["clocks"]["sm_mhz"])
PY
