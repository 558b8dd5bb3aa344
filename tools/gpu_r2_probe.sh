#!/bin/bash
# Round-2 probe: (1) is ONNXRuntime on the GPU box image?  (2) compute-sanitizer over smoke()  (3) baseline tests + bench
mkdir -p gpurun_out
{
  echo "== python import probes"
  python - <<'PY'
import importlib
for m in ("onnxruntime", "onnx", "cv2", "tensorrt"):
    try:
        mod = importlib.import_module(m)
        print(m, "PRESENT", getattr(mod, "__version__", "?"))
    except Exception as e:
        print(m, "absent:", type(e).__name__, e)
PY
  echo "== find libonnxruntime / onnxruntime anywhere"
  find / -xdev \( -iname '*onnxruntime*' -o -iname 'libonnx*' \) -not -path '/proc/*' 2>/dev/null | head -20
  echo "== pip list | grep -i onnx"
  python -m pip list 2>/dev/null | grep -i -E "onnx|opencv" || echo "(none)"
  echo "== nproc / cpu"
  nproc; grep -m1 'model name' /proc/cpuinfo
} > gpurun_out/r02_ort_probe.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_baseline.txt 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 0 python __graft_entry__.py smoke > gpurun_out/r02_sanitizer_memcheck.txt 2>&1
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 0 python __graft_entry__.py smoke > gpurun_out/r02_sanitizer_racecheck.txt 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_baseline.json 2> gpurun_out/r02_bench_baseline.err
tail -3 gpurun_out/r02_gputest_baseline.txt; tail -5 gpurun_out/r02_sanitizer_memcheck.txt; tail -5 gpurun_out/r02_sanitizer_racecheck.txt; cat gpurun_out/r02_ort_probe.txt | head -30
