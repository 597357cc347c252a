#!/bin/bash
# GPU pass: CUDA-core parity, the tcgen05 cross-check under its own timeout, whole-model parity; compact report
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_blocks.py -m gpu -q 2>&1 > gpurun_out/t_blocks.log
timeout 300 python -m pytest tests/test_gpu_conv_tc.py -m gpu -q 2>&1 > gpurun_out/t_tc.log
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q -s 2>&1 > gpurun_out/t_models.log
for f in t_blocks t_tc t_models; do echo "== $f"; grep -E "^\[|^      cos|^E   +(Assert|assert|Runtime)|passed|failed|^FAILED|^ERROR" gpurun_out/$f.log | cut -c1-220 | head -60; done
echo "== bench"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")})
    print("e2e", d["e2e"]); print("roofline", d["roofline"]); print("step", d["step_tflops"]); print("extra", d["extra"]); print("cpu", d["cpu_baseline"])
    for k, v in d["kernels"].items(): print(" ", k, {a: round(b, 3) for a, b in v.items()})
except Exception as e:
    print("no bench line", e)
PY
