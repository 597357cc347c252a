#!/bin/bash
# GPU pass: full -m gpu suite (no -x, prints kept), then the default bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -rA --no-header -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"
grep -E "^\[|passed|failed|FAILED|PASSED" gpurun_out/t_all.log | tail -60
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")})
    print("e2e", d["e2e"]); print("roofline", d["roofline"]); print("extra", d["extra"]); print("cpu", d["cpu_baseline"])
    for k, v in d["kernels"].items(): print(" ", k, {a: round(b, 3) for a, b in v.items()})
except Exception as e:
    print("no bench line", e)
PY
