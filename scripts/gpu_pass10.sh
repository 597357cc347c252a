#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -q -rA --no-header -p no:cacheprovider -k "cuda_graph" > gpurun_out/t_graph.log 2>&1; echo "pytest rc=$?"; grep -E "^\[|largest|passed|failed|^E  |Error" gpurun_out/t_graph.log | head -30
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -5 gpurun_out/bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")})
print("e2e", d["e2e"]); print("roofline", {k: d["roofline"][k] for k in ("kernel","achieved","frac")}); print("extra", d["extra"])
PY
