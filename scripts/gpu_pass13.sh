#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_blocks.py -m gpu -q --no-header -p no:cacheprovider -k "graph or bucket" 2>&1 | tail -5
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err | cut -c1-200
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}); print("e2e", d["e2e"]["value"]); print("extra", d["extra"]); print(d["roofline"]["share_of_step"])
PY
timeout 900 python bench.py --no-cpu-baseline --encoder resnet18 > gpurun_out/bench_r18.json 2> gpurun_out/bench_r18.err; echo "bench r18 rc=$?"; tail -2 gpurun_out/bench_r18.err | cut -c1-200
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r18.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}); print("e2e", d["e2e"]["value"]); print("extra", d["extra"]); print(d["config"]["workload"][:80])
PY
