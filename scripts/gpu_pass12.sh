#!/bin/bash
# validation pass: smoke(), default bench, reference arm (short)
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -6 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err | cut -c1-200
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")})
print("e2e", d["e2e"]); r = d["roofline"]; print("roofline", {k: r.get(k) for k in ("kernel", "achieved", "frac", "executed_frac", "share_of_step", "traffic")})
print("extra", d["extra"]); print("cpu", d["cpu_baseline"])
for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"])[:8]: print("  %-46s" % k, {a: round(b, 3) for a, b in v.items()})
PY
