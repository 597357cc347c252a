#!/bin/bash
# final-style pass: bench (graph), 1 GPU; ncu launch list + full capture of the dominant kernels
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err | cut -c1-300
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")})
print("e2e", d["e2e"]); r = d["roofline"]; print("roofline", {k: r[k] for k in ("kernel", "achieved", "frac", "executed_frac", "share_of_step", "avg_launch_ms")})
print("extra", d["extra"]); print("cpu", d["cpu_baseline"])
for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"])[:14]: print("  %-46s" % k, {a: round(b, 3) for a, b in v.items()})
PY
timeout 600 python bench.py --quick --no-graph --steps 1 --warmup 3 > gpurun_out/quick.json 2> gpurun_out/quick.err; rc=$?; echo "quick rc=$rc"; cat gpurun_out/quick.json
if [ $rc -eq 0 ]; then
  L=$(python -c "import json;print(json.loads(open('gpurun_out/quick.json').read().strip().splitlines()[-1])['gpu_launches'])"); echo "launches per step: $L"
  timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s $((3*L+200)) -c $((L+2500)) --csv --log-file gpurun_out/launches_r01b.csv python bench.py --quick --no-graph --steps 1 --warmup 3 > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"; tail -2 gpurun_out/ncu_list.log | cut -c1-200
fi
timeout 300 python scripts/conv_microbench.py --only vgg_256_256 --reps 1 > gpurun_out/plain_micro.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:shift -c 6 -o gpurun_out/prof_shift_r01b python scripts/conv_microbench.py --only vgg_256_256 --reps 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
