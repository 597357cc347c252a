#!/bin/bash
# bench (default), then ncu: launch list of one step + full capture of the dominant kernels on one VGG layer
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")})
print("e2e", d["e2e"]); print("roofline", d["roofline"]); print("extra", d["extra"]); print("cpu", d["cpu_baseline"])
for k, v in d["kernels"].items(): print(" ", k, {a: round(b, 3) for a, b in v.items()})
PY
timeout 600 python bench.py --quick --steps 1 --warmup 3 > gpurun_out/quick.json 2> gpurun_out/quick.err; rc=$?; echo "quick rc=$rc"; cat gpurun_out/quick.json
if [ $rc -eq 0 ]; then
  L=$(python -c "import json;print(json.loads(open('gpurun_out/quick.json').read().strip().splitlines()[-1])['gpu_launches'])")
  echo "launches per step: $L"
  timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s $((3*L+200)) -c $((L+400)) --csv --log-file gpurun_out/launches_r01.csv python bench.py --quick --steps 1 --warmup 3 > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"; tail -2 gpurun_out/ncu_list.log
fi
timeout 300 python scripts/conv_microbench.py --only vgg_256_256 --reps 1 > gpurun_out/plain_micro.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:shift -c 6 -o gpurun_out/prof_shift_r01 python scripts/conv_microbench.py --only vgg_256_256 --reps 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
