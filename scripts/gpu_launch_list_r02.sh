#!/bin/bash
# ncu launch list of one eager training step with the current defaults (the short half of gpu_profile_pass_r02.sh)
mkdir -p gpurun_out
timeout 600 python bench.py --quick --no-graph --steps 1 --warmup 3 > gpurun_out/quick_r02.json 2> gpurun_out/quick_r02.err; rc=$?; echo "quick rc=$rc"; cat gpurun_out/quick_r02.json
if [ $rc -eq 0 ]; then
  L=$(python -c "import json;print(json.loads(open('gpurun_out/quick_r02.json').read().strip().splitlines()[-1])['gpu_launches'])"); echo "launches per step: $L"
  timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s $((3*L+200)) -c $((L+2500)) --csv --log-file gpurun_out/launches_r02.csv python bench.py --quick --no-graph --steps 1 --warmup 3 > gpurun_out/ncu_list_r02.log 2>&1; echo "ncu list rc=$?"; tail -2 gpurun_out/ncu_list_r02.log | cut -c1-200
fi
