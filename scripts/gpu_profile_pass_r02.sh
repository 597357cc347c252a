#!/bin/bash
# round-2 profile pass (1 GPU): bench line, reference arm, ncu launch list of one eager step, ncu --set full of the dominant
# kernels; scripts/summarise_profiles.py r02 turns gpurun_out/ into profiles/r02_*
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r02.json 2> gpurun_out/bench_ref_r02.err; echo "reference arm rc=$?"
timeout 600 python bench.py --quick --no-graph --steps 1 --warmup 3 > gpurun_out/quick_r02.json 2> gpurun_out/quick_r02.err; rc=$?; echo "quick rc=$rc"; cat gpurun_out/quick_r02.json
if [ $rc -eq 0 ]; then
  L=$(python -c "import json;print(json.loads(open('gpurun_out/quick_r02.json').read().strip().splitlines()[-1])['gpu_launches'])"); echo "launches per step: $L"
  timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s $((3*L+200)) -c $((L+2500)) --csv --log-file gpurun_out/launches_r02.csv python bench.py --quick --no-graph --steps 1 --warmup 3 > gpurun_out/ncu_list_r02.log 2>&1; echo "ncu list rc=$?"; tail -2 gpurun_out/ncu_list_r02.log | cut -c1-200
fi
timeout 300 python scripts/conv_microbench.py --only vgg_256_256 --reps 1 --mode f16 > gpurun_out/plain_micro_r02.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:shift -c 6 -o gpurun_out/prof_shift_r02 python scripts/conv_microbench.py --only vgg_256_256 --reps 1 --mode f16 > gpurun_out/ncu_full_r02.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full_r02.log
timeout 300 python scripts/step_breakdown.py --mode f16 > gpurun_out/step_breakdown_r02.txt 2>&1; echo "breakdown rc=$?"
timeout 300 python scripts/stream_breakdown.py > gpurun_out/stream_breakdown_r02.txt 2>&1; echo "stream breakdown rc=$?"
timeout 300 python scripts/conv_microbench.py --reps 3 --mode f16 > gpurun_out/micro_r02.txt 2>&1; echo "microbench rc=$?"
