#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/conv_microbench.py --only thin_16_16 --batch 128 --reps 1 > gpurun_out/plain_micro.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"shift" -c 4 -o gpurun_out/prof_thin2_r01 python scripts/conv_microbench.py --only thin_16_16 --batch 128 --reps 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
