#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/t_all.log
timeout 600 python scripts/step_breakdown.py --mode bf16 --out gpurun_out/breakdown_bf16.json > gpurun_out/breakdown_bf16.log 2>&1; echo "rc=$?"; cat gpurun_out/breakdown_bf16.log | tail -105
