#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/step_breakdown.py --mode bf16 --out gpurun_out/breakdown_bf16.json > gpurun_out/breakdown_bf16.log 2>&1; echo "rc=$?"; grep -A60 "CUPTI" gpurun_out/breakdown_bf16.log | head -64
