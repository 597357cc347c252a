#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_models.py -m gpu -q --no-header -p no:cacheprovider -k "linear or cuda_core or bf16x1" > gpurun_out/t_fix.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/t_fix.log
timeout 600 python scripts/step_breakdown.py --mode bf16 --out gpurun_out/breakdown_bf16.json > gpurun_out/breakdown_bf16.log 2>&1; echo "rc=$?"; cat gpurun_out/breakdown_bf16.log | tail -120
timeout 600 python scripts/step_breakdown.py --mode bf16x1 --no-cupti > gpurun_out/breakdown_bf16x1.log 2>&1; echo "rc=$?"; head -50 gpurun_out/breakdown_bf16x1.log
timeout 600 python scripts/conv_microbench.py --mode bf16 --json gpurun_out/microbench_bf16.json > gpurun_out/microbench_bf16.log 2>&1; cat gpurun_out/microbench_bf16.log
