#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv_tc.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/t_shift.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/t_shift.log
timeout 300 python scripts/conv_microbench.py --mode bf16 --json gpurun_out/microbench_shift.json > gpurun_out/microbench_shift.log 2>&1; cat gpurun_out/microbench_shift.log
