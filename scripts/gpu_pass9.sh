#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_resnet.py -m gpu -q -rA --no-header -p no:cacheprovider > gpurun_out/t_resnet.log 2>&1; echo "pytest rc=$?"; grep -E "^\[|passed|failed|^E  |Error" gpurun_out/t_resnet.log | head -50
