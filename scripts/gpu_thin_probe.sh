#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_blocks.py -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -4
timeout 300 python scripts/conv_microbench.py --only thin_ --batch 128 --reps 5 2>&1 | tail -5
timeout 300 python scripts/conv_microbench.py --only dis_ --batch 128 --reps 5 2>&1 | tail -6
timeout 300 python scripts/conv_microbench.py --only vgg_ --batch 64 --reps 3 2>&1 | tail -8
timeout 600 python bench.py --quick --steps 5 --warmup 3 2>&1 | tail -2
