#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/t_all.log
timeout 600 python bench.py --quick --steps 5 --warmup 3 2>&1 | tail -2
