#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_blocks.py -m gpu -q --no-header -p no:cacheprovider -x -k bucket 2>&1 | tail -2
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 scripts/dp_profile.py 2>&1 | grep -v -i warn | tail -3 | cut -c1-400
bash scripts/gpu_multi_bench.sh 2 2>&1 | tail -5 | cut -c1-300
