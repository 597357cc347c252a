#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -q --no-header -p no:cacheprovider -x -k "cuda_graph_replay" -s 2>&1 | grep -v -i warn | tail -5
N=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 8 --warmup 3 --quick 2>&1 | grep quick
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 8 --warmup 3 --quick --no-overlap 2>&1 | grep quick
