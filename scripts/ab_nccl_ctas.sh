#!/bin/bash
# 2-GPU A/B: how many SMs NCCL may take away from the persistent convolution kernels (NCCL_MAX_CTAS), then the NCCL
# data-parallel correctness tests and the 2-GPU bench line
mkdir -p gpurun_out
run() {  # name, env...
    name=$1; shift
    env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
        bench.py --gpus 2 --quick --steps 20 --warmup 3 > gpurun_out/ab_nccl_$name.json 2> gpurun_out/ab_nccl_$name.err
    echo "$name rc=$? $(tail -1 gpurun_out/ab_nccl_$name.json | cut -c1-200)"
}
run default AFFGW_DUMMY=1
run ctas4 NCCL_MAX_CTAS=4
run ctas2 NCCL_MAX_CTAS=2
run ctas8 NCCL_MAX_CTAS=8
timeout 900 python -m pytest tests/test_gpu_dp_nccl.py -q --no-header -p no:cacheprovider -x > gpurun_out/t_dp.log 2>&1; echo "dp tests rc=$?"; tail -2 gpurun_out/t_dp.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-rec-extra > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
python scripts/show_bench.py gpurun_out/bench_n2.json 2>/dev/null | head -8
