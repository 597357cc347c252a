#!/bin/bash
# A/B of ops.relaxed_forward in dis_update's no_grad generator forward + the parity tests at the benchmarked configuration
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_c50.py -q --no-header -p no:cacheprovider -x -s -k "dis_update or gen_update" > gpurun_out/t_relaxed.log 2>&1; echo "pytest rc=$?"; grep -h "dis_update\|gen_update\|passed\|failed" gpurun_out/t_relaxed.log | cut -c1-330
timeout 600 python -m pytest tests/test_gpu_models.py -q --no-header -p no:cacheprovider -x -k "graph or early or shared" > gpurun_out/t_relaxed2.log 2>&1; echo "pytest2 rc=$?"; tail -2 gpurun_out/t_relaxed2.log
for v in "0 26" "1 26" "1 19" "0 26" "1 26"; do
    set -- $v
    AFFGW_RELAXED_DIS_FWD=$1 AFFGW_RELAXED_VGG_FROM=$2 timeout 600 python bench.py --quick --steps 20 --warmup 3 > gpurun_out/ab_relaxed_$1_$2.json 2> gpurun_out/ab_relaxed_$1_$2.err
    echo "relaxed=$1 from=$2 rc=$? $(tail -1 gpurun_out/ab_relaxed_$1_$2.json | cut -c1-120)"
done
