#!/bin/bash
# round-2, session 3: microbench of every step shape + ncu --set full of the thin (16-channel, full resolution) layer's kernels
# and of the tiny-map im2col GEMMs
mkdir -p gpurun_out
timeout 600 python scripts/conv_microbench.py --reps 3 --mode f16 --batch 128 > gpurun_out/micro_b128.log 2>&1; echo "micro rc=$?"
timeout 300 python scripts/conv_microbench.py --only thin_16_16 --reps 1 --mode f16 --batch 128 > /dev/null 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"shift|split_pos|fold|amax" -c 14 -o gpurun_out/prof_thin16_r02 python scripts/conv_microbench.py --only thin_16_16 --reps 1 --mode f16 --batch 128 > gpurun_out/ncu_thin.log 2>&1; echo "ncu thin rc=$?"; tail -2 gpurun_out/ncu_thin.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"igemm|wgrad_tcgen05" -c 9 -o gpurun_out/prof_dis512_r02 python scripts/conv_microbench.py --only dis_512_512 --reps 1 --mode f16 --batch 128 > gpurun_out/ncu_dis512.log 2>&1; echo "ncu dis rc=$?"; tail -2 gpurun_out/ncu_dis512.log
