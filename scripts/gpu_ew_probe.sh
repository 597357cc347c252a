#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --quick --no-graph --steps 1 --warmup 3 > gpurun_out/quick.json 2> gpurun_out/quick.err; echo "quick rc=$?"; cat gpurun_out/quick.json
timeout 1200 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --clock-control none -k regex:"norm_|split_positions|conv_fold|avgpool|add_act|gate_" -s 2500 -c 400 --csv --log-file gpurun_out/ew_sol.csv python bench.py --quick --no-graph --steps 1 --warmup 3 > gpurun_out/ncu_ew.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_ew.log | cut -c1-200
