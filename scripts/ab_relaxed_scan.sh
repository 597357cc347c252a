#!/bin/bash
# how deep can ops.relaxed_forward go? worst discriminator tensor of the complete dis_update (50 planes, batch 8) and step time
mkdir -p gpurun_out
for f in 26 19 14 7 0; do
    AFFGW_RELAXED_VGG_FROM=$f timeout 300 python -m pytest tests/test_gpu_parity_c50.py -q --no-header -p no:cacheprovider -x -s -k "dis_update_with_its_own and f16" > gpurun_out/t_scan_$f.log 2>&1
    echo "from=$f rc=$? $(grep -h 'complete dis_update' gpurun_out/t_scan_$f.log | cut -c1-260)"
    AFFGW_RELAXED_VGG_FROM=$f timeout 600 python bench.py --quick --steps 20 --warmup 3 > gpurun_out/ab_scan_$f.json 2> gpurun_out/ab_scan_$f.err
    echo "   bench from=$f rc=$? $(tail -1 gpurun_out/ab_scan_$f.json | cut -c1-120)"
done
