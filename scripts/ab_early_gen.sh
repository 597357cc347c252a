#!/bin/bash
# A/B of the trainer's stream options on one box: AFFGW_EARLY_GEN (gen_update's generator forward issued inside dis_update beside the
# discriminator pass), AFFGW_GEN_HEADS (classifier beside discriminator in gen_update), AFFGW_SIDE_TEXT (text encoder beside the image
# encoder) - edit the variable pair of the loop for the option under test
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py -q --no-header -p no:cacheprovider -x -s \
    -k "early_generator or shared_generator or graph or short_final" > gpurun_out/t_early.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t_early.log
for v in "1 0" "1 1"; do
    set -- $v
    AFFGW_EARLY_GEN=$1 AFFGW_SIDE_TEXT=$2 timeout 600 python bench.py --quick --steps 20 --warmup 3 > gpurun_out/ab_early_$1_$2.json 2> gpurun_out/ab_early_$1_$2.err
    echo "early=$1 side_text=$2 rc=$?"
    python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open('gpurun_out/ab_early_%s_%s.json' % (sys.argv[1], sys.argv[2])).read().strip().splitlines()[-1])
    print(d)
except Exception as e:
    print('no line', e)
PY
done
