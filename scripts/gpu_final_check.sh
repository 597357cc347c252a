#!/bin/bash
# round-end style validation: GPU tests, smoke(), default bench, reference arm
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | grep -v -i warn | tail -5
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step']); r=d['roofline']; print({k:r[k] for k in ('kernel','achieved','frac','executed_frac','share_of_step')}); print(d['extra']); print(d['cpu_baseline']['value'])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
