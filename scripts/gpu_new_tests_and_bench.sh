#!/bin/bash
# new / changed tests first (fast feedback), then the whole GPU suite, then the default bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_wire.py tests/test_gpu_resnet.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/t_new.log 2>&1; echo "new rc=$?"; tail -15 gpurun_out/t_new.log
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/t_all.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks')})
print(d['roofline'])
PY
