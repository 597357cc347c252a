#!/bin/bash
# GPU pass: micro-benchmark of every conv shape, bench with the tcgen05 wgrad, ncu launch list + full capture of the top kernels
mkdir -p gpurun_out
timeout 600 python scripts/conv_microbench.py --json gpurun_out/microbench.json > gpurun_out/microbench.log 2>&1; echo "microbench rc=$?"; cat gpurun_out/microbench.log | tail -20
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")})
    print("e2e", d["e2e"]); print("roofline", d["roofline"]); print("extra", d["extra"]); print("cpu", d["cpu_baseline"])
    for k, v in d["kernels"].items(): print(" ", k, {a: round(b, 3) for a, b in v.items()})
except Exception as e:
    print("no bench line", e)
PY
# ncu: full capture of the tcgen05 kernels on one mid-size layer (short command)
timeout 300 python scripts/conv_microbench.py --only vgg_256_256 --reps 1 > gpurun_out/plain_micro.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tcgen05 -c 6 -o gpurun_out/prof_conv_r01 python scripts/conv_microbench.py --only vgg_256_256 --reps 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
