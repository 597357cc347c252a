#!/bin/bash
# 8-GPU (or N-GPU) data-parallel bench under torchrun, as the driver launches it
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"; grep -v Warning gpurun_out/bench_n$N.err | tail -4 | cut -c1-300
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "n_gpus", "ms_per_step", "gpu_launches", "scaling")}); print("e2e", d["e2e"]); print(d["extra"]); print(d["clocks"])
PY
