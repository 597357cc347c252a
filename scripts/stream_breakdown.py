"""One eager training step with per-launch CUDA events on every streaming (HBM-bound) libaffgw call, grouped by
(entry point, algorithmic bytes): where the non-convolution time of the step goes.   python scripts/stream_breakdown.py [batch]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import affganwriting_b200 as A
from affganwriting_b200 import _lib, ops, load_data as LD
from affganwriting_b200.trainer import Trainer
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
A.set_precision("f16")
dev = torch.device("cuda", 0)
t = Trainer(num_writers=500, device=dev, wgrad_stream=False)     # one stream: every kernel is timed alone
batch = LD.batch_to_device(bench.synthetic_batch(B, 50, 1234), dev)
for _ in range(2):
    t.train_step_eager(batch)
ops.start_kernel_timing()
t.train_step_eager(batch)
recs, _lib.PROFILE = _lib.PROFILE, None
ops.stop_kernel_timing()
torch.cuda.synchronize()
agg = {}
for name, e0, e1, nb in recs:
    d = agg.setdefault((name, nb), [0, 0.0])
    d[0] += 1
    d[1] += e0.elapsed_time(e1)
tot = sum(v[1] for v in agg.values())
print(f"streaming kernels: {tot:.2f} ms in one eager step (batch {B})")
for (name, nb), (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{ms:8.3f} ms  x{n:3d}  {nb / 1e6:9.1f} MB/launch  {nb * n / ms / 1e6:7.0f} GB/s  {name}")
