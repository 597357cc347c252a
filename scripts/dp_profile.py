"""Where does a data-parallel step spend its extra time?  torchrun --nproc-per-node 2 scripts/dp_profile.py
Profiles one CUDA-graph training step per rank with torch.profiler (CUPTI) and prints, on rank 0: busy time, idle gaps, NCCL
kernel time, pack / unpack time."""
import os, sys, collections
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import affganwriting_b200 as A
from affganwriting_b200 import load_data as LD
from affganwriting_b200.trainer import Trainer
import bench
from torch.profiler import profile, ProfilerActivity

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
A.set_precision("bf16")
torch.manual_seed(0)
tr = Trainer(num_writers=500, device=dev, cuda_graph=True)
batch = LD.batch_to_device(bench.synthetic_batch(64, 50, 1234 + rank), dev)
for _ in range(8):
    tr.train_step(batch)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        tr.train_step(batch)
    torch.cuda.synchronize()
if rank == 0:
    evs = sorted([e for e in prof.events() if e.device_type.name == "CUDA"], key=lambda e: e.time_range.start)
    t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
    busy, cur_end, gaps = 0.0, t0, []
    for e in evs:
        s, en = e.time_range.start, e.time_range.end
        if s > cur_end:
            gaps.append((s - cur_end, e.name[:60]))
            busy += en - s
            cur_end = en
        elif en > cur_end:
            busy += en - cur_end
            cur_end = en
    print(f"world {world}: span {(t1 - t0) / 2e3:.2f} ms/step, busy {busy / 2e3:.2f} ms/step, idle {(t1 - t0 - busy) / 2e3:.2f} ms/step")
    by = collections.defaultdict(lambda: [0.0, 0])
    for e in evs:
        k = "nccl" if "nccl" in e.name.lower() else ("bucket" if "bucket" in e.name else ("adam" if "Optim" in e.name or "multi_tensor" in e.name else None))
        if k:
            by[k][0] += (e.time_range.end - e.time_range.start) / 2e3
            by[k][1] += 1
    print({k: (round(v[0], 3), v[1] // 2) for k, v in by.items()})
    print("largest gaps (us, next kernel):", [(round(g, 1), n) for g, n in sorted(gaps, reverse=True)[:8]])
if world > 1:
    dist.destroy_process_group()
