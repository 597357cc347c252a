"""Time the complete iteration with the native recogniser (bench.py's `full_iteration_with_recogniser`) alone.
    python scripts/rec_iteration_probe.py [--no-graph]"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import affganwriting_b200 as A
from affganwriting_b200.trainer import Trainer
from affganwriting_b200 import load_data as LD
import bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
A.set_precision("f16")
torch.manual_seed(0)
batch = LD.batch_to_device(bench.synthetic_batch(64, 50, 1234), dev)
t2 = Trainer(num_writers=500, device=dev, cuda_graph="--no-graph" not in sys.argv, rec=True)
for _ in range(Trainer.GRAPH_WARMUP + 2):
    t2.train_step(batch)
torch.cuda.synchronize()
for rep in range(8):
    t0 = time.perf_counter()
    t2.train_step(batch)
    torch.cuda.synchronize()
    print("iteration %d: %.1f ms   allocated %.2f GB reserved %.2f GB" % (rep, (time.perf_counter() - t0) * 1e3,
          torch.cuda.memory_allocated() / 2**30, torch.cuda.memory_reserved() / 2**30), flush=True)
