import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import affganwriting_b200 as A
from affganwriting_b200.trainer import Trainer
import bench
from affganwriting_b200 import load_data as LD
A.set_precision("bf16")
batch = LD.batch_to_device(bench.synthetic_batch(4, 50, 7), torch.device('cuda'))
dev = torch.device("cuda", 0)
torch.manual_seed(0); a = Trainer(device=dev)
b = Trainer(device=dev); b.model.load_state_dict(a.model.state_dict())
g = Trainer(device=dev, cuda_graph=True); g.model.load_state_dict(a.model.state_dict())
for it in range(7):
    la, lb, lg = a.train_step(batch), b.train_step(batch), g.train_step(batch)
    print(it, " ".join(f"{k}: {float(la[k]):.5f}/{float(lb[k]):.5f}/{float(lg[k]):.5f}" for k in ("cla", "dis", "gen")))
def cmp(x, y, tag):
    rows = []
    sx, sy = x.model.state_dict(), y.model.state_dict()
    for k, v in sx.items():
        if v.is_floating_point():
            rows.append((float((v - sy[k]).abs().max()), k))
    rows.sort(reverse=True)
    print(tag, rows[:6])
import os
print("env", {k: v for k, v in os.environ.items() if k.startswith("AFFGW")})
cmp(a, b, "eager vs eager")
cmp(a, g, "eager vs graph")
