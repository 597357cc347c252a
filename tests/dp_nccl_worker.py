"""Worker of tests/test_gpu_dp_nccl.py: one rank of a 2-GPU data-parallel run over NCCL (launched with torchrun).

Checks, per sub-network (cla / dis / gen) of one training iteration with the ResNet-18 style encoder (BASELINE.json
configs[2]):
  * the gradients GradientReducer leaves in .grad equal the MEAN over ranks of the gradients each rank computed on its own
    shard (gathered with an independent all_gather before the reduction);
  * tensors without a gradient stay without one;
  * after the Adam step the weights are bit-identical on both ranks.
Rank 0 writes a JSON report to argv[1].
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    out_path = sys.argv[1]
    encoder = sys.argv[2] if len(sys.argv) > 2 else "resnet18"
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import affganwriting_b200 as A
    from affganwriting_b200 import load_data as LD
    from affganwriting_b200.trainer import Trainer
    import bench

    A.set_precision("f16")
    torch.manual_seed(100 + rank)                     # different initial weights per rank: broadcast_module must fix that
    tr = Trainer(num_writers=500, device=dev, encoder=None if encoder == "vgg" else encoder, bucket_bytes=8 << 20)
    batch = LD.batch_to_device(bench.synthetic_batch(4, 50, seed=7 + rank), dev)      # a different shard per rank
    report = {"world": world, "encoder": encoder, "subnets": {}}
    ok = True
    # identical weights after the constructor's broadcast
    for name, p in tr.model.named_parameters():
        ref = p.detach().clone()
        dist.broadcast(ref, src=0)
        if not torch.equal(ref, p.detach()):
            ok = False
            report.setdefault("not_broadcast", []).append(name)
    for sub in ("cla", "dis", "gen"):
        tr._fwd_bwd(sub, batch, 0)
        params = [p for grp in tr.opt[sub].param_groups for p in grp["params"]]
        live = [p for p in params if p.grad is not None]
        own = torch.cat([p.grad.reshape(-1) for p in live]).clone()
        gathered = [torch.empty_like(own) for _ in range(world)]
        dist.all_gather(gathered, own)
        expect = torch.stack(gathered).double().mean(0)
        n_buckets = tr.red[sub].reduce()
        got = torch.cat([p.grad.reshape(-1) for p in live]).double()
        err = float((got - expect).abs().max() / expect.abs().max())
        differ = float((gathered[0].double() - gathered[1].double()).abs().max() / expect.abs().max())
        live_ids = {id(p) for p in live}
        none_kept = all(p.grad is None for p in params if id(p) not in live_ids)
        tr.opt[sub].step()
        from affganwriting_b200 import ops
        ops.weights_updated(params)
        same = True
        for p in live:
            ref = p.detach().clone()
            dist.broadcast(ref, src=0)
            same = same and bool(torch.equal(ref, p.detach()))
        report["subnets"][sub] = {"tensors_with_grad": len(live), "tensors_without_grad": len(params) - len(live),
                                  "buckets": n_buckets, "max_rel_err_vs_mean_of_rank_gradients": err,
                                  "rank_gradients_differ_by": differ, "weights_identical_after_step": same,
                                  "none_grads_kept": none_kept}
        ok = ok and err <= 1e-6 and differ > 1e-3 and same and none_kept and n_buckets >= 1
    report["ok"] = bool(ok)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    report["ok_all_ranks"] = bool(int(flag.item()))
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(report, f, indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
