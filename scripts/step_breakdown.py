"""Where one training step spends its time: per-(kernel kind, conv shape) CUDA-event timing and a CUPTI kernel-name table.

    python scripts/step_breakdown.py [--mode bf16] [--batch 64] [--out gpurun_out/breakdown.json]
"""
import argparse, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import affganwriting_b200 as A
from affganwriting_b200 import ops
from affganwriting_b200.trainer import Trainer
import bench
from affganwriting_b200 import load_data as LD


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="bf16")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-cupti", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    A.set_precision(args.mode)
    torch.manual_seed(0)
    tr = Trainer(num_writers=500, device=dev, wgrad_stream=False)     # one stream: every kernel is timed alone (the step number is the eager single-stream step)
    batch = LD.batch_to_device(bench.synthetic_batch(args.batch, 50, 1234), dev)
    for _ in range(3):
        tr.train_step(batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        tr.train_step(batch)
    e1.record(); torch.cuda.synchronize()
    step_ms = e0.elapsed_time(e1) / 3
    print(f"step {step_ms:.1f} ms (mode {args.mode}, batch {args.batch})")
    import time
    torch.cuda.synchronize(); t0 = time.perf_counter(); tr.train_step(batch); t1 = time.perf_counter(); torch.cuda.synchronize()
    print(f"host-side issue time of one step (no sync): {1e3 * (t1 - t0):.1f} ms")
    ops.start_kernel_timing()
    tr.train_step(batch)
    rec = ops.stop_kernel_timing(by_shape=True)
    rows = sorted(((v["ms"], k, v) for k, v in rec.items()), reverse=True)
    tot = sum(r[0] for r in rows)
    print(f"convolution launches: {tot:.1f} ms")
    out = {"step_ms": step_ms, "conv_ms": tot, "rows": []}
    for ms, (name, tag), v in rows:
        tf = v["flops"] / (ms * 1e-3) / 1e12 if ms > 0 else 0
        out["rows"].append({"kind": name, "shape": tag, "launches": v["launches"], "ms": ms, "tflops": tf})
        if ms > 0.004 * tot:
            print(f"  {ms:8.3f} ms  x{v['launches']:3d}  {tf:7.1f} TF/s  {name:20s} {tag}")
    if not args.no_cupti:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            tr.train_step(batch)
            torch.cuda.synchronize()
        ev = {}
        for e in prof.events():
            if e.device_type == torch.autograd.DeviceType.CUDA:
                d = ev.setdefault(e.name, [0, 0.0])
                d[0] += 1; d[1] += e.device_time / 1e3 if hasattr(e, "device_time") else e.cuda_time / 1e3
        krows = sorted(((v[1], k, v[0]) for k, v in ev.items()), reverse=True)
        ktot = sum(r[0] for r in krows)
        print(f"CUPTI: {ktot:.1f} ms of kernels / memcpys in one step")
        out["kernels"] = []
        for ms, k, n in krows[:75]:
            print(f"  {ms:8.3f} ms  x{n:4d}  {k[:110]}")
            out["kernels"].append({"name": k, "launches": n, "ms": ms})
    if args.out:
        json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
