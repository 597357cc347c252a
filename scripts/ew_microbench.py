"""Streaming-kernel micro-benchmark: CUPTI durations (torch.profiler) of the normalisation / operand-split / fold kernels on
the step's activation shapes, with the achieved GB/s against their algorithmic bytes (fp32 elements read + written).
    python scripts/ew_microbench.py
"""
import os, sys, collections
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import affganwriting_b200 as A
from affganwriting_b200 import ops
from torch.profiler import profile, ProfilerActivity

SHAPES = [(64, 64, 64, 216), (64, 128, 64, 216), (64, 256, 32, 108), (64, 512, 16, 54), (64, 512, 8, 27), (128, 16, 64, 216)]
# algorithmic bytes per element (fp32 reads + writes; bf16 planes = 2 x 2 bytes)
BYTES = {"norm_stats_partial": 4, "norm_apply": 8, "norm_bwd_reduce": 8, "norm_bwd_apply": 12, "split_positions": 8,
         "conv_fold": 8}


def main():
    A.set_precision("bf16")
    flush = torch.empty(200 << 20, dtype=torch.uint8, device="cuda")
    for shp in SHAPES:
        n, c, h, w = shp
        x = ops.to_internal(torch.randn(*shp, device="cuda")).requires_grad_()
        wgt = (torch.randn(c, c, 3, 3, device="cuda") * 0.05).requires_grad_()
        def run():
            flush.zero_()
            y = ops.instance_norm(x, act="relu")
            z = ops.conv2d(y, wgt, None, pad=1, pad_mode="reflect")
            z.backward(torch.ones_like(z))
            x.grad = None; wgt.grad = None
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                run()
            torch.cuda.synchronize()
        dur = collections.defaultdict(list)
        for ev in prof.events():
            if ev.device_type.name == "CUDA":
                dur[ev.name].append(ev.device_time if hasattr(ev, "device_time") else ev.cuda_time)
        elems = n * c * h * w
        print(f"--- {shp}: {elems / 1e6:.1f} M elements")
        for name, v in sorted(dur.items(), key=lambda kv: -sum(kv[1])):
            key = next((k for k in BYTES if k in name), None)
            us = sum(v) / len(v)
            if key:
                print(f"   {key:22s} {us:8.1f} us x{len(v) // 3}  {elems * BYTES[key] / us / 1e3:7.0f} GB/s")
            elif "tcgen05" in name or "wgrad" in name:
                print(f"   {name.split('::')[-1][:40]:22s} {us:8.1f} us x{len(v) // 3}")


if __name__ == "__main__":
    main()
