"""Per-layer micro-benchmark of the convolution kernels at the step's shapes (batch 64).
CUDA-event timing on the launching stream, warm-up first, median of `reps` launches; prints TFLOP/s (direct-conv
MAC x 2) for forward, dgrad and wgrad.  Used for kernel tuning and as the short command ncu profiles.

    python scripts/conv_microbench.py [--only NAME] [--reps 5] [--batch 64] [--mode bf16]
"""
import argparse, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import affganwriting_b200 as A
from affganwriting_b200 import ops

# name, H, W, Cin, Cout, k, pad, pad_mode, upsample
SHAPES = [
    ("vgg_64_64@64x216", 64, 216, 64, 64, 3, 1, "zero", 1),
    ("vgg_64_128@64x216", 64, 216, 64, 128, 3, 1, "zero", 1),
    ("vgg_128_128@64x216", 64, 216, 128, 128, 3, 1, "zero", 1),
    ("vgg_128_256@32x108", 32, 108, 128, 256, 3, 1, "zero", 1),
    ("vgg_256_256@32x108", 32, 108, 256, 256, 3, 1, "zero", 1),
    ("vgg_256_512@16x54", 16, 54, 256, 512, 3, 1, "zero", 1),
    ("vgg_512_512@16x54", 16, 54, 512, 512, 3, 1, "zero", 1),
    ("vgg_512_512@8x27", 8, 27, 512, 512, 3, 1, "zero", 1),
    ("dec_res_512@8x27", 8, 27, 512, 512, 3, 1, "reflect", 1),
    ("dec_up_512_256@8x27", 8, 27, 512, 256, 5, 2, "reflect", 2),
    ("dec_up_256_128@16x54", 16, 54, 256, 128, 5, 2, "reflect", 2),
    ("dec_up_128_64@32x108", 32, 108, 128, 64, 5, 2, "reflect", 2),
    ("mix_1024_512@8x27", 8, 27, 1024, 512, 1, 0, "zero", 1),
    ("thin_16_16@64x216", 64, 216, 16, 16, 3, 1, "reflect", 1),
    ("thin_16_32@64x216", 64, 216, 16, 32, 3, 1, "reflect", 1),
    ("thin_32_32@32x108", 32, 108, 32, 32, 3, 1, "reflect", 1),
    ("thin_32_64@32x108", 32, 108, 32, 64, 3, 1, "reflect", 1),
    ("dis_64_64@16x54", 16, 54, 64, 64, 3, 1, "reflect", 1),
    ("dis_128_128@8x27", 8, 27, 128, 128, 3, 1, "reflect", 1),
    ("dis_256_256@4x14", 4, 14, 256, 256, 3, 1, "reflect", 1),
    ("dis_512_512@2x7", 2, 7, 512, 512, 3, 1, "reflect", 1),
    ("dis_256_512@4x14", 4, 14, 256, 512, 3, 1, "reflect", 1),
    ("dis_512_1024@2x7", 2, 7, 512, 1024, 3, 1, "reflect", 1),
    ("iaff_512_128@8x27", 8, 27, 512, 128, 1, 0, "zero", 1),
    ("iaff_128_512@8x27", 8, 27, 128, 512, 1, 0, "zero", 1),
]


def median_ms(fn, reps):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--mode", default="bf16")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    A.set_precision(args.mode)
    torch.manual_seed(0)
    flush = torch.empty(160 << 20, dtype=torch.uint8, device="cuda")    # > L2 (126 MB)
    rows = []
    for name, h, w, ci, co, k, p, pm, up in SHAPES:
        if args.only and args.only not in name:
            continue
        x = ops.to_internal(torch.randn(args.batch, ci, h, w, device="cuda")).requires_grad_()
        wgt = (torch.randn(co, ci, k, k, device="cuda") * (2.0 / (ci * k * k)) ** 0.5).requires_grad_()
        b = torch.zeros(co, device="cuda", requires_grad=True)
        y = ops.conv2d(x, wgt, b, pad=p, pad_mode=pm, upsample=up)
        gy = torch.randn_like(y)
        flops = 2.0 * y.shape[0] * y.shape[2] * y.shape[3] * co * ci * k * k
        for _ in range(2):
            x.grad = wgt.grad = b.grad = None
            ops.conv2d(x, wgt, b, pad=p, pad_mode=pm, upsample=up).backward(gy)
        torch.cuda.synchronize()
        res = {}
        for _ in range(args.reps):
            flush.zero_()
            x.grad = wgt.grad = b.grad = None
            ops.start_kernel_timing()
            ops.conv2d(x, wgt, b, pad=p, pad_mode=pm, upsample=up).backward(gy)
            for kn, d in ops.stop_kernel_timing().items():
                res.setdefault(kn, []).append(d["ms"])
        row = {"shape": name, "gflop": flops / 1e9}
        for kn, v in res.items():
            v.sort()
            ms = v[len(v) // 2]
            row[kn] = {"ms": ms, "tflops": flops / ms / 1e9}
        rows.append(row)
        print(name, f"{flops / 1e9:8.1f} GF ", "  ".join(f"{kn.replace('conv_', '')}: {d['ms']:.3f} ms {d['tflops']:.0f} TF" for kn, d in row.items() if isinstance(d, dict)), flush=True)
    if args.json:
        json.dump(rows, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
