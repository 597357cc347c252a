"""GPU sanity check of the tcgen05 conv kernels (fwd / dgrad / wgrad, 1 and 3 passes) against torch fp32 convolutions."""
import os, sys
import torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import affganwriting_b200 as A
from affganwriting_b200 import ops
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
CASES = [
    # N, H, W, Cin, Cout, k, stride, pad, pad_mode, upsample, pre_act
    (2, 8, 27, 64, 64, 3, 1, 1, "zero", 1, "none"),
    (2, 8, 27, 128, 128, 3, 1, 1, "reflect", 1, "none"),
    (3, 16, 54, 64, 128, 3, 1, 1, "zero", 1, "none"),
    (2, 8, 27, 512, 256, 5, 1, 2, "reflect", 2, "none"),
    (2, 8, 27, 1024, 512, 1, 1, 0, "zero", 1, "none"),
    (2, 64, 216, 15, 64, 3, 1, 1, "zero", 1, "none"),
    (2, 64, 216, 50, 64, 3, 1, 1, "zero", 1, "none"),
    (2, 64, 216, 1, 16, 7, 1, 3, "reflect", 1, "none"),
    (2, 64, 216, 16, 16, 3, 1, 1, "reflect", 1, "lrelu"),
    (2, 32, 108, 16, 32, 3, 1, 1, "reflect", 1, "lrelu"),
    (2, 32, 108, 16, 32, 1, 1, 0, "zero", 1, "none"),
    (2, 16, 54, 32, 64, 3, 1, 1, "reflect", 1, "lrelu"),
    (2, 64, 216, 64, 1, 7, 1, 3, "reflect", 1, "none"),
    (4, 2, 7, 1024, 500, 2, 7, 0, "zero", 1, "lrelu"),
    (4, 2, 7, 512, 1024, 3, 1, 1, "reflect", 1, "lrelu"),
    (5, 7, 9, 192, 72, 3, 1, 1, "replicate", 1, "none"),
]
def ref_conv(x, w, b, k, s, p, pm, up, pre):
    if pre == "lrelu": x = F.leaky_relu(x, 0.2)
    if up == 2: x = F.interpolate(x, scale_factor=2)
    if p: x = F.pad(x, (p, p, p, p), mode={"zero": "constant"}.get(pm, pm))
    return F.conv2d(x, w, b, stride=s)
def rel(a, b): return float((a.detach().double() - b.detach().double()).abs().max() / (b.detach().double().abs().max() + 1e-30))
bad = 0
for mode, tol in (("bf16x1", 2e-2), ("bf16", 2e-4)):
    A.set_precision(mode)
    for case in CASES:
        n, h, w_, ci, co, k, s, p, pm, up, pre = case
        g = torch.Generator(device="cuda").manual_seed(1)
        x = torch.randn(n, ci, h, w_, device="cuda", generator=g)
        wgt = torch.randn(co, ci, k, k, device="cuda", generator=g) * (2.0 / (ci * k * k)) ** 0.5
        b = torch.randn(co, device="cuda", generator=g)
        xi = ops.to_internal(x).detach().clone().requires_grad_(); wi = wgt.clone().requires_grad_(); bi = b.clone().requires_grad_()
        y = ops.conv2d(xi, wi, bi, stride=s, pad=p, pad_mode=pm, upsample=up, pre_act=pre)
        xr = x.double().requires_grad_(); wr = wgt.double().requires_grad_(); br = b.double().requires_grad_()
        yr = ref_conv(xr, wr, br, k, s, p, pm, up, pre)
        gy = torch.randn(yr.shape, device="cuda", generator=g)
        y.backward(gy.float()); yr.backward(gy)
        e = (rel(y, yr), rel(xi.grad, xr.grad), rel(wi.grad, wr.grad), rel(bi.grad, br.grad))
        ok = all(v <= tol for v in e)
        bad += not ok
        print(f"{mode:7s} {str(case):70s} y {e[0]:.1e} dx {e[1]:.1e} dw {e[2]:.1e} db {e[3]:.1e} {'ok' if ok else 'FAIL'}", flush=True)
    # 2-D linear
    x = torch.randn(64, 768, device="cuda"); wgt = torch.randn(1024, 768, device="cuda") * 0.03; b = torch.randn(1024, device="cuda")
    xi = x.clone().requires_grad_(); wi = wgt.clone().requires_grad_()
    y = ops.linear(xi, wi, b); yr = F.linear(x, wgt, b)
    y.backward(torch.ones_like(y))
    e = (rel(y, yr), rel(xi.grad, torch.ones_like(yr) @ wgt), rel(wi.grad, torch.ones_like(yr).t() @ x))
    ok = all(v <= tol for v in e); bad += not ok
    print(f"{mode:7s} linear 64x768->1024 y {e[0]:.1e} dx {e[1]:.1e} dw {e[2]:.1e} {'ok' if ok else 'FAIL'}")
A.set_precision("fp32")
print("FAILURES", bad)
sys.exit(1 if bad else 0)
