"""tcgen05 implicit-GEMM convolution kernels (forward, dgrad, MN-major wgrad; 1 and 3 MMA passes) through the public
entry points, against (a) a plain torch float64 convolution of the same op on the same inputs and (b) the CUDA-core
kernels of this library.  This is the check that pins the UMMA shared-memory / instruction descriptors, the software
128B swizzle, the in-gather padding / upsampling / zero-insertion index maps and the split-bf16 operand planes.

Tolerances (relative to the largest reference magnitude): 3 passes (mode 'bf16x3') 2e-4 - the split keeps ~16 mantissa
bits per operand; 1 pass (mode 'bf16x1') 2e-2 - one bf16 rounding per operand; the modes 'f16' (the one
bench.py measures: fp16 operand planes, 11 bits each) and 'bf16' run the forward GEMM with 3 passes and both backward GEMMs
with 1, so y is held to the first bar and dx / dw to a single-rounding bar (3e-3 for fp16, 2e-2 for bf16)."""
import pytest
import torch
import torch.nn.functional as F

import affganwriting_b200 as A
from affganwriting_b200 import ops
from affgw_testutil import cosine

pytestmark = pytest.mark.gpu

CASES = [
    # N, H, W, Cin, Cout, k, stride, pad, pad_mode, upsample, pre_act
    (2, 8, 27, 64, 64, 3, 1, 1, "zero", 1, "none"),
    (2, 8, 27, 128, 128, 3, 1, 1, "reflect", 1, "none"),
    (3, 16, 54, 64, 128, 3, 1, 1, "zero", 1, "none"),
    (2, 8, 27, 512, 512, 3, 1, 1, "reflect", 1, "none"),          # decoder ResBlock
    (2, 8, 27, 512, 256, 5, 1, 2, "reflect", 2, "none"),          # decoder up-conv (nearest x2 folded into the gather)
    (1, 32, 108, 128, 64, 5, 1, 2, "reflect", 2, "none"),
    (2, 8, 27, 1024, 512, 1, 1, 0, "zero", 1, "none"),            # GenModel_FC.mix
    (2, 64, 216, 15, 64, 3, 1, 1, "zero", 1, "none"),             # first VGG conv, config 1 (channels padded to 16)
    (2, 64, 216, 50, 64, 3, 1, 1, "zero", 1, "none"),             # first VGG conv, config 2 (channels padded to 56)
    (2, 64, 216, 1, 16, 7, 1, 3, "reflect", 1, "none"),           # Dis / Cla stem
    (2, 64, 216, 16, 16, 3, 1, 1, "reflect", 1, "lrelu"),         # ActFirstResBlock, 16 channels
    (2, 32, 108, 16, 32, 1, 1, 0, "zero", 1, "none"),             # learned shortcut
    (2, 16, 54, 32, 64, 3, 1, 1, "reflect", 1, "lrelu"),
    (2, 64, 216, 64, 1, 7, 1, 3, "reflect", 1, "none"),           # decoder output conv (Cout = 1)
    (4, 2, 7, 1024, 500, 2, 7, 0, "zero", 1, "lrelu"),            # head: kernel 2, stride 7 (zero-insertion dgrad)
    (4, 2, 7, 512, 1024, 3, 1, 1, "reflect", 1, "lrelu"),
    (6, 4, 14, 256, 256, 3, 1, 1, "reflect", 1, "lrelu"),         # tiny map, wide layer: 128-position tiles (MT = 1)
    (3, 4, 14, 256, 512, 3, 1, 1, "zero", 1, "none"),             # same for a zero-padded layer (ResNet stage 3/4 shapes)
    (5, 7, 9, 192, 72, 3, 1, 1, "replicate", 1, "none"),          # ragged everything
    (1, 64, 216, 64, 64, 3, 1, 1, "zero", 1, "none"),             # many pixel splits in wgrad
]
TOL = {"bf16x3": 2e-4, "bf16x1": 2e-2, "bf16": 2e-4, "f16": 2e-4}
TOL_BWD = {"bf16x3": 2e-4, "bf16x1": 2e-2, "bf16": 2e-2, "f16": 3e-3}


def ref_conv(x, w, b, s, p, pm, up, pre):
    if pre == "lrelu":
        x = F.leaky_relu(x, 0.2)
    if up == 2:
        x = F.interpolate(x, scale_factor=2)
    if p:
        x = F.pad(x, (p, p, p, p), mode={"zero": "constant"}.get(pm, pm))
    return F.conv2d(x, w, b, stride=s)


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


@pytest.fixture(params=["bf16x3", "bf16x1", "bf16", "f16"])
def tc_mode(request):
    A.set_precision(request.param)
    A.force_simt(False)
    ops.force_simt_wgrad(False)
    yield request.param
    A.force_simt(False)
    ops.force_simt_wgrad(False)
    A.set_precision("fp32")


@pytest.fixture(params=[True, False], ids=["thin", "tensorcore"])
def thin_route(request):
    """Single-channel-sided stencils have their own fp32 CUDA-core kernels; both routes are checked."""
    prev = ops.use_thin_kernels(request.param)
    yield request.param
    ops.use_thin_kernels(prev)


@pytest.mark.parametrize("case", [CASES[9], CASES[13], (3, 20, 45, 1, 64, 5, 1, 2, "zero", 1, "none"),
                                  (3, 20, 45, 16, 1, 3, 1, 1, "replicate", 1, "none")])
def test_single_channel_stencils(case, thin_route, tc_mode):
    """conv_thin.cu (1 -> N and N -> 1 stencils, forward / dgrad / wgrad) and the tensor-core route for the same layers."""
    if not thin_route and case[5] != 7:
        pytest.skip("covered by the generic cases")
    n, h, w_, ci, co, k, s, p, pm, up, pre = case
    # (one bf16 rounding per operand through a tanh on a 64-channel 7x7 sum: looser than the linear cases)
    tol = TOL[tc_mode] if (thin_route or tc_mode == "bf16x3") else (3e-3 if tc_mode == "f16" else 6e-2)
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.randn(n, ci, h, w_, device="cuda", generator=g)
    wgt = torch.randn(co, ci, k, k, device="cuda", generator=g) * (2.0 / (ci * k * k)) ** 0.5
    b = torch.randn(co, device="cuda", generator=g)
    xi = ops.to_internal(x).detach().clone().requires_grad_()
    wi, bi = wgt.clone().requires_grad_(), b.clone().requires_grad_()
    ops.start_kernel_timing()
    y = ops.conv2d(xi, wi, bi, stride=s, pad=p, pad_mode=pm, upsample=up, pre_act=pre, post_act="tanh")
    xr, wr, br = x.double().requires_grad_(), wgt.double().requires_grad_(), b.double().requires_grad_()
    yr = torch.tanh(ref_conv(xr, wr, br, s, p, pm, up, pre))
    gy = torch.randn(yr.shape, device="cuda", generator=g)
    y.backward(gy)
    names = set(ops.stop_kernel_timing())
    assert any("thin" in nm for nm in names) == thin_route, names
    yr.backward(gy.double())
    e = dict(y=rel(y, yr), dx=rel(xi.grad, xr.grad), dw=rel(wi.grad, wr.grad), db=rel(bi.grad, br.grad))
    assert all(v <= tol for v in e.values()), e


@pytest.mark.parametrize("case", CASES)
def test_tc_conv_matches_float64_torch(case, tc_mode):
    n, h, w_, ci, co, k, s, p, pm, up, pre = case
    tol = TOL[tc_mode]
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(n, ci, h, w_, device="cuda", generator=g)
    wgt = torch.randn(co, ci, k, k, device="cuda", generator=g) * (2.0 / (ci * k * k)) ** 0.5
    b = torch.randn(co, device="cuda", generator=g)
    xi = ops.to_internal(x).detach().clone().requires_grad_()
    wi, bi = wgt.clone().requires_grad_(), b.clone().requires_grad_()
    recs_before = A.launch_count()
    y = ops.conv2d(xi, wi, bi, stride=s, pad=p, pad_mode=pm, upsample=up, pre_act=pre)
    assert A.launch_count() > recs_before
    xr, wr, br = x.double().requires_grad_(), wgt.double().requires_grad_(), b.double().requires_grad_()
    yr = ref_conv(xr, wr, br, s, p, pm, up, pre)
    assert y.shape == yr.shape and y.dtype == torch.float32
    gy = torch.randn(yr.shape, device="cuda", generator=g)
    y.backward(gy)
    yr.backward(gy.double())
    e = dict(y=rel(y, yr), dx=rel(xi.grad, xr.grad), dw=rel(wi.grad, wr.grad), db=rel(bi.grad, br.grad))
    assert e["y"] <= tol and e["db"] <= tol and e["dx"] <= TOL_BWD[tc_mode] and e["dw"] <= TOL_BWD[tc_mode], e
    assert cosine(wi.grad, wr.grad) >= (0.9999999 if tc_mode == "bf16x3" else 0.9999)


def test_tc_linear_matches_float64_torch(tc_mode):
    """nn.Linear (TextEncoder_FC.fc, modules_tro.py:272-282) rides the same kernels as a 1x1 convolution."""
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(64, 768, device="cuda", generator=g)
    wgt = torch.randn(1024, 768, device="cuda", generator=g) * 0.03
    b = torch.randn(1024, device="cuda", generator=g)
    xi, wi = x.clone().requires_grad_(), wgt.clone().requires_grad_()
    y = ops.linear(xi, wi, b, post_act="relu")
    yr = torch.relu(F.linear(x.double(), wgt.double(), b.double()))
    gy = torch.randn(y.shape, device="cuda", generator=g)
    y.backward(gy)
    gz = gy.double() * (y > 0)          # the mask of OUR forward: a sign flip of a near-zero pre-activation is not a dgrad error
    assert rel(y, yr) <= TOL[tc_mode]
    assert rel(xi.grad, gz @ wgt.double()) <= TOL_BWD[tc_mode]
    assert rel(wi.grad, gz.t() @ x.double()) <= TOL_BWD[tc_mode]


@pytest.mark.parametrize("case", [CASES[1], CASES[4], CASES[8], CASES[14], CASES[16]])
def test_tc_matches_cuda_core_kernels(case):
    """Same convolution through the fp32 CUDA-core kernels and through the 3-pass tcgen05 kernels of this library."""
    n, h, w_, ci, co, k, s, p, pm, up, pre = case
    A.set_precision("bf16x3")
    try:
        g = torch.Generator(device="cuda").manual_seed(3)
        x = ops.to_internal(torch.randn(n, ci, h, w_, device="cuda", generator=g))
        wgt = torch.randn(co, ci, k, k, device="cuda", generator=g) * (2.0 / (ci * k * k)) ** 0.5
        outs = {}
        for simt in (True, False):
            A.force_simt(simt)
            xi, wi = x.clone().requires_grad_(), wgt.clone().requires_grad_()
            y = ops.conv2d(xi, wi, None, stride=s, pad=p, pad_mode=pm, upsample=up, pre_act=pre)
            gy = torch.randn(y.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(7))
            y.backward(gy)
            outs[simt] = (y.detach(), xi.grad, wi.grad)
        for a, b_, what in zip(outs[False], outs[True], ("y", "dx", "dw")):
            assert rel(a, b_) <= 2e-4, what
    finally:
        A.force_simt(False)
        A.set_precision("fp32")


def test_tc_wgrad_matches_cuda_core_wgrad():
    """MN-major tcgen05 weight-gradient kernel vs the CUDA-core wgrad with forward / dgrad kept on tensor cores."""
    A.set_precision("bf16x3")
    try:
        for (n, h, w_, ci, co, k, s, p, pm, up, pre) in (CASES[0], CASES[4], CASES[17]):
            g = torch.Generator(device="cuda").manual_seed(11)
            x = ops.to_internal(torch.randn(n, ci, h, w_, device="cuda", generator=g))
            wgt = (torch.randn(co, ci, k, k, device="cuda", generator=g) * 0.05).requires_grad_()
            grads = {}
            for simt in (True, False):
                ops.force_simt_wgrad(simt)
                wgt.grad = None
                y = ops.conv2d(x, wgt, None, stride=s, pad=p, pad_mode=pm, upsample=up)
                gy = torch.randn(y.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
                y.backward(gy)
                grads[simt] = wgt.grad.detach().clone()
            assert cosine(grads[False], grads[True]) >= 0.9999999
            assert rel(grads[False], grads[True]) <= 2e-4
    finally:
        ops.force_simt_wgrad(False)
        A.set_precision("fp32")


def test_fp32_mode_never_uses_tensor_cores_and_bf16_always_does():
    from affganwriting_b200 import ops as O_
    x = ops.to_internal(torch.randn(2, 64, 8, 27, device="cuda")).requires_grad_()
    w = torch.randn(64, 64, 3, 3, device="cuda", requires_grad=True)
    for mode, want in (("fp32", "simt"), ("f16", "tcgen05"), ("bf16", "tcgen05"), ("bf16x3", "tcgen05"), ("bf16x1", "tcgen05")):
        A.set_precision(mode)
        O_.start_kernel_timing()
        ops.conv2d(x, w, None, pad=1).sum().backward()
        names = set(O_.stop_kernel_timing())
        assert names and all(want in n for n in names), (mode, names)
    A.set_precision("fp32")


@pytest.mark.parametrize("case", [CASES[3], CASES[4], CASES[5], CASES[1], CASES[0]])      # position-space layers
@pytest.mark.parametrize("grad_scale", [1.0, 3e-7])
def test_fp16_operand_planes(case, grad_scale):
    """ops.operand_format("f16"): fp16 operand planes, one MMA per product in forward, dgrad and wgrad (the decoder's route).
    11 significant bits per operand -> 2e-3 of the largest reference magnitude; tiny incoming gradients (3e-7: far below
    fp16's normal range) exercise the per-tensor power-of-two scale of the dY planes."""
    n, h, w_, ci, co, k, s, p, pm, up, pre = case
    A.set_precision("bf16")
    try:
        g = torch.Generator(device="cuda").manual_seed(21)
        x = torch.randn(n, ci, h, w_, device="cuda", generator=g)
        wgt = torch.randn(co, ci, k, k, device="cuda", generator=g) * (2.0 / (ci * k * k)) ** 0.5
        b = torch.randn(co, device="cuda", generator=g)
        xi = ops.to_internal(x).detach().clone().requires_grad_()
        wi, bi = wgt.clone().requires_grad_(), b.clone().requires_grad_()
        ops.start_kernel_timing()
        with ops.operand_format("f16"):
            y = ops.conv2d(xi, wi, bi, stride=s, pad=p, pad_mode=pm, upsample=up, pre_act=pre)
        xr, wr, br = x.double().requires_grad_(), wgt.double().requires_grad_(), b.double().requires_grad_()
        yr = ref_conv(xr, wr, br, s, p, pm, up, pre)
        gy = torch.randn(yr.shape, device="cuda", generator=g) * grad_scale
        y.backward(gy)
        names = set(ops.stop_kernel_timing(by_kernel=True))
        ops.stop_stream_timing()
        assert names and all(nm.endswith(" f16") and ", 1" in nm for nm in names), names
        yr.backward(gy.double())
        e = dict(y=rel(y, yr), dx=rel(xi.grad, xr.grad), dw=rel(wi.grad, wr.grad), db=rel(bi.grad, br.grad))
        assert e["y"] <= 2e-3 and e["dx"] <= 2e-3 and e["dw"] <= 2e-3 and e["db"] <= 2e-4, e
        assert cosine(wi.grad, wr.grad) >= 0.999999 and cosine(xi.grad, xr.grad) >= 0.999999
    finally:
        A.set_precision("fp32")
