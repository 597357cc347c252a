"""tcgen05 implicit-GEMM kernel vs the CUDA-core kernel on identical bf16 operands (forward and dgrad), through the
public entry points.  Both accumulate in fp32, so they must agree to accumulation-order noise; this is the check
that pins the UMMA shared-memory / instruction descriptors and the software 128B swizzle."""
import pytest
import torch

import affganwriting_b200 as A
from affganwriting_b200 import ops
from affgw_testutil import rel_err

pytestmark = pytest.mark.gpu

CASES = [
    # N, H, W, Cin, Cout, k, pad, pad_mode, upsample
    (2, 8, 27, 64, 64, 3, 1, "zero", 1),
    (2, 8, 27, 128, 128, 3, 1, "reflect", 1),
    (3, 16, 54, 64, 128, 3, 1, "zero", 1),
    (2, 8, 27, 512, 512, 3, 1, "reflect", 1),
    (2, 8, 27, 512, 256, 5, 2, "reflect", 2),
    (1, 32, 108, 128, 64, 5, 2, "reflect", 2),
    (2, 8, 27, 1024, 512, 1, 0, "zero", 1),
    (5, 7, 9, 64, 192, 3, 1, "replicate", 1),
    (1, 64, 216, 64, 64, 3, 1, "zero", 1),
]


@pytest.mark.parametrize("case", CASES)
def test_tc_matches_simt(case):
    n, h, w, ci, co, k, p, pm, up = case
    A.set_precision("bf16")
    try:
        g = torch.Generator(device="cuda").manual_seed(hash(case) & 0xFFFF)
        x = ops.to_internal(torch.randn(n, ci, h, w, device="cuda", generator=g)).requires_grad_()
        wgt = (torch.randn(co, ci, k, k, device="cuda", generator=g) * (2.0 / (ci * k * k)) ** 0.5).requires_grad_()
        b = torch.randn(co, device="cuda", generator=g)
        outs = {}
        for simt in (True, False):
            A.force_simt(simt)
            x.grad = None
            y = ops.conv2d(x, wgt, b, pad=p, pad_mode=pm, upsample=up, post_act="relu")
            gy = torch.randn(y.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(7))
            y.float().backward(gy)
            outs[simt] = (y.detach().float(), x.grad.detach().float())
        assert rel_err(outs[False][0], outs[True][0]) <= 1e-2
        assert rel_err(outs[False][1], outs[True][1]) <= 3e-2     # dgrad through a bf16 padded-gradient buffer
        # and against an fp32 torch convolution of the same (bf16-rounded) operands
        xr = x.detach().float()
        if up == 2:
            xr = torch.nn.functional.interpolate(xr, scale_factor=2)
        if p:
            xr = torch.nn.functional.pad(xr, (p, p, p, p), mode={"zero": "constant"}.get(pm, pm))
        ref = torch.relu(torch.nn.functional.conv2d(xr, wgt.detach().bfloat16().float(), b))
        assert rel_err(outs[False][0], ref) <= 1e-2
    finally:
        A.force_simt(False)
        A.set_precision("fp32")


def test_tc_path_is_taken():
    """The bf16 route must really launch the tcgen05 kernel for 64-aligned channel counts."""
    import ctypes
    from affganwriting_b200 import _lib as L
    d = L.ConvDesc()
    d.N, d.H, d.W, d.Cin, d.Cout, d.KH, d.KW = 2, 8, 27, 512, 512, 3, 3
    d.stride, d.pad, d.pad_mode, d.upsample, d.Ho, d.Wo = 1, 1, 1, 1, 8, 27
    d.in_pitch, d.out_pitch, d.x_dtype, d.w_dtype, d.y_dtype = 512, 512, 1, 1, 1
    assert L.lib().affgw_conv_tc_block_n(ctypes.byref(d)) == 128
    d.Cout, d.out_pitch = 64, 64
    assert L.lib().affgw_conv_tc_block_n(ctypes.byref(d)) == 64
    d.Cin, d.in_pitch = 50, 50
    assert L.lib().affgw_conv_tc_block_n(ctypes.byref(d)) == 0


BLOCKS = [
    # in, out, k, pad, norm, act, pad_type, H, W
    (128, 128, 3, 1, "in", "relu", "reflect", 8, 27),
    (64, 64, 3, 1, "in", "relu", "zero", 16, 54),
    (256, 128, 5, 2, "in", "relu", "reflect", 8, 27),
    (128, 64, 1, 0, "none", "none", "zero", 8, 27),
]


@pytest.mark.parametrize("case", BLOCKS)
def test_tc_conv2dblock_against_bf16_storage_oracle(case):
    """One Conv2dBlock on tensor-core-eligible channel counts, forward + backward, against the oracle under the bf16
    storage model: a single block is not chaotic, so this is the tight numerical check of the tcgen05 path
    (image-level bounds are dominated by the network's own sensitivity, see tests/test_gpu_models.py)."""
    from affganwriting_b200.blocks import Conv2dBlock
    from affgw_testutil import cosine
    from oracle import affgw_oracle as O
    from oracle import weights as W
    ci, co, k, p, norm, act, pt, h, w = case
    A.set_precision("bf16")
    try:
        m = Conv2dBlock(ci, co, k, 1, p, norm=norm, activation=act, pad_type=pt)
        sd = W.make_state({kk: list(v.shape) for kk, v in m.state_dict().items()})
        m.load_state_dict(sd)
        m = m.cuda()
        g = torch.Generator().manual_seed(3)
        x_cpu = torch.randn(3, ci, h, w, generator=g).bfloat16().float()
        gy_cpu = torch.randn(3, co, h, w, generator=g)
        x = x_cpu.cuda().requires_grad_()
        y = m(x)
        y.float().backward(gy_cpu.cuda())
        xo = x_cpu.clone().requires_grad_()
        sdo = {kk: v.clone().requires_grad_() for kk, v in sd.items()}
        with O.storage_model("bf16"):
            yo = O.conv2d_block(xo, sdo, "", k, 1, p, norm, act, pt)
        yo.backward(gy_cpu)
        assert rel_err(y, yo) <= 1e-2
        assert cosine(x.grad, xo.grad) >= 0.999
        assert cosine(m.conv.weight.grad, sdo["conv.weight"].grad) >= 0.999
        assert rel_err(m.conv.weight.grad, sdo["conv.weight"].grad) <= 3e-2
    finally:
        A.set_precision("fp32")


WG_CASES = [
    # N, H, W, Cin, Cout, k, pad, pad_mode, upsample
    (2, 8, 27, 64, 64, 3, 1, "zero", 1),          # Cin = 64: two taps share one 128-row M tile, odd tap count
    (2, 8, 27, 128, 128, 3, 1, "reflect", 1),
    (3, 16, 54, 64, 128, 3, 1, "zero", 1),
    (2, 8, 27, 512, 256, 5, 2, "reflect", 2),
    (2, 8, 27, 1024, 512, 1, 0, "zero", 1),
    (1, 64, 216, 64, 64, 3, 1, "zero", 1),        # many pixel splits
    (5, 7, 9, 192, 64, 3, 1, "replicate", 1),
]


@pytest.mark.parametrize("case", WG_CASES)
def test_tc_wgrad_matches_simt(case):
    """MN-major tcgen05 weight-gradient kernel vs the CUDA-core wgrad on identical bf16 operands."""
    from affgw_testutil import cosine
    n, h, w, ci, co, k, p, pm, up = case
    A.set_precision("bf16")
    try:
        g = torch.Generator(device="cuda").manual_seed(11)
        x = ops.to_internal(torch.randn(n, ci, h, w, device="cuda", generator=g))
        wgt = (torch.randn(co, ci, k, k, device="cuda", generator=g) * 0.05).requires_grad_()
        grads = {}
        for simt in (True, False):
            ops.force_simt_wgrad(simt)
            wgt.grad = None
            y = ops.conv2d(x, wgt, None, pad=p, pad_mode=pm, upsample=up)
            gy = torch.randn(y.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
            y.float().backward(gy)
            grads[simt] = wgt.grad.detach().clone()
        assert cosine(grads[False], grads[True]) >= 0.99999
        assert rel_err(grads[False], grads[True]) <= 2e-3
    finally:
        ops.force_simt_wgrad(False)
        A.set_precision("fp32")
