"""CPU: the C-ABI library loads, exports every symbol include/affgw.h declares, the Python binding covers them all,
the drop-in classes reproduce the reference's checkpoint key layout, and the product never touches the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "affgw.h")).read()
    return sorted(set(re.findall(r"\b(affgw_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from affganwriting_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build with __graft_entry__.build()"
    h = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 45
    for n in names:
        assert hasattr(h, n), n
    assert h.affgw_version() == 111


def test_python_binding_covers_the_header():
    from affganwriting_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    _lib.lib()


def test_conv_descriptor_layout_matches_header():
    from affganwriting_b200 import _lib
    src = open(os.path.join(ROOT, "include", "affgw.h")).read()
    body = src[src.index("typedef struct affgw_conv_desc {"):src.index("} affgw_conv_desc;")]
    fields = []
    for line in body.splitlines()[1:]:
        line = line.split("/*")[0]
        if "int32_t" in line:
            fields += [f.strip() for f in line.replace("int32_t", "").replace(";", "").split(",") if f.strip()]
    assert fields == [f[0] for f in _lib.ConvDesc._fields_]
    assert ctypes.sizeof(_lib.ConvDesc) == 4 * len(fields)


def test_host_side_argument_validation_needs_no_gpu():
    from affganwriting_b200 import _lib
    h = _lib.lib()
    d = _lib.ConvDesc()
    assert h.affgw_conv_tc_supported(ctypes.byref(d)) == 0
    assert b"non-positive" in h.affgw_last_error()
    d.N, d.H, d.W, d.Cin, d.Cout, d.KH, d.KW = 2, 8, 27, 512, 512, 3, 3
    d.stride, d.pad, d.pad_mode, d.upsample, d.Ho, d.Wo = 1, 1, 1, 1, 8, 27
    d.in_pitch, d.out_pitch, d.x_dtype, d.w_dtype, d.y_dtype = 512, 512, 1, 1, 1
    d.algo, d.passes, d.grad_dtype = 2, 3, 0
    assert h.affgw_conv_tc_supported(ctypes.byref(d)) == 1
    assert h.affgw_conv2d_dgrad_ws_bytes(ctypes.byref(d)) == 2 * 10 * 29 * 512 * 4     # reflect: folded path, fp32 dx
    d.pad_mode = 0
    assert h.affgw_conv2d_dgrad_ws_bytes(ctypes.byref(d)) == 0                          # zero pad: direct
    d.passes = 2
    assert h.affgw_conv_tc_supported(ctypes.byref(d)) == 0 and b"passes" in h.affgw_last_error()
    d.passes = 1
    d.Ho = 9
    assert h.affgw_conv_tc_supported(ctypes.byref(d)) == 0 and b"output extent" in h.affgw_last_error()
    for layout in (1, 2):                                                               # im2col tiles / shifted-kernel planes
        assert h.affgw_pack_weight_tc_bytes(512, 512, 3, 3, 512, 0, 1, layout) == 512 * 512 * 9 * 2
        assert h.affgw_pack_weight_tc_bytes(512, 512, 3, 3, 512, 0, 3, layout) == 512 * 512 * 9 * 2 * 2
        assert h.affgw_pack_weight_tc_bytes(512, 512, 3, 3, 510, 0, 1, layout) < 0     # c_store must be a multiple of 8
    assert h.affgw_pack_weight_tc_bytes(512, 512, 3, 3, 512, 0, 1, 7) < 0
    d.Ho, d.passes = 8, 3
    assert h.affgw_conv_tc_layout(ctypes.byref(d), 0) == 2 and h.affgw_conv_tc_layout(ctypes.byref(d), 1) == 2
    assert h.affgw_conv_tc_prefer_shift(0) == 1
    assert h.affgw_conv_tc_layout(ctypes.byref(d), 0) == 1
    assert h.affgw_conv_tc_prefer_shift(1) == 0
    d.stride, d.Ho, d.Wo = 2, 4, 14                                                      # strided: im2col kernel only
    assert h.affgw_conv_tc_layout(ctypes.byref(d), 0) == 1
    assert h.affgw_operand_planes_bytes(100, 56, 3) == 100 * 56 * 2 * 2


@pytest.mark.parametrize("key,build", [("gen_c50", "gen"), ("dis", "dis"), ("cla", "cla")])
def test_state_dict_keys_match_reference(key, build, specs):
    from affganwriting_b200 import modules_tro as M
    m = {"gen": lambda: M.GenModel_FC(12), "dis": M.DisModel, "cla": lambda: M.WriterClaModel(500)}[build]()
    mine = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert mine == specs[key]
    assert list(mine) == list(specs[key])          # same registration order, as torch.save would write it


def test_block_constructor_signatures(specs):
    from affganwriting_b200 import blocks as B
    for name, sp in specs.items():
        if not name.startswith("blocks.conv_"):
            continue
        m = B.Conv2dBlock(**sp["ctor"])
        assert {k: list(v.shape) for k, v in m.state_dict().items()} == sp["spec"], name
    m = B.AdaptiveInstanceNorm2d(512)
    assert {k: list(v.shape) for k, v in m.state_dict().items()} == specs["blocks.adain_plain"]["spec"]
    with pytest.raises(AssertionError):
        B.Conv2dBlock(4, 4, 3, 1, norm="groupnorm")
    with pytest.raises(AssertionError):
        m(__import__("torch").zeros(1, 512, 2, 2))      # "Please assign AdaIN weight first"


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "affganwriting_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("oracle/", "").replace("CPU oracle", "") or fn == "__never__", fn


def test_missing_library_fails_loudly(monkeypatch):
    from affganwriting_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libaffgw.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.lib()


def _conv_desc(n, h, w, cin, cout, k, pad, pad_mode=1, stride=1, stride_w=0, passes=3, upsample=1):
    from affganwriting_b200 import _lib
    d = _lib.ConvDesc()
    d.N, d.H, d.W, d.Cin, d.Cout, d.KH, d.KW = n, h, w, cin, cout, k, k
    sw = stride_w or stride
    d.stride, d.stride_w, d.pad, d.pad_mode, d.upsample = stride, stride_w, pad, pad_mode, upsample
    d.Ho, d.Wo = (h * upsample + 2 * pad - k) // stride + 1, (w * upsample + 2 * pad - k) // sw + 1
    c8 = lambda c: (c + 7) // 8 * 8                                               # noqa: E731
    d.in_pitch, d.out_pitch, d.x_dtype, d.w_dtype, d.y_dtype = c8(cin), c8(cout), 1, 1, 1
    d.algo, d.passes, d.grad_dtype = 2, passes, 0
    return d


def test_kernel_routing_of_the_step_shapes_needs_no_gpu():
    """Which tcgen05 kernel a layer lands on: position-space kernels for stride-1 filters unless the padding
    overhead is too large, 128-position tiles only where 256-position tiles would leave SMs idle, wide-N tiles by channel
    count; the descriptor's column stride (Resnet18.py's (2, 1) stem) is honoured by the geometry check."""
    from affganwriting_b200 import _lib
    h = _lib.lib()
    SHIFT, IM2COL = _lib.WLAYOUT_SHIFT, _lib.WLAYOUT_IM2COL
    lay = lambda d, dg=0: h.affgw_conv_tc_layout(ctypes.byref(d), dg)             # noqa: E731
    tn = lambda d, which: h.affgw_conv_tc_tile_n(ctypes.byref(d), which)          # noqa: E731
    tm = lambda d, which: h.affgw_conv_tc_tile_m(ctypes.byref(d), which)          # noqa: E731
    vgg = _conv_desc(64, 32, 108, 256, 256, 3, 1, pad_mode=0)
    assert lay(vgg) == SHIFT and tn(vgg, 0) == 256 and tn(vgg, 1) == 256 and tn(vgg, 2) == 128 and tm(vgg, 0) == 256
    res = _conv_desc(64, 8, 27, 512, 512, 3, 1)                                   # ~146 tiles of 256 positions: keep them
    assert lay(res) == SHIFT and tm(res, 0) == 256 and tm(res, 1) == 256
    d414 = _conv_desc(128, 4, 14, 256, 256, 3, 1)                                 # wide layer on a tiny map: 1.7x padding accepted,
    assert lay(d414) == SHIFT and tn(d414, 0) == 256 and tm(d414, 0) == 128       # 128-position tiles
    d27 = _conv_desc(128, 2, 7, 512, 512, 3, 1)                                   # 2.6x padding: im2col kernels
    assert lay(d27) == IM2COL and tn(d27, 0) == 128 and tm(d27, 0) == 128
    thin = _conv_desc(128, 64, 216, 16, 16, 3, 1)
    assert lay(thin) == SHIFT and tn(thin, 0) == 16 and tn(thin, 2) == 16 and tm(thin, 0) == 512
    c64 = _conv_desc(64, 64, 216, 64, 64, 3, 1, pad_mode=0)
    assert tn(c64, 0) == 64 and tm(c64, 0) == 256                                 # N-packed hi/lo weights: 2 M-tiles
    assert tm(_conv_desc(64, 64, 216, 64, 64, 3, 1, pad_mode=0, passes=1), 0) == 512
    mix = _conv_desc(64, 8, 27, 1024, 512, 1, 0, pad_mode=0)
    assert lay(mix) == IM2COL
    shortcut = _conv_desc(128, 64, 216, 16, 32, 1, 0, pad_mode=0)                 # thin full-resolution 1x1: bulk-copy pipeline
    assert lay(shortcut) == SHIFT
    up = _conv_desc(64, 8, 27, 512, 256, 5, 2, upsample=2)
    assert lay(up) == SHIFT and up.Ho == 16 and up.Wo == 54
    stem = _conv_desc(2, 64, 216, 56, 96, 3, 1, pad_mode=0, stride=2, stride_w=1)  # rows / 2, columns kept
    assert (stem.Ho, stem.Wo) == (32, 216) and h.affgw_conv_tc_supported(ctypes.byref(stem)) == 1 and lay(stem) == IM2COL
    stem.Wo = 108                                                                  # the extent a (2, 2) stride would give
    assert h.affgw_conv_tc_supported(ctypes.byref(stem)) == 0 and b"output extent" in h.affgw_last_error()


def test_host_side_validation_of_optimiser_and_wire_format_needs_no_gpu():
    import torch
    from affganwriting_b200 import load_data as LD
    from affganwriting_b200.optim import Adam
    with pytest.raises(ValueError):
        Adam([])
    p = torch.zeros(3, requires_grad=True)
    for bad in (dict(lr=-1.0), dict(eps=-1e-8), dict(betas=(1.0, 0.999)), dict(betas=(0.9, -0.1))):
        with pytest.raises(ValueError):
            Adam([p], **bad)
    opt = Adam([p], lr=2e-4)
    assert opt.param_groups[0]["lr"] == 2e-4 and opt.param_groups[0]["betas"] == (0.9, 0.999) and opt.param_groups[0]["eps"] == 1e-8
    opt.step()                                                   # no gradient anywhere: nothing to do, no CUDA needed
    assert opt.state == {} and opt.state_dict() == {"state": {}, "param_groups": [dict(
        lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, params=[0])]}
    p.grad = torch.ones(3)
    with pytest.raises(RuntimeError):
        opt.step()                                               # CPU tensors: the product has no CPU path
    opt.zero_grad()
    assert p.grad is None
    ref = torch.optim.Adam([torch.zeros(3, requires_grad=True), torch.zeros(2, requires_grad=True)])
    with pytest.raises(ValueError):
        opt.load_state_dict(ref.state_dict())                    # parameter group of another size
    with pytest.raises(RuntimeError):
        LD.decode_u8(torch.zeros(16, dtype=torch.uint8))         # CPU tensor
    assert LD.batch_to_device(("src", 3), "cpu") == ("src", 3)   # non-tensor members pass through untouched


def test_pass_selection_contexts_need_no_gpu():
    """Host-side state of the per-layer tensor-core pass selection: ops.conv_passes and ops.relaxed_forward (the no_grad
    generator forward of dis_update / GraphedGenerator(relaxed=True)) - mode gating, nesting, restoration."""
    from affganwriting_b200 import ops
    from affganwriting_b200.vgg_tro_channel3_modi import RELAXED_FROM
    assert ops.precision() == "fp32" and not ops.relaxed()
    with ops.relaxed_forward(True):
        assert not ops.relaxed()                              # fp32 / bf16 modes: never relaxed
    ops.set_precision("bf16")
    try:
        with ops.relaxed_forward(True):
            assert not ops.relaxed()
        ops.set_precision("f16")
        assert ops._state["passes"] == (3, 1, 1)
        with ops.relaxed_forward(False):
            assert not ops.relaxed()
        with ops.relaxed_forward(True):
            assert ops.relaxed() and ops.relaxed_from(RELAXED_FROM) == RELAXED_FROM
            with ops.relaxed_forward(True, 26):               # generation: a shallower relaxation (inference.GraphedGenerator)
                assert ops.relaxed_from(RELAXED_FROM) == 26
                with ops.conv_passes(fwd=1):
                    assert ops._state["passes"] == (1, 1, 1)
                with ops.conv_passes(fwd=None):               # Conv2dBlock.fwd_passes = None keeps the mode's counts
                    assert ops._state["passes"] == (3, 1, 1)
                assert ops._state["passes"] == (3, 1, 1)
            assert ops.relaxed_from(RELAXED_FROM) == RELAXED_FROM
        assert not ops.relaxed()
        # the VGG layers that stay at three passes are the first five convolutions (features 0, 3, 6, 9, 13)
        assert RELAXED_FROM == 14
    finally:
        ops.set_precision("fp32")


def test_compute_entry_points_fail_loudly_without_a_device():
    """No CPU fallback (include/affgw.h conventions): on a host without a GPU the library reports that and a compute entry
    point returns a negative code with a message - it never touches the (host) buffers it was handed."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("only meaningful on a GPU-less host")
    from affganwriting_b200 import _lib
    h = _lib.lib()
    assert h.affgw_device_ok() == 0 and h.affgw_last_error()
    buf = (ctypes.c_float * 64)(*([1.5] * 64))
    p = ctypes.addressof(buf)
    with pytest.raises(RuntimeError, match="affgw_act_bwd failed"):
        _lib.call("affgw_act_bwd", p, p, p, 0, 64, 1, None)
    assert list(buf) == [1.5] * 64


@pytest.mark.parametrize("shape", [(32, 108, 256, 256, 3, 1, 1), (8, 27, 512, 512, 3, 1, 1), (64, 216, 16, 16, 3, 1, 1),
                                   (8, 27, 512, 256, 5, 2, 2), (64, 216, 50, 64, 3, 1, 1), (4, 14, 256, 256, 3, 1, 1)])
def test_position_frames_hold_every_window_for_any_batch(shape):
    """Storage geometry of the planar operand planes (host code - no GPU): for every batch size a short final
    batch can have (and the 2B / 4B concatenations of the discriminator passes), the frame a position-space kernel reads keeps
    (a) a zero lead as long as the filter window's reach backwards, (b) room behind the last position for the rounding of the
    position range to whole CTA tiles plus the window's reach forwards, (c) 8-position (one 128-byte core matrix) granularity."""
    from affganwriting_b200 import _lib
    h = _lib.lib()
    H, W, cin, cout, k, pad, up = shape
    for n in list(range(1, 66)) + [127, 128, 192, 255, 256]:
        d = _conv_desc(n, H, W, cin, cout, k, pad, upsample=up)
        if h.affgw_conv_tc_layout(ctypes.byref(d), 0) != _lib.WLAYOUT_SHIFT:
            continue
        fx, fy = _lib.PosFrame(), _lib.PosFrame()
        assert h.affgw_conv_pos_frames(ctypes.byref(d), ctypes.byref(fx), ctypes.byref(fy)) == 0
        for f, c in ((fx, cin), (fy, cout)):
            assert (f.N, f.Hp, f.Wp) == (n, H * up + 2 * pad, W * up + 2 * pad)
            assert f.G == 2 * ((c + 15) // 16)                                     # 8-channel groups, even
            span = (k - 1) * (f.Wp + 1)                                            # reach of the filter window in positions
            q = f.N * f.Hp * f.Wp
            tile = max(h.affgw_conv_tc_tile_m(ctypes.byref(d), 0), h.affgw_conv_tc_tile_m(ctypes.byref(d), 1))
            assert f.lead >= span and f.lead % 8 == 0
            assert f.QA % 32 == 0 and f.QA - f.lead >= (q + tile - 1) // tile * tile + span
            for passes in (1, 3):
                assert h.affgw_position_planes_bytes(ctypes.byref(f), passes) == (2 if passes == 3 else 1) * f.G * f.QA * 16
