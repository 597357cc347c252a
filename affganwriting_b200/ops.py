"""Autograd-visible operators over the libaffgw C ABI.

Every operator takes / returns ordinary torch tensors with the reference's logical shapes (N, C, H, W), stored
channels-last (NHWC in memory) so that no layout copies happen between operators.  PyTorch is used for tensor
allocation, streams and the autograd graph only: all arithmetic below runs in libaffgw kernels.
"""
import collections
import ctypes as C

import torch
from torch.autograd import Function

from . import _lib as L  # noqa: N812

import os as _os
_FUSE_DB = _os.environ.get("AFFGW_FUSE_DB", "1") != "0"
_THIN = _os.environ.get("AFFGW_THIN", "1") != "0"
_WGRAD_LATE = _os.environ.get("AFFGW_WGRAD_FORK", "late") == "late"
_WGRAD_PRIO = int(_os.environ.get("AFFGW_WGRAD_PRIO", "0"))
# "passes": tensor-core MMAs per product of (forward, input-gradient, weight-gradient) GEMMs: 3 = split operands, 1 = single
_state = {"mode": "fp32", "passes": (3, 3, 3), "force_simt": False, "simt_wgrad": False, "fmt": None, "grad_accum": False,
          "wgrad_side": None, "scratch_tag": None, "bn_record": None}
_MODES = {"fp32": (3, 3, 3), "f16": (3, 1, 1), "bf16": (3, 1, 1), "bf16x3": (3, 3, 3), "bf16x1": (1, 1, 1)}
_err_flag = {}
_profile = {"records": None}


def start_kernel_timing():
    """Begin recording (name, start_event, end_event, algorithmic_flops) for every convolution launch and (entry point, events,
    algorithmic bytes) for every streaming kernel; the events sit on the launching (current) stream.  Used by bench.py for the
    roofline numbers; off by default."""
    _profile["records"] = []
    L.PROFILE = []


def stop_stream_timing():
    """-> {entry point: {"launches", "ms", "bytes"}} of the streaming (HBM-bound) kernels since start_kernel_timing();
    call before stop_kernel_timing() (which synchronises) or after - it synchronises itself."""
    recs, L.PROFILE = L.PROFILE or [], None
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1, nb in recs:
        d = out.setdefault(name, {"launches": 0, "ms": 0.0, "bytes": 0})
        d["launches"] += 1
        d["ms"] += e0.elapsed_time(e1)
        d["bytes"] += nb
    return out


def stop_kernel_timing(by_shape=False, by_kernel=False):
    """-> {name: {"launches", "ms", "flops"}} ; synchronises the device once.  by_shape=True keys on (name, shape tag),
    by_kernel=True on the CUDA kernel function the call ran (e.g. conv_shift_tcgen05_kernel<256, 3>)."""
    recs, _profile["records"] = _profile["records"] or [], None
    if L.PROFILE is not None and not L.PROFILE:
        L.PROFILE = None
    torch.cuda.synchronize()
    out = {}
    for name, tag, e0, e1, fl in recs:
        kern = name
        if isinstance(tag, tuple):
            tag, kern = tag
        key = kern if by_kernel else ((name, tag) if by_shape else name)
        d = out.setdefault(key, {"launches": 0, "ms": 0.0, "flops": 0.0, "kind": name})
        d["launches"] += 1
        d["ms"] += e0.elapsed_time(e1)
        d["flops"] += fl
    return out


def _kernel_name(d, which, passes):
    """CUDA kernel function a tcgen05 convolution call lands on (matches the names in an ncu / CUPTI launch list; calls on
    fp16 operand planes run the same functions and are listed separately with an " f16" suffix)."""
    lay = L.lib().affgw_conv_tc_layout(C.byref(d), int(which == 1))
    bn = L.lib().affgw_conv_tc_tile_n(C.byref(d), which)
    if lay == L.WLAYOUT_SHIFT:
        sfx = " f16" if d.operand_fmt == L.FMT_F16 else ""
        if which == 2:
            return "conv_wgrad_shift_kernel<%d, %d>%s" % (bn, passes, sfx)
        return "conv_shift_tcgen05_kernel<%d, %d, %d>%s" % (bn, passes, L.lib().affgw_conv_tc_tile_m(C.byref(d), which) // 128, sfx)
    sfx = " f16" if d.operand_fmt == L.FMT_F16 else ""
    return (("conv_wgrad_tcgen05_kernel<%d, %d>" if which == 2 else "conv_igemm_tcgen05_kernel<%d, %d, float>") % (bn, passes)) + sfx


class _timed:
    def __init__(self, name, flops, tag=None):
        self.on = _profile["records"] is not None
        self.name, self.flops, self.tag = name, flops, tag

    def __enter__(self):
        if self.on:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if self.on:
            self.e1.record()
            _profile["records"].append((self.name, self.tag, self.e0, self.e1, self.flops))


def set_precision(mode):
    """Activations, statistics and gradients are stored in fp32 in every mode; the mode selects the convolution engine:
       'fp32'   : CUDA-core FFMA convolutions (<= 1e-4 against the CPU reference)
       'f16'    : tcgen05 tensor-core convolutions on FP16 operand planes (11 significant bits per plane against bf16's 8, same
                  tensor throughput, fp32 TMEM accumulation).  Forward GEMMs on split operands (hi + remainder plane, three
                  MMAs per product), backward GEMMs one MMA per product.  fp16's range is handled by power-of-two scales the
                  kernels undo exactly: weights x 2^8, every dY tensor x 2^k from its max-abs (device side, no host sync).
                  Image 2.5e-5 and gradient cosine 0.99998 per tensor against fp32 in the operand-rounding model
                  (scripts/precision_sweep.py): this is the mode bench.py measures; the decoder's up-convolutions spend part of
                  that margin on a single forward pass (modules_tro.Decoder).
       'bf16'   : the same on BF16 operand planes (no scales needed: fp32's range).  FORWARD GEMMs three MMAs per product
                  (a_hi*w_hi + a_lo*w_hi + a_hi*w_lo, fp32 TMEM accumulation): forward rounding is amplified layer by layer
                  and only the split meets the 2e-2 image bar.  BACKWARD GEMMs (input and weight gradients) are linear in
                  dY, their rounding is not amplified, and run one MMA per product: gradient cosine stays >= 0.9997 per
                  tensor against the fp32 reference (scripts/precision_sweep.py, DESIGN.md "precision").
       'bf16x3' : three MMAs per product in all three GEMMs (the round-1 'bf16' mode).
       'bf16x1' : one bf16 MMA per product everywhere (fastest; on this network the 2^-9 forward operand rounding is
                  amplified to ~2e-1 on the image at random init)."""
    if mode not in _MODES:
        raise ValueError("precision must be one of " + ", ".join(repr(m) for m in _MODES))
    _state["mode"] = mode
    _state["passes"] = _MODES[mode]


def precision():
    return _state["mode"]


class conv_passes:
    """Context manager: the tensor-core convolutions whose FORWARD is issued inside run with the given MMAs per product
    (1 or 3) in their forward / input-gradient / weight-gradient GEMMs; None keeps the mode's value.  The backward GEMMs of
    a layer use what was in force at its forward.  conv_passes(n) sets all three."""

    def __init__(self, fwd=None, dgrad=None, wgrad=None):
        if dgrad is None and wgrad is None and fwd is not None:
            dgrad = wgrad = fwd
        assert all(v in (None, 1, 3) for v in (fwd, dgrad, wgrad))
        self.req = (fwd, dgrad, wgrad)

    def __enter__(self):
        self.prev = _state["passes"]
        if _state["mode"] != "fp32":
            _state["passes"] = tuple(p if r is None else r for p, r in zip(self.prev, self.req))

    def __exit__(self, *a):
        _state["passes"] = self.prev


class relaxed_forward:
    """Context manager for a generator forward whose image only FEEDS the discriminator (dis_update's no_grad pair,
    network_tro.py:117-118): inside, the layers that declare themselves tolerant (`relaxed()` is true for them) run their forward
    GEMM with one fp16 MMA per product instead of three.  Which layers tolerate it is decided by the discriminator's gradients
    (scripts/precision_sweep.py disfwd, then measured on B200 at the benchmarked shapes): VGG convolutions 6-16 and the decoder's
    ResBlock convolutions leave the worst discriminator tensor at cosine 0.99994, every layer single-pass 0.9994 (the first five
    VGG layers stay at three passes).  Mode 'f16' only."""

    def __init__(self, flag=True, vgg_from=None):
        self.flag = bool(flag)
        self.vgg_from = vgg_from        # `features` index of the first single-pass VGG convolution; None = the encoder's default

    def __enter__(self):
        self.prev = (_state.get("relaxed", False), _state.get("relaxed_from"))
        _state["relaxed"] = self.flag and _state["mode"] == "f16"
        _state["relaxed_from"] = self.vgg_from

    def __exit__(self, *a):
        _state["relaxed"], _state["relaxed_from"] = self.prev


def relaxed():
    return bool(_state.get("relaxed", False))


def relaxed_from(default):
    v = _state.get("relaxed_from")
    return default if v is None else v


class operand_format:
    """Context manager: the position-space convolutions whose forward is issued inside run on FP16 operand planes, one MMA per
    product in all three GEMMs (fmt = "f16"), instead of the mode's bf16 planes.  fp16 carries 11 significant bits per plane
    against bf16's 8, which is what lets the decoder's convolutions - the last layers in front of the image, where forward rounding
    is amplified least - drop from three tensor-core passes to one inside the accuracy bars (scripts/precision_sweep.py:
    image 1.8e-3, worst per-tensor gradient cosine 0.99956 at 50 planes, batch 8).  fp16's range is handled by power-of-two
    scales that the kernels undo exactly: weights x 2^8, dY x a per-tensor 2^k from its max-abs (affgw_amax_scale).
    Only active in mode 'bf16'; layers the position-space kernels do not take keep the mode's bf16 route."""

    def __init__(self, fmt):
        assert fmt in (None, "bf16", "f16")
        self.fmt = None if fmt == "bf16" else fmt

    def __enter__(self):
        self.prev = _state["fmt"]
        _state["fmt"] = self.fmt

    def __exit__(self, *a):
        _state["fmt"] = self.prev


class accumulate_into_grad:
    """Context manager (opt-in, used by trainer.Trainer around forward + backward): a tensor-core weight / bias gradient whose
    parameter already HAS a `.grad` is added into that buffer by the weight-gradient kernel's own reduction (it accumulates
    anyway) and `None` is returned to autograd, instead of materialising a fresh gradient that autograd then adds with an ATen
    kernel (one extra launch and three tensor passes per shared parameter: the two decodes of gen_update, the two backward
    calls of dis_update).  Same numbers, same `.grad` afterwards.  NOT for `torch.autograd.grad(...)` users: the functional
    API expects returned gradients - hence opt-in."""

    def __init__(self, flag=True):
        self.flag = bool(flag)

    def __enter__(self):
        self.prev = _state["grad_accum"]
        _state["grad_accum"] = self.flag

    def __exit__(self, *a):
        _state["grad_accum"] = self.prev


class wgrad_side_stream:
    """Context manager (opt-in, used by trainer.Trainer around forward + backward): the tensor-core WEIGHT-gradient GEMMs of the
    backward passes issued inside are launched on a second CUDA stream, forked from the launching stream after the layer's dY
    operand planes are written and joined when the context exits.  Nothing in a backward pass reads a weight gradient, so this
    takes ~15 ms of tensor-bound kernels off the dependency chain dY -> dgrad -> norm backward -> next layer's dY: they run
    beside the HBM-bound kernels of the following layers and fill the SMs a kernel's last wave leaves idle.  Inside a CUDA-graph
    capture the fork / join become graph edges (two branches of one graph).

    Each weight gradient is written straight into the parameter's `.grad` (created here when it is None; accumulated inside
    the kernel's own reduction when it exists) and `None` is returned to autograd, so no autograd kernel ever touches a buffer
    the side stream is still writing.  The operand planes the side stream reads are kept alive until the join (they are
    allocated on, and returned to, the launching stream).  Same numbers, same `.grad` afterwards - once the context has exited;
    not for `torch.autograd.grad(...)`.  Implies accumulate_into_grad."""

    _streams = {}

    def __init__(self, flag=True):
        self.flag = bool(flag)

    def __enter__(self):
        self.prev = (_state["wgrad_side"], _state["grad_accum"])
        if self.flag:
            dev = torch.cuda.current_device()
            st = wgrad_side_stream._streams.get(dev)
            if st is None:
                st = wgrad_side_stream._streams[dev] = torch.cuda.Stream(device=dev, priority=_WGRAD_PRIO)
            _state["wgrad_side"] = {"stream": st, "keep": [], "forked": False}
            _state["grad_accum"] = True
        return self

    def join(self):
        side = _state["wgrad_side"]
        if side is not None and side["forked"]:
            torch.cuda.current_stream().wait_stream(side["stream"])
            side["forked"] = False
        if side is not None:
            side["keep"].clear()

    def __exit__(self, *a):
        if self.flag:
            self.join()
        _state["wgrad_side"], _state["grad_accum"] = self.prev


class bn_updates_twice:
    """Context manager: the running-statistics updates of the train-mode BatchNorm forwards issued inside are applied a second
    time, in the same order, when the context exits - what a second, identical run of the same forward would leave behind (the
    exponential average is order-dependent when one layer is called with several inputs: text encoder and iAFF see the
    label and the swapped label, modules_tro.py:230-259).  Used when one generator forward stands for the two the reference
    runs per iteration (Trainer(share_generator_forward=True))."""

    def __enter__(self):
        self.prev = _state["bn_record"]
        _state["bn_record"] = []
        return self

    def __exit__(self, exc_type, *a):
        rec, _state["bn_record"] = _state["bn_record"], self.prev
        if exc_type is None:
            for rm, rv, nbt, mean, var, cc, mom in rec:
                L.call("affgw_bn_update_running", rm.data_ptr(), rv.data_ptr(), L.ptr(nbt), mean.data_ptr(), var.data_ptr(), cc,
                       mom, L.stream())


class scratch_scope:
    """Context manager: the small self-resetting device scratch words of the launches issued inside (the running maximum of
    ops._amax_scale) are private to `tag`.  Two CUDA graphs captured on the same stream but replayed CONCURRENTLY (Trainer's
    cla_update beside dis_update) must not share them."""

    def __init__(self, tag):
        self.tag = tag

    def __enter__(self):
        self.prev = _state["scratch_tag"]
        _state["scratch_tag"] = self.tag

    def __exit__(self, *a):
        _state["scratch_tag"] = self.prev


class _on_wgrad_stream:
    """Fork the side stream from the launching stream and make it current; `keep` = tensors it reads (alive until the join)."""

    def __init__(self, side, keep):
        self.side = side
        if side is not None:
            side["keep"].append(keep)

    def __enter__(self):
        if self.side is not None:
            self.side["stream"].wait_stream(torch.cuda.current_stream())
            self.side["forked"] = True
            self.ctx = torch.cuda.stream(self.side["stream"])
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.side is not None:
            self.ctx.__exit__(*a)


def _grad_slot(p):
    """The parameter's existing gradient buffer when it can take an in-place accumulation, else None."""
    if not _state["grad_accum"] or p is None or not p.is_leaf:
        return None
    g = p.grad
    if g is None or g.dtype != torch.float32 or g.shape != p.shape or not g.is_contiguous() or g.device != p.device:
        return None
    return g


class wgrad_passes(conv_passes):
    """conv_passes(wgrad=n): used around the discriminator / classifier trunks, whose weight gradients sum many
    cancelling terms (mean-reduced losses over all logits) and keep the split operands."""

    def __init__(self, n):
        super().__init__(None, None, n)


def act_dtype():
    return torch.float32


def use_thin_kernels(flag=True):
    """Route single-channel-sided stencils (7x7 stems, decoder output conv) to the fp32 CUDA-core kernels (default) or,
    when off, through the tensor-core kernels like every other convolution (tests)."""
    global _THIN
    prev, _THIN = _THIN, bool(flag)
    return prev


def force_simt(flag=True):
    """Route every convolution through the CUDA-core kernel (used by tests to cross-check tcgen05)."""
    _state["force_simt"] = bool(flag)


def force_simt_wgrad(flag=True):
    """Keep forward / dgrad on tcgen05 but compute weight gradients on CUDA cores (cross-check of the MN-major path)."""
    _state["simt_wgrad"] = bool(flag)


def _nb(t):
    return 0 if t is None else t.numel() * t.element_size()


def _err_tensor(device):
    key = (device.type, device.index)
    if key not in _err_flag:
        _err_flag[key] = torch.zeros(1, dtype=torch.int32, device=device)
    return _err_flag[key]


def check_device_errors(device=None):
    """Raise if any kernel flagged an out-of-range label / writer id since the last check (one host sync)."""
    for key, t in _err_flag.items():
        if int(t.item()) != 0:
            t.zero_()
            raise RuntimeError("libaffgw: out-of-range token or class index")


# ------------------------------------------------------------------------------------------------ layout helpers
def _pitch4(x):
    """(N, C, H, W, pitch) if the 4-D tensor is addressable as NHWC with a channel pitch, else None."""
    n, c, h, w = x.shape
    sn, sc, sh, sw = x.stride()
    if w > 1:
        pitch = sw
    elif h > 1:
        pitch = sh
    elif n > 1:
        pitch = sn
    else:
        pitch = c
    ok = (c == 1 or sc == 1) and pitch >= c and (w == 1 or sw == pitch) and (h == 1 or sh == w * pitch) and \
         (n == 1 or sn == h * w * pitch)
    return (n, c, h, w, pitch) if ok else None


def empty_cl(n, c, h, w, dtype, device, zero=False):
    """NCHW-shaped tensor stored NHWC."""
    base = (torch.zeros if zero else torch.empty)((n, h, w, c), dtype=dtype, device=device)
    return base.permute(0, 3, 1, 2)


def to_internal(x, c_pad=None, dtype=None):
    """Bring a caller tensor (typically NCHW fp32, network_tro.py:30-36) to the internal layout / dtype.
    c_pad zero-pads the channel dimension (the 50 -> 64 style-image planes of the first VGG conv)."""
    L.require_cuda(x)
    dtype = dtype or act_dtype()
    if x.dim() != 4:
        if x.dtype == dtype and x.is_contiguous():
            return x
        return _cast(x.contiguous(), dtype)
    n, c, h, w = x.shape
    cp = c_pad or c
    g = _pitch4(x)
    if g is not None and g[4] == c and cp == c:
        return x if x.dtype == dtype else _cast(x, dtype)
    if x.dtype != torch.float32 or not x.is_contiguous():
        x = x.float().contiguous()
    out = empty_cl(n, cp, h, w, dtype, x.device)
    L.call("affgw_nchw_to_nhwc", x.data_ptr(), out.data_ptr(), L.dt(out), n, c, h * w, cp, L.stream())
    return out


def _cast(x, dtype):
    out = torch.empty_like(x, dtype=dtype)
    L.call("affgw_cast", x.data_ptr(), L.dt(x), out.data_ptr(), L.dt(out), x.numel(), L.stream())
    return out


def _dense_cl(t, dtype=None):
    """Dense NHWC tensor (pitch == C) of the requested dtype; used on incoming gradients."""
    if t.dim() == 4:
        g = _pitch4(t)
        if g is None or g[4] != g[1]:
            t = to_internal(t if t.dtype == torch.float32 else t.float(), dtype=dtype or t.dtype)
    elif not t.is_contiguous():
        t = t.contiguous()
    if dtype is not None and t.dtype != dtype:
        t = _cast(t, dtype)
    return t


class _ToInternal(Function):
    """Boundary conversion for inputs that require grad (dis_update real images, network_tro.py:108-109)."""

    @staticmethod
    def forward(ctx, x, c_pad, dtype):
        ctx.shape = x.shape
        return to_internal(x.detach(), c_pad, dtype)

    @staticmethod
    def backward(ctx, g):
        if len(ctx.shape) != 4:
            g = g.contiguous()
            return (g if g.dtype == torch.float32 else _cast(g, torch.float32)), None, None
        n, c, h, w = ctx.shape
        g = _dense_cl(g)
        out = torch.empty((n, c, h, w), dtype=torch.float32, device=g.device)
        L.call("affgw_nhwc_to_nchw", g.data_ptr(), out.data_ptr(), L.dt(g), n, c, h * w, g.shape[1], L.stream())
        return out, None, None


def _is_internal(x, c_pad, dtype):
    if x.dtype != dtype or not x.is_cuda:
        return False
    if x.dim() != 4:
        return x.dim() == 0 or x.stride(-1) == 1
    return _pitch4(x) is not None and (c_pad is None or c_pad == x.shape[1])


def input_to_internal(x, c_pad=None, dtype=None):
    """Identity for tensors our own operators produced; converts caller tensors at the boundary."""
    dtype = dtype or act_dtype()
    if _is_internal(x, c_pad, dtype):
        return x
    if x.requires_grad and torch.is_grad_enabled():
        return _ToInternal.apply(x, c_pad, dtype)
    return to_internal(x, c_pad, dtype)


# ------------------------------------------------------------------------------------------------ convolution
ConvCfg = collections.namedtuple("ConvCfg", "stride pad pad_mode upsample pre_act post_act out_dtype stride_w")


class _WeightCache:
    """Packed operand copies of a parameter, rebuilt when the parameter's version counter moves.  An entry is a mutable list
    [version key, parameter address, packed tensor, builder, parameter]; `builder(out)` re-packs in place when given the old
    tensor, so a CUDA graph that read the packed copy keeps seeing the current weights (refresh_packed)."""

    @staticmethod
    def version(weight):
        # autograd's version counter catches ordinary in-place updates; fused / multi-tensor optimiser kernels do not always
        # bump it (torch.optim.Adam(fused=True) does not), so an explicit per-parameter epoch (weights_updated) is part of
        # the key as well
        return (weight._version, weight.__dict__.get("_affgw_epoch", 0))

    @staticmethod
    def get(weight, kind, builder):
        cache = weight.__dict__.setdefault("_affgw_packed", {})
        ent = cache.get(kind)
        ver = _WeightCache.version(weight)
        if ent is None or ent[1] != weight.data_ptr():
            ent = cache[kind] = [ver, weight.data_ptr(), builder(None), builder, weight]
        elif ent[0] != ver:
            ent[2] = builder(ent[2])            # same buffer, new contents
            ent[0] = ver
        return ent[2]


def weights_updated(module_or_params):
    """Tell the packed-operand cache that these parameters were modified in place by something autograd's version counter
    does not see (a fused optimiser step, a raw-pointer copy).  Cheap: one integer per parameter."""
    params = module_or_params.parameters() if hasattr(module_or_params, "parameters") else module_or_params
    for p in params:
        p.__dict__["_affgw_epoch"] = p.__dict__.get("_affgw_epoch", 0) + 1


def clear_weight_cache(module):
    """Drop every packed operand copy of `module`'s parameters.  The graphed generators call this before capturing so that
    their graph re-packs the weights it reads on every replay instead of pointing at a copy an earlier eager call made."""
    for p in module.parameters():
        p.__dict__.pop("_affgw_packed", None)


def packed_entries(module_or_params):
    """Every packed operand copy currently cached for these parameters (a list of cache entries to hand to refresh_packed).
    trainer.Trainer collects them when it captures its CUDA graphs: the graphs read these buffers and contain no packing
    kernels; the packing runs once per optimiser step, next to the step (off the critical path when the step is overlapped)."""
    params = module_or_params.parameters() if hasattr(module_or_params, "parameters") else module_or_params
    return [ent for p in params for ent in p.__dict__.get("_affgw_packed", {}).values()]


def refresh_packed(entries):
    """Re-pack, in place and on the current stream, every entry whose parameter has changed since it was packed; returns the
    number of packing kernels launched."""
    n = 0
    for ent in entries:
        w = ent[4]
        ver = _WeightCache.version(w)
        if ent[0] != ver and ent[1] == w.data_ptr():
            ent[2] = ent[3](ent[2])
            ent[0] = ver
            n += 1
    return n


def _w4(weight):
    return weight if weight.dim() == 4 else weight.view(weight.shape[0], weight.shape[1], 1, 1)


def _pack(weight, dtype, ipad, flip):
    w4 = _w4(weight.detach())
    co, ci, kh, kw = w4.shape

    def build(out):
        rows = ci if flip else co
        if out is None:
            out = torch.empty((rows, kh, kw, ipad), dtype=dtype, device=weight.device)
        L.call("affgw_pack_weight", w4.contiguous().data_ptr(), out.data_ptr(), L.dt(out), co, ci, kh, kw, ipad,
               int(flip), L.stream())
        return out
    return _WeightCache.get(weight, ("plain", dtype, ipad, flip), build)


def _pack_tc(weight, ipad, flip, passes, layout, fmt=0):
    w4 = _w4(weight.detach())
    co, ci, kh, kw = w4.shape

    def build(out):
        if out is None:
            nbytes = L.lib().affgw_pack_weight_tc_bytes(co, ci, kh, kw, ipad, int(flip), passes, layout)
            if nbytes <= 0:
                raise RuntimeError("affgw_pack_weight_tc_bytes: bad configuration")
            out = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=weight.device)  # 16-bit storage (bf16 or fp16 bits)
        L.call("affgw_pack_weight_tc_fmt", w4.contiguous().data_ptr(), out.data_ptr(), co, ci, kh, kw, ipad, int(flip), passes,
               layout, fmt, L.stream())
        return out
    return _WeightCache.get(weight, ("tc", ipad, flip, passes, layout, fmt), build)


def prefer_shift_kernel(flag=True):
    """Route stride-1 convolutions through the shared-memory-window tcgen05 kernel (default) or the im2col one."""
    return bool(L.lib().affgw_conv_tc_prefer_shift(int(bool(flag))))


def _up8(c):
    return (c + 7) // 8 * 8


def _split_planes(x, rows, c, pitch, passes, pre_act="none", fmt=0, scale=None):
    """fp32 activations [rows][pitch] -> 16-bit operand planes [1 or 2][rows][c_store] of the tcgen05 kernels (bf16, or fp16
    of x * scale[0])."""
    cs = _up8(c)
    planes = torch.empty((2 if passes == 3 else 1, rows, cs), dtype=torch.bfloat16, device=x.device)
    L.call("affgw_split_planes_fmt", x.data_ptr(), L.dt(x), planes.data_ptr(), rows, c, pitch, cs, passes, L.ACT[pre_act],
           fmt, L.ptr(scale), L.stream(), nbytes=rows * c * x.element_size() + _nb(planes))
    return planes


def _pos_frames(d):
    fx, fy = L.PosFrame(), L.PosFrame()
    L.call("affgw_conv_pos_frames", C.byref(d), C.byref(fx), C.byref(fy))
    return fx, fy


def _split_positions(src, frame, hs, ws, c, pitch, up, origin, pad_mode, pre_act, passes, colsum=None, fmt=0, scale=None):
    """fp32 NHWC tensor -> planar position planes [1 or 2][G][QA][8] (bf16, or fp16 of src * scale[0]) on the frame of a
    stride-1 convolution."""
    nbytes = L.lib().affgw_position_planes_bytes(C.byref(frame), passes)
    if nbytes <= 0:
        raise RuntimeError("affgw_position_planes_bytes: bad frame")
    planes = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=src.device)
    L.call("affgw_split_positions_fmt", src.data_ptr(), L.dt(src), planes.data_ptr(), C.byref(frame), hs, ws, c, pitch, up, origin,
           origin, L.PAD[pad_mode], L.ACT[pre_act], passes, L.ptr(colsum), fmt, L.ptr(scale), L.stream(),
           nbytes=frame.N * hs * ws * c * src.element_size() + _nb(planes))
    return planes


def _amax_scale(t):
    """Device-side per-tensor power-of-two scale of an fp16 operand: -> float32 [2] = (2^k, 2^-k), no host synchronisation."""
    out = torch.empty(2, dtype=torch.float32, device=t.device)
    key = (t.device.index, torch.cuda.current_stream().cuda_stream, _state["scratch_tag"])
    ws = _amax_ws.get(key)
    if ws is None:                      # 8 bytes of scratch per (device, stream): zero on entry, left zero by the kernel
        ws = _amax_ws[key] = torch.zeros(2, dtype=torch.int32, device=t.device)
    L.call("affgw_amax_scale", t.data_ptr(), t.numel(), out.data_ptr(), ws.data_ptr(), L.stream(), nbytes=_nb(t))
    return out


_amax_ws = {}


def _conv_geom(x, weight, cfg):
    """-> dict(N,H,W,Cx,pitch,Cout,Cin,KH,KW,Ho,Wo, two_d)"""
    w4 = _w4(weight)
    co, ci, kh, kw = w4.shape
    if x.dim() == 2:
        n, cx = x.shape
        h = w = 1
        pitch = x.stride(0) if n > 1 else max(cx, x.stride(0))
        if x.stride(1) != 1:
            raise RuntimeError("conv2d: 2-D input must have unit inner stride")
    else:
        g = _pitch4(x)
        if g is None:
            raise RuntimeError("conv2d: input is not channels-last addressable (call ops.to_internal first)")
        n, cx, h, w, pitch = g
    if cx < ci:
        raise RuntimeError(f"conv2d: input has {cx} channels, weight expects {ci}")
    hv, wv = h * cfg.upsample, w * cfg.upsample
    ho = (hv + 2 * cfg.pad - kh) // cfg.stride + 1
    wo = (wv + 2 * cfg.pad - kw) // cfg.stride_w + 1
    return dict(N=n, H=h, W=w, Cx=cx, pitch=pitch, Cout=co, Cin=ci, KH=kh, KW=kw, Ho=ho, Wo=wo, two_d=x.dim() == 2)


def _desc(g, cfg, cin, x_dt, w_dt, y_dt, algo, in_pitch=None, out_pitch=None, passes=1, grad_dt=0, pre_act=None, fmt=0):
    d = L.ConvDesc()
    d.N, d.H, d.W, d.Cin = g["N"], g["H"], g["W"], cin
    d.Cout, d.KH, d.KW = g["Cout"], g["KH"], g["KW"]
    d.stride, d.pad, d.pad_mode, d.upsample = cfg.stride, cfg.pad, L.PAD[cfg.pad_mode], cfg.upsample
    d.Ho, d.Wo = g["Ho"], g["Wo"]
    d.in_pitch, d.out_pitch = in_pitch or g["pitch"], out_pitch or g["Cout"]
    d.pre_act, d.post_act = L.ACT[cfg.pre_act if pre_act is None else pre_act], L.ACT[cfg.post_act]
    d.x_dtype, d.w_dtype, d.y_dtype = x_dt, w_dt, y_dt
    d.algo, d.passes, d.grad_dtype = algo, passes, grad_dt
    d.stride_w = cfg.stride_w
    d.operand_fmt = fmt
    return d


class _Conv2d(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, addend, cfg):
        L.require_cuda(x, weight)
        if x.dtype != torch.float32:
            raise RuntimeError("conv2d: activations are fp32 tensors (operand rounding happens inside the convolution)")
        g = _conv_geom(x, weight, cfg)
        y_dtype = torch.float32
        use_tc = _state["mode"] != "fp32" and not _state["force_simt"]
        passes, pd, pw = _state["passes"]
        need_wgrad = weight.requires_grad and torch.is_grad_enabled()
        # the x operand planes are shared by the forward and the weight-gradient GEMM: written with a remainder plane when
        # either needs it (a single-pass kernel reads the leading plane of a two-plane tensor)
        px = 3 if (passes == 3 or (need_wgrad and pw == 3)) else 1
        if g["two_d"]:
            y = torch.empty((g["N"], g["Cout"]), dtype=y_dtype, device=x.device)
        else:
            y = empty_cl(g["N"], g["Cout"], g["Ho"], g["Wo"], y_dtype, x.device)
        if addend is not None:
            addend = _dense_cl(addend, y_dtype)
        b32 = None if bias is None else bias.detach()
        flops = 2.0 * g["N"] * g["Ho"] * g["Wo"] * g["Cout"] * g["Cin"] * g["KH"] * g["KW"]
        tag = "%dx%dx%d c%d->%d k%d s%d u%d" % (g["N"], g["H"], g["W"], g["Cin"], g["Cout"], g["KH"], cfg.stride if cfg.stride == cfg.stride_w else cfg.stride * 10 + cfg.stride_w, cfg.upsample)
        planes = None
        thin = False
        fmt = L.FMT_BF16
        if use_tc and addend is None and cfg.pre_act == "none" and not g["two_d"] and g["pitch"] == g["Cin"] \
                and (g["Cin"] == 1 or g["Cout"] == 1) and _THIN:
            dthin = _desc(g, cfg, g["Cin"], L.F32, L.F32, L.F32, L.ALGO_SIMT)
            thin = bool(L.lib().affgw_conv_thin_supported(C.byref(dthin)))
        if thin:
            # single-channel-sided stencil (7x7 stems, decoder output conv): fp32 CUDA-core kernels, no operand planes
            ws = torch.empty(L.lib().affgw_conv_thin_ws_bytes(C.byref(dthin), 0), dtype=torch.uint8, device=x.device)
            with _timed("conv_fwd_thin", flops, (tag, "conv_thin fwd") if _profile["records"] is not None else tag):
                L.call("affgw_conv_thin_fwd", x.data_ptr(), _w4(weight.detach()).contiguous().data_ptr(), L.ptr(b32),
                       y.data_ptr(), ws.data_ptr(), C.byref(dthin), L.stream())
            layout = 0
        elif use_tc:
            cs = _up8(g["Cin"])
            d = _desc(g, cfg, g["Cin"], L.BF16, L.BF16, L.F32, L.ALGO_TC, in_pitch=cs, passes=passes, pre_act="none")
            layout = L.lib().affgw_conv_tc_layout(C.byref(d), 0)
            if not layout:
                raise RuntimeError("conv2d: tcgen05 kernels refused the shape: " + L.last_error())
            if _state["mode"] == "f16":
                fmt = L.FMT_F16
            elif _state["fmt"] == "f16" and _state["mode"] == "bf16" and layout == L.WLAYOUT_SHIFT:
                # ops.operand_format inside the bf16 mode: fp16 planes for this layer, single pass in all three GEMMs
                fmt = L.FMT_F16
                passes = pd = pw = px = 1
            if fmt:
                d = _desc(g, cfg, g["Cin"], L.BF16, L.BF16, L.F32, L.ALGO_TC, in_pitch=cs, passes=passes, pre_act="none", fmt=fmt)
            if layout == L.WLAYOUT_SHIFT:
                fx, _ = _pos_frames(d)
                planes = _split_positions(x, fx, g["H"], g["W"], g["Cin"], g["pitch"], cfg.upsample, cfg.pad, cfg.pad_mode,
                                          cfg.pre_act, px, fmt=fmt)
            else:
                planes = _split_planes(x, g["N"] * g["H"] * g["W"], g["Cin"], g["pitch"], px, cfg.pre_act, fmt=fmt)
            wp = _pack_tc(weight, cs, False, passes, layout, fmt)
            with _timed("conv_fwd_tcgen05", flops, (tag, _kernel_name(d, 0, passes)) if _profile["records"] is not None else tag):
                L.call("affgw_conv2d_fwd", planes.data_ptr(), wp.data_ptr(), L.ptr(b32), L.ptr(addend), y.data_ptr(),
                       C.byref(d), L.stream())
        else:
            d = _desc(g, cfg, g["Cin"], L.F32, L.F32, L.F32, L.ALGO_SIMT)
            wp = _pack(weight, torch.float32, g["Cin"], False)
            with _timed("conv_fwd_simt", flops, tag):
                L.call("affgw_conv2d_fwd", x.data_ptr(), wp.data_ptr(), L.ptr(b32), L.ptr(addend), y.data_ptr(), C.byref(d),
                       L.stream())
        ctx.cfg, ctx.g, ctx.use_tc, ctx.passes, ctx.tag = cfg, g, use_tc and not thin, (pd, pw, px), tag
        ctx.fmt = fmt
        ctx.layout = layout if use_tc else 0
        ctx.thin = thin
        ctx.has_bias, ctx.has_addend = bias is not None, addend is not None
        ctx.bias_ref = bias
        keep_x = (not use_tc) or thin or cfg.pre_act != "none" or _state["simt_wgrad"]
        ctx.save_for_backward(x if keep_x else None, weight, y if cfg.post_act != "none" else None, planes)
        ctx.x_meta = (x.shape, x.device)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y, planes = ctx.saved_tensors
        cfg, g, (pd, pw, px) = ctx.cfg, ctx.g, ctx.passes
        need_x, need_w, need_b, need_a = ctx.needs_input_grad[:4]
        if px != 3:
            pw = 1                      # the saved x planes have no remainder plane
        # dY planes feed both backward GEMMs: remainder plane when either runs split operands
        pdy = 3 if ((need_x and pd == 3) or (need_w and pw == 3)) else 1
        fmt = ctx.fmt
        dy_scale = None
        dev = ctx.x_meta[1]
        dz = _dense_cl(dy, torch.float32)
        if cfg.post_act != "none":
            t = torch.empty_like(dz)
            L.call("affgw_act_bwd", dz.data_ptr(), y.data_ptr(), t.data_ptr(), L.dt(dz), dz.numel(), L.ACT[cfg.post_act],
                   L.stream(), nbytes=3 * _nb(dz))
            dz = t
        st = L.stream()
        M = g["N"] * g["Ho"] * g["Wo"]
        cin, cout = g["Cin"], g["Cout"]
        db = dw = dx = None
        use_tc = ctx.use_tc
        # position-space layers: the bias gradient is summed by the kernel that splits dY into operand planes
        fuse_db = use_tc and ctx.layout == L.WLAYOUT_SHIFT and (need_w or need_x) and _FUSE_DB
        db_ret = True
        if ctx.has_bias and need_b:
            bias_p = ctx.bias_ref
            db = _grad_slot(bias_p) if (use_tc and not ctx.thin) else None      # atomics add into the existing .grad
            db_ret = db is None
            if db is None:
                db = torch.zeros(cout, dtype=torch.float32, device=dev)
            if not fuse_db:
                L.call("affgw_colsum", dz.data_ptr(), L.F32, db.data_ptr(), M, cout, cout, st)
        fwd_cfg = cfg._replace(post_act="none")
        flops = 2.0 * M * cout * cin * g["KH"] * g["KW"]
        if use_tc and (need_w or need_x):
            cs, cso = _up8(cin), _up8(cout)
            if ctx.layout == L.WLAYOUT_SHIFT:
                # dY on the forward convolution's position frame: one split feeds both dgrad and wgrad
                d0 = _desc(g, fwd_cfg, cin, L.BF16, L.BF16, L.BF16, L.ALGO_TC, in_pitch=cs, out_pitch=cso, passes=pdy)
                _, fy = _pos_frames(d0)
                if fmt == L.FMT_F16:
                    dy_scale = _amax_scale(dz)       # (2^k, 2^-k) in device memory: gradients need a per-tensor scale in fp16
                dzp = _split_positions(dz, fy, g["Ho"], g["Wo"], cout, cout, 1, 0, "zero", "none", pdy,
                                       colsum=db if fuse_db else None, fmt=fmt, scale=dy_scale)
            else:
                if fmt == L.FMT_F16:
                    dy_scale = _amax_scale(dz)
                dzp = _split_planes(dz, M, cout, cout, pdy, fmt=fmt, scale=dy_scale)
        if ctx.thin:
            dthin = _desc(g, fwd_cfg, cin, L.F32, L.F32, L.F32, L.ALGO_SIMT)
            if need_w:
                dw = torch.zeros(weight.shape, dtype=torch.float32, device=dev)
                with _timed("conv_wgrad_thin", flops, (ctx.tag, "conv_thin wgrad") if _profile["records"] is not None else ctx.tag):
                    L.call("affgw_conv_thin_wgrad", x.data_ptr(), dz.data_ptr(), dw.data_ptr(), C.byref(dthin), st)
            if need_x:
                dx = empty_cl(g["N"], cin, g["H"], g["W"], torch.float32, dev)
                ws = torch.empty(L.lib().affgw_conv_thin_ws_bytes(C.byref(dthin), 1), dtype=torch.uint8, device=dev)
                with _timed("conv_dgrad_thin", flops, (ctx.tag, "conv_thin dgrad") if _profile["records"] is not None else ctx.tag):
                    L.call("affgw_conv_thin_dgrad", dz.data_ptr(), _w4(weight.detach()).contiguous().data_ptr(), dx.data_ptr(),
                           ws.data_ptr(), C.byref(dthin), st)
            da = dz if (ctx.has_addend and need_a) else None
            return dx, dw, db, da, None
        dw_ret = True
        deferred = None
        if need_w:
            tc_w = use_tc and not _state["simt_wgrad"]
            dw = _grad_slot(weight) if tc_w else None   # the unpack kernel accumulates: dw += partials
            dw_ret = dw is None
            side = _state["wgrad_side"] if tc_w else None
            if dw is None:
                dw = torch.zeros(weight.shape, dtype=torch.float32, device=dev)
                if side is not None and weight.is_leaf and weight.grad is None:
                    weight.grad = dw            # ops.wgrad_side_stream: autograd never handles a buffer the side stream writes
                    dw_ret = False
                else:
                    side = None
            if tc_w:
                d = _desc(g, fwd_cfg, cin, L.BF16, L.BF16, L.BF16, L.ALGO_TC, in_pitch=cs, out_pitch=cso, passes=pw, fmt=fmt)
                ws_bytes = L.lib().affgw_conv2d_wgrad_ws_bytes(C.byref(d))
                if ws_bytes <= 0:
                    raise RuntimeError("conv2d_wgrad: tcgen05 kernel refused the shape: " + L.last_error())
                wsb = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                dwg = d

                def launch_wgrad():
                    with _on_wgrad_stream(side, (planes, dzp, dy_scale, wsb)):
                        with _timed("conv_wgrad_tcgen05", flops,
                                    (ctx.tag, _kernel_name(dwg, 2, pw)) if _profile["records"] is not None else ctx.tag):
                            L.call("affgw_conv2d_wgrad_scaled", planes.data_ptr(), dzp.data_ptr(), dw.data_ptr(), wsb.data_ptr(),
                                   C.byref(dwg), None if dy_scale is None else dy_scale.data_ptr() + 4, L.stream())
                if side is not None and need_x and _WGRAD_LATE:
                    deferred = launch_wgrad     # fork AFTER this layer's dgrad: the side stream starts when the dgrad CTAs drain
                else:
                    launch_wgrad()
            else:
                if x is None:
                    raise RuntimeError("conv2d backward: the CUDA-core wgrad needs the saved fp32 input")
                d = _desc(g, fwd_cfg, cin, L.F32, L.F32, L.F32, L.ALGO_SIMT)
                with _timed("conv_wgrad_simt", flops, ctx.tag):
                    L.call("affgw_conv2d_wgrad", x.data_ptr(), dz.data_ptr(), dw.data_ptr(), None, C.byref(d), st)
        if need_x:
            if use_tc:
                if g["two_d"]:
                    dx = torch.empty((g["N"], cin), dtype=torch.float32, device=dev)
                else:
                    dx = empty_cl(g["N"], cin, g["H"], g["W"], torch.float32, dev)
                d = _desc(g, fwd_cfg, cin, L.BF16, L.BF16, L.BF16, L.ALGO_TC, in_pitch=cs, out_pitch=cso, passes=pd,
                          grad_dt=L.F32, fmt=fmt)
                layout = L.lib().affgw_conv_tc_layout(C.byref(d), 1)
                if not layout:
                    raise RuntimeError("conv2d_dgrad: tcgen05 kernels refused the shape: " + L.last_error())
                wt = _pack_tc(weight, cso, True, pd, layout, fmt)
                base = dx
            else:
                if g["Cx"] != cin and not (cfg.pad_mode == "zero" and cfg.upsample == 1 and cfg.pre_act == "none"):
                    raise RuntimeError("conv2d backward: channel-padded input needs a zero-pad, stride-1 convolution")
                dense = g["pitch"] == cin
                if g["two_d"]:
                    base = torch.empty((g["N"], cin), dtype=torch.float32, device=dev) if dense else \
                        torch.zeros((g["N"], g["pitch"]), dtype=torch.float32, device=dev)
                    dx = base if dense else base[:, :g["Cx"]]
                else:
                    base = empty_cl(g["N"], g["pitch"], g["H"], g["W"], torch.float32, dev, zero=not dense)
                    dx = base if dense else base[:, :g["Cx"]]
                d = _desc(g, fwd_cfg, cin, L.F32, L.F32, L.F32, L.ALGO_SIMT)
                wt = _pack(weight, torch.float32, cout, True)
            ws_bytes = L.lib().affgw_conv2d_dgrad_ws_bytes(C.byref(d))
            if ws_bytes < 0:
                raise RuntimeError("conv2d_dgrad_ws_bytes: " + L.last_error())
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes else None
            src = dzp if use_tc else dz
            with _timed("conv_dgrad_tcgen05" if use_tc else "conv_dgrad_simt", flops,
                        (ctx.tag, _kernel_name(d, 1, pd)) if (use_tc and _profile["records"] is not None) else ctx.tag):
                L.call("affgw_conv2d_dgrad_scaled", src.data_ptr(), wt.data_ptr(), L.ptr(x), base.data_ptr(), L.ptr(ws),
                       C.byref(d), None if dy_scale is None else dy_scale.data_ptr() + 4, st)
            if use_tc and g["Cx"] != cin:      # the weight reads only the first `cin` channels of a wider input
                full = torch.zeros(ctx.x_meta[0], dtype=torch.float32, device=dev).contiguous(memory_format=torch.channels_last) \
                    if len(ctx.x_meta[0]) == 4 else torch.zeros(ctx.x_meta[0], dtype=torch.float32, device=dev)
                full[:, :cin] = dx
                dx = full
        if deferred is not None:
            deferred()
        da = dz if (ctx.has_addend and need_a) else None
        return dx, (dw if dw_ret else None), (db if db_ret else None), da, None


def conv2d(x, weight, bias=None, stride=1, pad=0, pad_mode="zero", upsample=1, pre_act="none", post_act="none",
           addend=None, out_dtype=None):
    """pad -> [nearest x2] -> conv -> +bias -> +addend -> activation   (blocks.py:150-163, modules_tro.py:594-598)"""
    sh, sw = (stride if isinstance(stride, (tuple, list)) else (stride, stride))     # (rows, columns), nn.Conv2d order
    cfg = ConvCfg(int(sh), int(pad), pad_mode, int(upsample), pre_act, post_act, out_dtype, int(sw))
    return _Conv2d.apply(x, weight, bias, addend, cfg)


def linear(x, weight, bias=None, post_act="none", out_dtype=None):
    """nn.Linear as a 1x1 convolution over the leading dimensions (modules_tro.py:222,252-259,272-282)."""
    lead = x.shape[:-1]
    x2 = x if x.dim() == 2 else x.reshape(-1, x.shape[-1])
    y = conv2d(x2, weight, bias, post_act=post_act, out_dtype=out_dtype)
    return y if x.dim() == 2 else y.view(*lead, y.shape[-1])


# ------------------------------------------------------------------------------------------------ normalisation
def _stats(x, G, P, Cc, eps, unbiased, want_var=False):
    dev = x.device
    ws = torch.empty(2 * G * Cc, dtype=torch.float32, device=dev)
    mean = torch.empty(G * Cc, dtype=torch.float32, device=dev)
    rstd = torch.empty(G * Cc, dtype=torch.float32, device=dev)
    var = torch.empty(G * Cc, dtype=torch.float32, device=dev) if want_var else None
    L.call("affgw_norm_stats", x.data_ptr(), L.dt(x), ws.data_ptr(), mean.data_ptr(), rstd.data_ptr(), L.ptr(var), G, P, Cc,
           float(eps), int(unbiased), L.stream(), nbytes=_nb(x))
    return mean, rstd, var


def _gpc(x, per_sample):
    if x.dim() == 4:
        n, c, h, w = x.shape
        return (n, h * w, c) if per_sample else (1, n * h * w, c)
    n, c = x.shape
    return (n, 1, c) if per_sample else (1, n, c)


class _InstanceNorm(Function):
    """Per-(n, c) normalisation with optional per-(n, c) affine, activation and residual:
    nn.InstanceNorm2d (+ReLU), AdaptiveInstanceNorm2d's F.batch_norm trick (blocks.py:197-204), mean_variance_norm."""

    @staticmethod
    def forward(ctx, x, gamma, beta, residual, act, eps, unbiased):
        x = _dense_cl(x)
        G, P, Cc = _gpc(x, True)
        mean, rstd, _ = _stats(x, G, P, Cc, eps, unbiased)
        if gamma is not None:
            gamma, beta = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        if residual is not None:
            residual = _dense_cl(residual, x.dtype)
        y = torch.empty_like(x)
        L.call("affgw_norm_apply", x.data_ptr(), L.dt(x), mean.data_ptr(), rstd.data_ptr(), L.ptr(gamma), L.ptr(beta),
               L.ptr(residual), y.data_ptr(), G, P, Cc, L.ACT[act], 1, L.stream(), nbytes=_nb(x) + _nb(y) + _nb(residual))
        ctx.act, ctx.unbiased, ctx.affine, ctx.has_res = act, unbiased, gamma is not None, residual is not None
        ctx.save_for_backward(x, mean, rstd, gamma, beta)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd, gamma, beta = ctx.saved_tensors
        G, P, Cc = _gpc(x, True)
        dy = _dense_cl(dy, x.dtype)
        s1 = torch.empty(G * Cc, dtype=torch.float32, device=x.device)
        s2 = torch.empty(G * Cc, dtype=torch.float32, device=x.device)
        dx = torch.empty_like(x)
        L.call("affgw_norm_bwd", dy.data_ptr(), x.data_ptr(), L.dt(x), mean.data_ptr(), rstd.data_ptr(), L.ptr(gamma),
               L.ptr(beta), s1.data_ptr(), s2.data_ptr(), dx.data_ptr(), G, P, Cc, L.ACT[ctx.act], 1, 1, int(ctx.unbiased),
               L.stream(), nbytes=_nb(dy) + _nb(x) + _nb(dx))
        return dx, (s2 if ctx.affine else None), (s1 if ctx.affine else None), (dy if ctx.has_res else None), None, None, None


def instance_norm(x, act="none", gamma=None, beta=None, residual=None, eps=1e-5, unbiased=False):
    return _InstanceNorm.apply(x, gamma, beta, residual, act, eps, unbiased)


class _BatchNorm(Function):
    """nn.BatchNorm1d / nn.BatchNorm2d (blocks.py:250-281, modules_tro.py:275,278) with fused activation."""

    @staticmethod
    def forward(ctx, x, weight, bias, buffers, training, momentum, eps, act):
        x = _dense_cl(x)
        G, P, Cc = _gpc(x, False)
        running_mean, running_var, nbt = buffers
        if training:
            if P <= 1:
                raise ValueError("Expected more than 1 value per channel when training, got input size "
                                 + str(list(x.shape)))
            mean, rstd, var = _stats(x, G, P, Cc, eps, False, want_var=True)
            if running_mean is not None:
                L.call("affgw_bn_update_running", running_mean.data_ptr(), running_var.data_ptr(), L.ptr(nbt),
                       mean.data_ptr(), var.data_ptr(), Cc, float(momentum), L.stream())
                if _state["bn_record"] is not None:         # ops.bn_updates_twice: replayed in order when the context exits
                    _state["bn_record"].append((running_mean, running_var, nbt, mean, var, Cc, float(momentum)))
        else:
            mean = torch.empty(Cc, dtype=torch.float32, device=x.device)
            rstd = torch.empty(Cc, dtype=torch.float32, device=x.device)
            L.call("affgw_bn_eval_stats", running_mean.data_ptr(), running_var.data_ptr(), mean.data_ptr(),
                   rstd.data_ptr(), Cc, float(eps), L.stream())
        w32, b32 = weight.detach(), bias.detach()
        y = torch.empty_like(x)
        L.call("affgw_norm_apply", x.data_ptr(), L.dt(x), mean.data_ptr(), rstd.data_ptr(), w32.data_ptr(), b32.data_ptr(),
               None, y.data_ptr(), G, P, Cc, L.ACT[act], 0, L.stream(), nbytes=_nb(x) + _nb(y))
        ctx.act, ctx.training = act, training
        ctx.save_for_backward(x, mean, rstd, weight, bias)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd, weight, bias = ctx.saved_tensors
        G, P, Cc = _gpc(x, False)
        dy = _dense_cl(dy, x.dtype)
        s1 = torch.empty(Cc, dtype=torch.float32, device=x.device)
        s2 = torch.empty(Cc, dtype=torch.float32, device=x.device)
        dx = torch.empty_like(x)
        L.call("affgw_norm_bwd", dy.data_ptr(), x.data_ptr(), L.dt(x), mean.data_ptr(), rstd.data_ptr(),
               weight.detach().data_ptr(), bias.detach().data_ptr(), s1.data_ptr(), s2.data_ptr(), dx.data_ptr(), G, P, Cc,
               L.ACT[ctx.act], 0, int(ctx.training), 0, L.stream(), nbytes=_nb(dy) + _nb(x) + _nb(dx))
        return dx, s2, s1, None, None, None, None, None


def batch_norm(x, bn, act="none"):
    """`bn` is an nn.BatchNorm{1,2}d used purely as the parameter / buffer container (state_dict keys)."""
    use_batch = bn.training or bn.running_mean is None
    bufs = (bn.running_mean if bn.training else bn.running_mean, bn.running_var, bn.num_batches_tracked)
    return _BatchNorm.apply(x, bn.weight, bn.bias, bufs, use_batch, bn.momentum, bn.eps, act)


# ------------------------------------------------------------------------------------------------ pooling / resize
class _MaxPool2(Function):
    @staticmethod
    def forward(ctx, x):
        x = _dense_cl(x)
        n, c, h, w = x.shape
        y = empty_cl(n, c, h // 2, w // 2, x.dtype, x.device)
        L.call("affgw_maxpool2_fwd", x.data_ptr(), y.data_ptr(), L.dt(x), n, h, w, c, L.stream(), nbytes=_nb(x) + _nb(y))
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        n, c, h, w = x.shape
        dy = _dense_cl(dy, x.dtype)
        dx = torch.empty_like(x)
        L.call("affgw_maxpool2_bwd", dy.data_ptr(), x.data_ptr(), dx.data_ptr(), L.dt(x), n, h, w, c, L.stream(),
               nbytes=_nb(dy) + _nb(x) + _nb(dx))
        return dx


def max_pool2(x):
    return _MaxPool2.apply(x)


class _MaxPool3(Function):
    """nn.MaxPool2d(kernel_size=3, stride=(sy, sx), padding=1): (2, 2) is the torchvision ResNet stem, (2, 1) and (1, 1) are
    Resnet18.py:45-46."""

    @staticmethod
    def forward(ctx, x, sy, sx):
        x = _dense_cl(x)
        n, c, h, w = x.shape
        y = empty_cl(n, c, (h - 1) // sy + 1, (w - 1) // sx + 1, x.dtype, x.device)
        L.call("affgw_maxpool3_fwd", x.data_ptr(), y.data_ptr(), L.dt(x), n, h, w, c, sy, sx, L.stream())
        ctx.save_for_backward(x)
        ctx.strides = (sy, sx)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        n, c, h, w = x.shape
        dy = _dense_cl(dy, x.dtype)
        dx = torch.empty_like(x)
        L.call("affgw_maxpool3_bwd", dy.data_ptr(), x.data_ptr(), dx.data_ptr(), L.dt(x), n, h, w, c, *ctx.strides, L.stream())
        return dx, None, None


def max_pool3(x, stride=2):
    sy, sx = stride if isinstance(stride, (tuple, list)) else (stride, stride)
    return _MaxPool3.apply(x, int(sy), int(sx))


def max_pool3s2(x):
    return _MaxPool3.apply(x, 2, 2)


class _ResizeBilinear(Function):
    """F.interpolate(size=(ho, wo), mode="bilinear", align_corners=False) (modules_tro.py:527)."""

    @staticmethod
    def forward(ctx, x, ho, wo):
        x = _dense_cl(x)
        n, c, h, w = x.shape
        y = empty_cl(n, c, ho, wo, x.dtype, x.device)
        L.call("affgw_resize_bilinear_fwd", x.data_ptr(), y.data_ptr(), L.dt(x), n, h, w, c, ho, wo, L.stream())
        ctx.shape = (n, c, h, w, ho, wo)
        return y

    @staticmethod
    def backward(ctx, dy):
        n, c, h, w, ho, wo = ctx.shape
        dy = _dense_cl(dy)
        dx = empty_cl(n, c, h, w, torch.float32, dy.device, zero=True)
        L.call("affgw_resize_bilinear_bwd", dy.data_ptr(), dx.data_ptr(), L.dt(dy), n, h, w, c, ho, wo, L.stream())
        return dx, None, None


def resize_bilinear(x, ho, wo):
    return _ResizeBilinear.apply(x, ho, wo)


class _AddAct(Function):
    """act(a + b): `out += residual; out = relu(out)` of the ResNet blocks."""

    @staticmethod
    def forward(ctx, a, b, act):
        a = _dense_cl(a)
        b = _dense_cl(b, a.dtype)
        y = torch.empty_like(a)
        L.call("affgw_add_act", a.data_ptr(), b.data_ptr(), y.data_ptr(), L.dt(a), a.numel(), L.ACT[act], L.stream())
        ctx.act = act
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = _dense_cl(dy, y.dtype)
        dz = torch.empty_like(y)
        L.call("affgw_act_bwd", dy.data_ptr(), y.data_ptr(), dz.data_ptr(), L.dt(y), y.numel(), L.ACT[ctx.act], L.stream())
        return dz, dz, None


def add_act(a, b, act="relu"):
    return _AddAct.apply(a, b, act)


class _AvgPool3s2Reflect(Function):
    @staticmethod
    def forward(ctx, x):
        x = _dense_cl(x)
        n, c, h, w = x.shape
        y = empty_cl(n, c, (h - 1) // 2 + 1, (w - 1) // 2 + 1, x.dtype, x.device)
        L.call("affgw_avgpool3s2_fwd", x.data_ptr(), y.data_ptr(), L.dt(x), n, h, w, c, L.stream(), nbytes=_nb(x) + _nb(y))
        ctx.shape = (n, c, h, w)
        return y

    @staticmethod
    def backward(ctx, dy):
        n, c, h, w = ctx.shape
        dy = _dense_cl(dy)
        dx = empty_cl(n, c, h, w, dy.dtype, dy.device)
        L.call("affgw_avgpool3s2_bwd", dy.data_ptr(), dx.data_ptr(), L.dt(dy), n, h, w, c, L.stream(), nbytes=_nb(dy) + _nb(dx))
        return dx


def avg_pool3s2_reflect(x):
    """nn.ReflectionPad2d(1) + nn.AvgPool2d(3, 2) (modules_tro.py:133-134)."""
    return _AvgPool3s2Reflect.apply(x)


class _ResizeNearest(Function):
    @staticmethod
    def forward(ctx, x, ho, wo):
        x = _dense_cl(x)
        n, c, h, w = x.shape
        y = empty_cl(n, c, ho, wo, x.dtype, x.device)
        L.call("affgw_resize_nearest_fwd", x.data_ptr(), y.data_ptr(), L.dt(x), n, h, w, c, ho, wo, L.stream())
        ctx.shape = (n, c, h, w, ho, wo)
        return y

    @staticmethod
    def backward(ctx, dy):
        n, c, h, w, ho, wo = ctx.shape
        dy = _dense_cl(dy)
        dx = empty_cl(n, c, h, w, dy.dtype, dy.device)
        L.call("affgw_resize_nearest_bwd", dy.data_ptr(), dx.data_ptr(), L.dt(dy), n, h, w, c, ho, wo, L.stream())
        return dx, None, None


def resize_nearest(x, ho, wo):
    if x.shape[2] == ho and x.shape[3] == wo:
        return x
    return _ResizeNearest.apply(x, ho, wo)


# ------------------------------------------------------------------------------------------------ iAFF pieces
class _Gate(Function):
    @staticmethod
    def forward(ctx, x, r, xl, xg):
        x, r, xl = _dense_cl(x), _dense_cl(r, x.dtype), _dense_cl(xl, x.dtype)
        xg = _dense_cl(xg, x.dtype)
        n, c, h, w = x.shape
        y = torch.empty_like(x)
        L.call("affgw_gate_fwd", x.data_ptr(), r.data_ptr(), xl.data_ptr(), xg.data_ptr(), y.data_ptr(), L.dt(x), n, h * w, c,
               L.stream())
        ctx.save_for_backward(x, r, xl, xg)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, r, xl, xg = ctx.saved_tensors
        n, c, h, w = x.shape
        dy = _dense_cl(dy, x.dtype)
        dx, dr, dxl, dxg = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x), torch.empty_like(xg)
        L.call("affgw_gate_bwd", dy.data_ptr(), x.data_ptr(), r.data_ptr(), xl.data_ptr(), xg.data_ptr(), dx.data_ptr(),
               dr.data_ptr(), dxl.data_ptr(), dxg.data_ptr(), L.dt(x), n, h * w, c, L.stream())
        return dx, dr, dxl, dxg


def iaff_gate(x, r, xl, xg):
    """x * w + r * (1 - w), w = sigmoid(xl + xg)  (blocks.py:289-292, 296-298)."""
    return _Gate.apply(x, r, xl, xg)


class _Gap(Function):
    @staticmethod
    def forward(ctx, x):
        x = _dense_cl(x)
        n, c, h, w = x.shape
        y = empty_cl(n, c, 1, 1, x.dtype, x.device)
        L.call("affgw_gap_fwd", x.data_ptr(), y.data_ptr(), L.dt(x), n, h * w, c, L.stream())
        ctx.shape = (n, c, h, w)
        return y

    @staticmethod
    def backward(ctx, dy):
        n, c, h, w = ctx.shape
        dy = _dense_cl(dy)
        dx = empty_cl(n, c, h, w, dy.dtype, dy.device)
        L.call("affgw_bcast_add", None, dy.data_ptr(), dx.data_ptr(), L.dt(dy), n, h * w, c, 1.0 / (h * w), L.stream())
        return dx


def global_avg_pool(x):
    return _Gap.apply(x)


class _Add(Function):
    @staticmethod
    def forward(ctx, a, b):
        a = _dense_cl(a)
        b = _dense_cl(b, a.dtype)
        y = torch.empty_like(a)
        L.call("affgw_add2", a.data_ptr(), b.data_ptr(), y.data_ptr(), L.dt(a), a.numel(), L.stream())
        return y

    @staticmethod
    def backward(ctx, dy):
        return dy, dy


def add(a, b):
    return _Add.apply(a, b)


# ------------------------------------------------------------------------------------------------ text encoder pieces
class _Embedding(Function):
    @staticmethod
    def forward(ctx, ids, table, dtype):
        L.require_cuda(ids, table)
        if ids.dtype != torch.int64:
            raise RuntimeError("embedding: token ids must be int64")
        ids = ids.contiguous()
        v, e = table.shape
        out = torch.empty(ids.shape + (e,), dtype=dtype, device=ids.device)
        L.call("affgw_embedding_fwd", ids.data_ptr(), table.detach().data_ptr(), out.data_ptr(), L.dt(out), ids.numel(), e, v,
               _err_tensor(ids.device).data_ptr(), L.stream())
        ctx.save_for_backward(ids)
        ctx.shape = (v, e)
        return out

    @staticmethod
    def backward(ctx, dout):
        (ids,) = ctx.saved_tensors
        v, e = ctx.shape
        dout = dout.contiguous()
        dt = torch.zeros((v, e), dtype=torch.float32, device=dout.device)
        L.call("affgw_embedding_bwd", ids.data_ptr(), dout.data_ptr(), dt.data_ptr(), L.dt(dout), ids.numel(), e, v, L.stream())
        return None, dt, None


def embedding(ids, table, dtype=None):
    return _Embedding.apply(ids, table, dtype or act_dtype())


class _TextTile(Function):
    @staticmethod
    def forward(ctx, chars, h, w, reps):
        chars = chars.contiguous()
        b, slots, c = chars.shape
        out = empty_cl(b, c, h, w, chars.dtype, chars.device)
        L.call("affgw_text_tile_fwd", chars.data_ptr(), out.data_ptr(), L.dt(chars), b, h, w, c, slots - 1, reps, L.stream())
        ctx.args = (b, slots, c, h, w, reps)
        return out

    @staticmethod
    def backward(ctx, dout):
        b, slots, c, h, w, reps = ctx.args
        dout = _dense_cl(dout)
        dch = torch.empty((b, slots, c), dtype=dout.dtype, device=dout.device)
        L.call("affgw_text_tile_bwd", dout.data_ptr(), dch.data_ptr(), L.dt(dout), b, h, w, c, slots - 1, reps, L.stream())
        return dch, None, None, None


def text_tile(chars, h, w, reps):
    """chars [B, ts+1, C] (last slot = PAD) -> content map [B, C, h, w]  (modules_tro.py:295-317)."""
    return _TextTile.apply(chars, h, w, reps)


# ------------------------------------------------------------------------------------------------ losses
class _BceLogits(Function):
    @staticmethod
    def forward(ctx, x, target):
        x = x.contiguous()
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        L.call("affgw_bce_logits_fwd", x.data_ptr(), L.dt(x), float(target), loss.data_ptr(), x.numel(), L.stream())
        ctx.save_for_backward(x)
        ctx.target = float(target)
        return loss

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        g = g.float().contiguous()
        dx = torch.empty_like(x)
        L.call("affgw_bce_logits_bwd", x.data_ptr(), L.dt(x), ctx.target, g.data_ptr(), dx.data_ptr(), x.numel(), L.stream())
        return dx, None


def bce_with_logits_const(x, target):
    """nn.BCEWithLogitsLoss against an all-`target` label tensor (modules_tro.py:152-168)."""
    return _BceLogits.apply(x, target)


class _SoftmaxCe(Function):
    @staticmethod
    def forward(ctx, x, y):
        x = x.contiguous()
        if y.dtype != torch.int64:
            raise RuntimeError("cross_entropy: class indices must be int64")
        y = y.contiguous()
        b, c = x.shape
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        L.call("affgw_softmax_ce_fwd", x.data_ptr(), L.dt(x), y.data_ptr(), loss.data_ptr(), b, c,
               _err_tensor(x.device).data_ptr(), L.stream())
        ctx.save_for_backward(x, y)
        return loss

    @staticmethod
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        b, c = x.shape
        g = g.float().contiguous()
        dx = torch.empty_like(x)
        L.call("affgw_softmax_ce_bwd", x.data_ptr(), L.dt(x), y.data_ptr(), g.data_ptr(), dx.data_ptr(), b, c, L.stream())
        return dx, None


def cross_entropy(x, y):
    return _SoftmaxCe.apply(x, y)


class _LabelSmoothKl(Function):
    @staticmethod
    def forward(ctx, x, y, pad_idx, smoothing):
        L.require_cuda(x, y)
        x = x.float().contiguous()
        if y.dtype != torch.int64:
            raise RuntimeError("label_smoothing_kl: targets must be int64")
        y = y.contiguous()
        rows, v = x.shape
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        L.call("affgw_label_smooth_kl_fwd", x.data_ptr(), y.data_ptr(), loss.data_ptr(), rows, v, int(pad_idx), float(smoothing),
               _err_tensor(x.device).data_ptr(), L.stream())
        ctx.save_for_backward(x, y)
        ctx.cfg = (int(pad_idx), float(smoothing))
        return loss

    @staticmethod
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        rows, v = x.shape
        g = g.float().contiguous()
        dx = torch.empty_like(x)
        L.call("affgw_label_smooth_kl_bwd", x.data_ptr(), y.data_ptr(), g.data_ptr(), dx.data_ptr(), rows, v, ctx.cfg[0],
               ctx.cfg[1], L.stream())
        return dx, None, None, None


def label_smoothing_kl(x, y, pad_idx, smoothing):
    """crit(log_softmax(x), y) of the reference (loss_tro.py:8-35, network_tro.py:44-45): KL(sum) against the smoothed
    one-hot; x [rows, V] logits, y [rows] int64."""
    return _LabelSmoothKl.apply(x, y, pad_idx, smoothing)


# ------------------------------------------------------------------------------------------------ line-level generator pieces
def conv_transpose2d(x, weight, bias=None, stride=2, pad=1):
    """F.conv_transpose2d(x, weight [Cin, Cout, K, K], bias, stride, padding) - FusedUpsample of the line-level generator
    (line_generation/model/pure_gen.py:268-279) - computed as what it is: the input gradient of the stride-`stride` convolution
    with the same weight, on the dgrad kernels (zero-insertion gather).  Forward only (generation); x [N, Cin, H, W] internal."""
    if torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad):
        raise RuntimeError("conv_transpose2d is a generation-only operator (call it under torch.no_grad())")
    x = _dense_cl(input_to_internal(x))
    n, cv, h, w = x.shape
    cv_w, cu, kh, kw = weight.shape
    if cv_w != cv or kh != kw:
        raise RuntimeError("conv_transpose2d: weight must be [Cin, Cout, K, K] with Cin matching the input")
    hu, wu = (h - 1) * stride - 2 * pad + kh, (w - 1) * stride - 2 * pad + kw
    # the forward convolution whose input gradient this is: U [n, cu, hu, wu] -> V [n, cv, h, w]
    g = dict(N=n, H=hu, W=wu, Cx=cu, pitch=cu, Cout=cv, Cin=cu, KH=kh, KW=kw, Ho=h, Wo=w, two_d=False)
    cfg = ConvCfg(int(stride), int(pad), "zero", 1, "none", "none", None, int(stride))
    st = L.stream()
    out = empty_cl(n, cu, hu, wu, torch.float32, x.device)
    use_tc = _state["mode"] != "fp32" and not _state["force_simt"]
    if use_tc:
        passes = _state["passes"][0]                      # this is a FORWARD computation: the mode's forward pass count
        fmt = L.FMT_F16 if _state["mode"] == "f16" else L.FMT_BF16
        cs, cso = _up8(cu), _up8(cv)
        d = _desc(g, cfg, cu, L.BF16, L.BF16, L.BF16, L.ALGO_TC, in_pitch=cs, out_pitch=cso, passes=passes, grad_dt=L.F32, fmt=fmt)
        layout = L.lib().affgw_conv_tc_layout(C.byref(d), 1)
        if not layout:
            raise RuntimeError("conv_transpose2d: tcgen05 kernels refused the shape: " + L.last_error())
        if layout == L.WLAYOUT_SHIFT:
            raise RuntimeError("conv_transpose2d: unexpected position-space layout for a strided convolution")
        planes = _split_planes(x, n * h * w, cv, cv, passes, fmt=fmt)
        wt = _pack_tc(weight, cso, True, passes, layout, fmt)
        src = planes
    else:
        d = _desc(g, cfg, cu, L.F32, L.F32, L.F32, L.ALGO_SIMT)
        wt = _pack(weight, torch.float32, cv, True)
        src = x
    ws_bytes = L.lib().affgw_conv2d_dgrad_ws_bytes(C.byref(d))
    if ws_bytes < 0:
        raise RuntimeError("conv_transpose2d: " + L.last_error())
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device) if ws_bytes else None
    L.call("affgw_conv2d_dgrad_scaled", src.data_ptr(), wt.data_ptr(), None, out.data_ptr(), L.ptr(ws), C.byref(d), None, st)
    if bias is not None:
        b = bias.detach().float().reshape(1, cu).expand(n, cu).contiguous()
        res = torch.empty_like(out)
        L.call("affgw_bcast_add", out.data_ptr(), b.data_ptr(), res.data_ptr(), L.dt(out), n, hu * wu, cu, 1.0, st)
        out = res
    return out


class _Blur3(Function):
    """Depthwise 3x3 binomial blur, zero padding (Blur, pure_gen.py:123-136); symmetric kernel: backward = the same blur."""

    @staticmethod
    def forward(ctx, x):
        x = _dense_cl(x)
        n, c, h, w = x.shape
        y = torch.empty_like(x)
        L.call("affgw_blur3", x.data_ptr(), y.data_ptr(), n, h, w, c, L.stream(), nbytes=2 * _nb(x))
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = _dense_cl(dy, torch.float32)
        n, c, h, w = dy.shape
        dx = torch.empty_like(dy)
        L.call("affgw_blur3", dy.data_ptr(), dx.data_ptr(), n, h, w, c, L.stream(), nbytes=2 * _nb(dy))
        return dx


def blur3(x):
    return _Blur3.apply(input_to_internal(x))


def pixel_norm(x, eps=1e-8):
    """x / sqrt(mean(x^2, dim=1) + eps) for a [rows, C] tensor (PixelNorm, pure_gen.py:306-311).  Forward only."""
    L.require_cuda(x)
    x = x.detach().float().contiguous()
    y = torch.empty_like(x)
    L.call("affgw_pixelnorm", x.data_ptr(), y.data_ptr(), x.shape[0], x.shape[1], float(eps), L.stream())
    return y
