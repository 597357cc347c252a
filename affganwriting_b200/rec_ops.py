"""Autograd operators of the recogniser that are not convolutions / GEMMs (libaffgw `rec.cu`): GRU cell, Dropout2d scaling,
feature-map -> sequence permutation, location-attention energy, attention soft-max + context.  fp32 CUDA tensors only."""
import torch
from torch.autograd import Function

from . import _lib as L  # noqa: N812


def _f32c(t):
    L.require_cuda(t)
    if t.dtype != torch.float32:
        raise RuntimeError("recogniser operators take float32 tensors")
    return t if t.is_contiguous() else t.contiguous()


class _GruCell(Function):
    """torch.nn.GRU cell (gate order r, z, n): h' = (1 - z) n + z h from gi = W_ih x + b_ih and gh = W_hh h + b_hh."""

    @staticmethod
    def forward(ctx, gi, gh, h):
        L.require_cuda(gi, gh, h)
        n, hid = h.shape
        if gi.stride(1) != 1:
            gi = gi.contiguous()
        gh, h = _f32c(gh), _f32c(h)
        out = torch.empty_like(h)
        L.call("affgw_gru_cell_fwd", gi.data_ptr(), gi.stride(0) if n > 1 else 3 * hid, gh.data_ptr(), h.data_ptr(), out.data_ptr(),
               n, hid, L.stream())
        ctx.save_for_backward(gi, gh, h)
        return out

    @staticmethod
    def backward(ctx, dout):
        gi, gh, h = ctx.saved_tensors
        n, hid = h.shape
        dout = _f32c(dout)
        dgi = torch.empty((n, 3 * hid), dtype=torch.float32, device=h.device)
        dgh, dh = torch.empty_like(dgi), torch.empty_like(h)
        L.call("affgw_gru_cell_bwd", dout.data_ptr(), gi.data_ptr(), gi.stride(0) if n > 1 else 3 * hid, gh.data_ptr(), h.data_ptr(),
               dgi.data_ptr(), dgh.data_ptr(), dh.data_ptr(), n, hid, L.stream())
        return dgi, dgh, dh


def gru_cell(gi, gh, h):
    return _GruCell.apply(gi, gh, h)


class _ScaleNC(Function):
    """x [N, C, H, W] (channels-last) times a per-(sample, channel) factor m [N, C]: Dropout2d with the caller's mask."""

    @staticmethod
    def forward(ctx, x, m):
        from . import ops
        x = ops._dense_cl(x)
        n, c, h, w = x.shape
        m = _f32c(m.reshape(n, c))
        y = torch.empty_like(x)
        L.call("affgw_scale_nc", x.data_ptr(), m.data_ptr(), y.data_ptr(), n, h * w, c, L.stream())
        ctx.save_for_backward(m)
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import ops
        (m,) = ctx.saved_tensors
        dy = ops._dense_cl(dy, torch.float32)
        n, c, h, w = dy.shape
        dx = torch.empty_like(dy)
        L.call("affgw_scale_nc", dy.data_ptr(), m.data_ptr(), dx.data_ptr(), n, h * w, c, L.stream())
        return dx, None


def scale_nc(x, m):
    return _ScaleNC.apply(x, m)


class _MulConst(Function):
    """a * mask for a mask that needs no gradient (inter-layer GRU dropout)."""

    @staticmethod
    def forward(ctx, a, mask):
        a, mask = _f32c(a), _f32c(mask)
        if a.shape != mask.shape:
            raise RuntimeError("mul: shapes differ")
        y = torch.empty_like(a)
        L.call("affgw_mul2", a.data_ptr(), mask.data_ptr(), y.data_ptr(), a.numel(), L.stream())
        ctx.save_for_backward(mask)
        return y

    @staticmethod
    def backward(ctx, dy):
        (mask,) = ctx.saved_tensors
        dy = _f32c(dy)
        da = torch.empty_like(dy)
        L.call("affgw_mul2", dy.data_ptr(), mask.data_ptr(), da.data_ptr(), dy.numel(), L.stream())
        return da, None


def mul_mask(a, mask):
    return _MulConst.apply(a, mask)


class _MapToSeq(Function):
    """[B, C, H, W] channels-last feature map -> [W, B, H * C] sequence (encoder_vgg.py:711-713)."""

    @staticmethod
    def forward(ctx, x):
        from . import ops
        x = ops._dense_cl(x)
        b, c, h, w = x.shape
        y = torch.empty((w, b, h * c), dtype=torch.float32, device=x.device)
        L.call("affgw_map_seq", x.data_ptr(), y.data_ptr(), b, h, w, c, 1, L.stream())
        ctx.dims = (b, c, h, w)
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import ops
        b, c, h, w = ctx.dims
        dy = _f32c(dy)
        dx = ops.empty_cl(b, c, h, w, torch.float32, dy.device)
        L.call("affgw_map_seq", dy.data_ptr(), dx.data_ptr(), b, h, w, c, 0, L.stream())
        return dx


def map_to_seq(x):
    return _MapToSeq.apply(x)


class _AttnEnergy(Function):
    """energy[n, t] = v . tanh(e[sample[n], t] + hp[n] + loc[n, t]) + vb   (attention.py:145-158)."""

    @staticmethod
    def forward(ctx, e, sample, hp, loc, v, vb):
        e, hp, loc, v, vb = _f32c(e), _f32c(hp), _f32c(loc), _f32c(v), _f32c(vb)
        n, t, f = loc.shape
        energy = torch.empty((n, t), dtype=torch.float32, device=loc.device)
        L.call("affgw_attn_energy_fwd", e.data_ptr(), sample.data_ptr(), hp.data_ptr(), loc.data_ptr(), v.data_ptr(), vb.data_ptr(),
               energy.data_ptr(), n, t, f, L.stream())
        ctx.save_for_backward(e, sample, hp, loc, v)
        ctx.vb_shape = vb.shape
        return energy

    @staticmethod
    def backward(ctx, denergy):
        e, sample, hp, loc, v = ctx.saved_tensors
        n, t, f = loc.shape
        denergy = _f32c(denergy)
        de, dhp, dv = torch.zeros_like(e), torch.zeros_like(hp), torch.zeros_like(v)
        dvb = torch.zeros(ctx.vb_shape, dtype=torch.float32, device=v.device)
        dloc = torch.empty_like(loc)
        L.call("affgw_attn_energy_bwd", denergy.data_ptr(), e.data_ptr(), sample.data_ptr(), hp.data_ptr(), loc.data_ptr(),
               v.data_ptr(), de.data_ptr(), dhp.data_ptr(), dloc.data_ptr(), dv.data_ptr(), dvb.data_ptr(), n, t, f, L.stream())
        return de, None, dhp, dloc, dv, dvb


def attn_energy(e, sample, hp, loc, v, vb):
    return _AttnEnergy.apply(e, sample, hp, loc, v, vb)


class _AttnCtx(Function):
    """(attn, ctx): attn[n] = softmax_t(energy[n]); ctx[n] = sum_t attn[n, t] enc[sample[n], t]   (decoder.py:36-40)."""

    @staticmethod
    def forward(ctx_, energy, enc, sample):
        energy, enc = _f32c(energy), _f32c(enc)
        n, t = energy.shape
        f = enc.shape[-1]
        attn = torch.empty_like(energy)
        ctx = torch.empty((n, f), dtype=torch.float32, device=enc.device)
        L.call("affgw_attn_ctx_fwd", energy.data_ptr(), enc.data_ptr(), sample.data_ptr(), attn.data_ptr(), ctx.data_ptr(), n, t, f,
               L.stream())
        ctx_.save_for_backward(attn, enc, sample)
        return attn, ctx

    @staticmethod
    def backward(ctx_, dattn, dctx):
        attn, enc, sample = ctx_.saved_tensors
        n, t = attn.shape
        f = enc.shape[-1]
        dctx = _f32c(dctx) if dctx is not None else torch.zeros((n, f), dtype=torch.float32, device=enc.device)
        dattn = _f32c(dattn) if dattn is not None else None
        denergy = torch.empty_like(attn)
        denc = torch.zeros_like(enc)
        L.call("affgw_attn_ctx_bwd", L.ptr(dattn), dctx.data_ptr(), attn.data_ptr(), enc.data_ptr(), sample.data_ptr(),
               denergy.data_ptr(), denc.data_ptr(), n, t, f, L.stream())
        return denergy, denc, None


def attn_softmax_context(energy, enc, sample):
    return _AttnCtx.apply(energy, enc, sample)
