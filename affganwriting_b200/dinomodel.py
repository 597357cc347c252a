"""DINOv2 ViT style encoder (BASELINE.json configs[3]): drop-in for the reference's wrapper `GAN_word/dinomodel.py:7-166`
(`ImageEncoderDINOv2`: 50-plane patch embedding, replicate-pad to multiples of 14, five taps - the stem tokens and four
transformer blocks - each reduced to 512 channels by a 1x1 convolution, the last one bilinearly resized to 8 x 27).

The reference obtains the backbone with `torch.hub.load(<local dinov2 checkout>, "dinov2_<arch>", source="local")`; that
checkout is not part of the reference tree.  Here the backbone is built directly (`DinoVisionTransformer`: the public DINOv2
ViT definition - pre-LayerNorm blocks with eps 1e-6, qkv bias, LayerScale on both residual branches, erf-GELU MLP, flat
`blocks`, no register tokens - with the hub model's parameter names, so a DINOv2 checkpoint given as `ckpt_path` loads), and
every forward runs on libaffgw: patch embedding, qkv / proj / fc1 / fc2 and the reducers on the tcgen05 GEMM kernels,
LayerNorm / attention / GELU / LayerScale-residual on `vit.cu`, the resize on the bilinear kernel.

GENERATION ONLY (configs[3] is batch-256 inference): the transformer kernels have no backward.  Like the reference's fallback
path (`_pos_embed_tokens`, dinomodel.py:103-117), positional embeddings are added only when the stored grid happens to have
exactly 1 + Hp*Wp entries - for the hub model's 37 x 37 grid and a 64 x 216 image (5 x 16 patches) they are not.
"""
import os

import torch
from torch import nn

from . import _lib as L  # noqa: N812
from . import ops

ARCHS = {"vits14": dict(embed_dim=384, depth=12, num_heads=6), "vitb14": dict(embed_dim=768, depth=12, num_heads=12),
         "vitl14": dict(embed_dim=1024, depth=24, num_heads=16)}


def layer_norm(x, m):
    x = x.contiguous()
    y = torch.empty_like(x)
    L.call("affgw_layernorm_fwd", x.data_ptr(), m.weight.data_ptr(), m.bias.data_ptr(), y.data_ptr(), x.numel() // x.shape[-1],
           x.shape[-1], float(m.eps), L.stream())
    return y


def gelu(x):
    y = torch.empty_like(x)
    L.call("affgw_gelu_fwd", x.data_ptr(), y.data_ptr(), x.numel(), L.stream())
    return y


def scale_residual(x, t, gamma):
    y = torch.empty_like(x)
    L.call("affgw_scale_residual", x.data_ptr(), t.contiguous().data_ptr(), L.ptr(gamma), y.data_ptr(), x.numel(), x.shape[-1], L.stream())
    return y


def attention(qkv, batch, tokens, heads):
    """qkv [batch * tokens, 3 D] (the Linear's output, i.e. [B][N][3][H][hd]) -> [batch * tokens, D]"""
    d = qkv.shape[-1] // 3
    out = torch.empty((batch * tokens, d), dtype=torch.float32, device=qkv.device)
    L.call("affgw_attention_fwd", qkv.data_ptr(), out.data_ptr(), batch, tokens, heads, d // heads, float((d // heads) ** -0.5),
           L.stream())
    return out


class _PatchEmbed(nn.Module):
    def __init__(self, dim, patch=14, in_chans=3):
        super().__init__()
        self.patch_size = (patch, patch)
        self.proj = nn.Conv2d(in_chans, dim, kernel_size=patch, stride=patch)
        self.norm = nn.Identity()


class _Attention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.num_heads = heads
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)


class _LayerScale(nn.Module):
    def __init__(self, dim, init=1.0):
        super().__init__()
        self.gamma = nn.Parameter(init * torch.ones(dim))


class _Mlp(nn.Module):
    def __init__(self, dim, ratio=4):
        super().__init__()
        self.fc1, self.act, self.fc2 = nn.Linear(dim, dim * ratio), nn.GELU(), nn.Linear(dim * ratio, dim)


class Block(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.norm1, self.attn, self.ls1 = nn.LayerNorm(dim, eps=1e-6), _Attention(dim, heads), _LayerScale(dim)
        self.norm2, self.mlp, self.ls2 = nn.LayerNorm(dim, eps=1e-6), _Mlp(dim), _LayerScale(dim)

    def forward(self, tok, batch, tokens):
        """tok [batch * tokens, D] -> same:  x + ls1(attn(norm1(x))), then x + ls2(mlp(norm2(x)))"""
        a = self.attn
        h = ops.linear(layer_norm(tok, self.norm1), a.qkv.weight, a.qkv.bias)
        h = ops.linear(attention(h, batch, tokens, a.num_heads), a.proj.weight, a.proj.bias)
        tok = scale_residual(tok, h, self.ls1.gamma)
        h = gelu(ops.linear(layer_norm(tok, self.norm2), self.mlp.fc1.weight, self.mlp.fc1.bias))
        h = ops.linear(h, self.mlp.fc2.weight, self.mlp.fc2.bias)
        return scale_residual(tok, h, self.ls2.gamma)


class DinoVisionTransformer(nn.Module):
    """Parameter container with the hub model's names (cls_token, pos_embed, mask_token, patch_embed.proj, blocks.N.*, norm)."""

    def __init__(self, embed_dim=1024, depth=24, num_heads=16, grid=37):
        super().__init__()
        self.embed_dim = self.num_features = embed_dim
        self.num_heads = num_heads
        self.patch_embed = _PatchEmbed(embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, grid * grid + 1, embed_dim))
        self.mask_token = nn.Parameter(torch.zeros(1, embed_dim))
        self.blocks = nn.ModuleList([Block(embed_dim, num_heads) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)


class ImageEncoderDINOv2(nn.Module):
    def __init__(self, repo_dir=None, arch="vitl14", ckpt_path=None, in_channels=50, final_size=(8, 27), tap_blocks=None,
                 probe_size=(48, 540)):
        """Same arguments as dinomodel.py:21-30.  `repo_dir` (the local torch.hub checkout the reference loads the class from) is
        accepted and ignored; `arch` is "vits14" | "vitb14" | "vitl14" (the giant model's SwiGLU MLP is not built) or a dict
        {embed_dim, depth, num_heads}."""
        super().__init__()
        self.output_dim = 512
        self.final_size = tuple(final_size)
        if isinstance(arch, str):
            if arch not in ARCHS:
                raise ValueError(f"unsupported DINOv2 arch {arch!r} (have {sorted(ARCHS)})")
            arch = ARCHS[arch]
        self.model = DinoVisionTransformer(**arch)
        if ckpt_path is not None and os.path.isfile(ckpt_path):               # dinomodel.py:42-51 (lenient load)
            sd = torch.load(ckpt_path, map_location="cpu")
            if isinstance(sd, dict) and "state_dict" in sd:
                sd = sd["state_dict"]
            elif isinstance(sd, dict) and "model" in sd:
                sd = sd["model"]
            self.model.load_state_dict({(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}, strict=False)
        patch = self.model.patch_embed
        old = patch.proj                                                      # dinomodel.py:53-72: in_channels != 3
        new = nn.Conv2d(in_channels, old.out_channels, kernel_size=old.kernel_size, stride=old.stride, padding=old.padding,
                        bias=old.bias is not None)
        with torch.no_grad():
            new.weight[:, :3] = old.weight
            if in_channels > 3:
                new.weight[:, 3:] = old.weight[:, :1].repeat(1, in_channels - 3, 1, 1)
        patch.proj = new
        self.embed_dim = self.model.embed_dim
        self.patch_size = patch.patch_size
        n_blocks = len(self.model.blocks)
        if tap_blocks is None:
            idxs = torch.linspace(0, n_blocks - 1, steps=4).round().to(torch.int64).tolist()
            tap_blocks = sorted(set(idxs))
        self.tap_blocks = list(tap_blocks)
        self.reduce_layers = nn.ModuleList([nn.Conv2d(self.embed_dim, 512, kernel_size=1) for _ in range(1 + len(self.tap_blocks))])

    @torch.no_grad()
    def encode_with_intermediate(self, x):
        B, _, H, W = x.shape
        ph, pw = self.patch_size
        pad_h, pad_w = (ph - H % ph) % ph, (pw - W % pw) % pw
        if pad_h or pad_w:
            x = torch.nn.functional.pad(x.float(), (0, pad_w, 0, pad_h), mode="replicate")      # dinomodel.py:131-137
        Hp, Wp = x.shape[-2] // ph, x.shape[-1] // pw
        m = self.model
        D = self.embed_dim
        proj = m.patch_embed.proj
        spatial = ops.conv2d(ops.input_to_internal(x), proj.weight, proj.bias, stride=ph, pad=0)   # [B, D, Hp, Wp], NHWC storage
        n_tok = 1 + Hp * Wp
        tok = torch.empty((B, n_tok, D), dtype=torch.float32, device=x.device)
        tok[:, 0, :] = m.cls_token.reshape(1, D)
        tok[:, 1:, :] = ops._dense_cl(spatial).permute(0, 2, 3, 1).reshape(B, Hp * Wp, D)
        if m.pos_embed.shape[1] == n_tok:                                     # dinomodel.py:112-114
            tok = tok + m.pos_embed
        tok = tok.reshape(B * n_tok, D)

        def tap(i, t):
            fmap = t.reshape(B, n_tok, D)[:, 1:, :].reshape(B, Hp, Wp, D).permute(0, 3, 1, 2)      # channels-last view of the tokens
            r = self.reduce_layers[i]
            return ops.conv2d(ops.input_to_internal(fmap), r.weight, r.bias)
        results, red = [tap(0, tok)], 1
        for i, blk in enumerate(m.blocks):
            tok = blk(tok, B, n_tok)
            if i in self.tap_blocks:
                results.append(tap(red, tok))
                red += 1
        results[-1] = ops.resize_bilinear(results[-1], *self.final_size)
        return results

    def forward(self, x):
        return self.encode_with_intermediate(x)
