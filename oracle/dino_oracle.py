"""TEST INFRASTRUCTURE - the DINOv2 style encoder of BASELINE.json configs[3] on the CPU.

The reference ships only the WRAPPER (GAN_word/dinomodel.py:7-166); its backbone comes from `torch.hub.load(<local checkout of
facebookresearch/dinov2>, "dinov2_vitl14", source="local")` - an un-vendored dependency that is absent here (SURVEY.md §8(c),
appendix E).  So there are two layers of evidence:

  * `StandInViT` restates the PUBLIC DINOv2 ViT definition (pre-LayerNorm blocks with eps 1e-6, qkv bias, LayerScale on both
    residual branches, erf-GELU MLP, no register tokens, flat `blocks`) with the hub model's attribute / parameter names, at any
    size.  It cannot be compared with the real hub module in this container; it IS pinned against an independent implementation
    of the same published model - transformers.Dinov2Model (the implementation that loads the official checkpoints through
    convert_dinov2_to_hf.py's key map) - by oracle/make_golden_dino_hf.py: every block output 1e-7, and HF's layers + the
    wrapper's reducers reproduce the UNMODIFIED reference wrapper's golden maps to 1.6e-7 (tests/golden/dino_hf.npz).
  * the reference WRAPPER itself is pinned: oracle/make_golden_dino.py runs the UNMODIFIED dinomodel.ImageEncoderDINOv2 with
    `torch.hub.load` answered by a small StandInViT, and `dino_encoder` below (a functional restatement of wrapper + blocks over
    the state_dict) is checked against it (tests/golden/dino.npz).
"""
import torch
import torch.nn.functional as F
from torch import nn


class _PatchEmbed(nn.Module):
    def __init__(self, dim, patch=14, in_chans=3):
        super().__init__()
        self.patch_size = (patch, patch)
        self.proj = nn.Conv2d(in_chans, dim, kernel_size=patch, stride=patch)
        self.norm = nn.Identity()

    def forward(self, x):
        return self.norm(self.proj(x).flatten(2).transpose(1, 2))


class _Attention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.num_heads, self.scale = heads, (dim // heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0] * self.scale, qkv[1], qkv[2]
        attn = (q @ k.transpose(-2, -1)).softmax(dim=-1)
        return self.proj((attn @ v).transpose(1, 2).reshape(B, N, C))


class _LayerScale(nn.Module):
    def __init__(self, dim, init=1.0):
        super().__init__()
        self.gamma = nn.Parameter(init * torch.ones(dim))

    def forward(self, x):
        return x * self.gamma


class _Mlp(nn.Module):
    def __init__(self, dim, ratio=4):
        super().__init__()
        self.fc1, self.act, self.fc2 = nn.Linear(dim, dim * ratio), nn.GELU(), nn.Linear(dim * ratio, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class _Block(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.norm1, self.attn, self.ls1 = nn.LayerNorm(dim, eps=1e-6), _Attention(dim, heads), _LayerScale(dim)
        self.norm2, self.mlp, self.ls2 = nn.LayerNorm(dim, eps=1e-6), _Mlp(dim), _LayerScale(dim)

    def forward(self, x):
        x = x + self.ls1(self.attn(self.norm1(x)))
        return x + self.ls2(self.mlp(self.norm2(x)))


class StandInViT(nn.Module):
    """Interface the reference wrapper touches: patch_embed(.proj, .patch_size), cls_token, pos_embed, blocks, embed_dim."""

    def __init__(self, embed_dim=1024, depth=24, num_heads=16, grid=37):
        super().__init__()
        self.embed_dim = self.num_features = embed_dim
        self.num_heads = num_heads
        self.patch_embed = _PatchEmbed(embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, grid * grid + 1, embed_dim))
        self.mask_token = nn.Parameter(torch.zeros(1, embed_dim))
        self.blocks = nn.ModuleList([_Block(embed_dim, num_heads) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)


def dino_encoder(x, sd, heads, tap_blocks, final_size=(8, 27), patch=14):
    """dinomodel.py:127-163 + the blocks, functional over the wrapper's state_dict (`model.*`, `reduce_layers.*`)."""
    B, _, H, W = x.shape
    pad_h, pad_w = (patch - H % patch) % patch, (patch - W % patch) % patch
    if pad_h or pad_w:
        x = F.pad(x, (0, pad_w, 0, pad_h), mode="replicate")
    Hp, Wp = x.shape[-2] // patch, x.shape[-1] // patch
    tok = F.conv2d(x, sd["model.patch_embed.proj.weight"], sd["model.patch_embed.proj.bias"], stride=patch).flatten(2).transpose(1, 2)
    tok = torch.cat((sd["model.cls_token"].expand(B, -1, -1), tok), dim=1)
    if sd["model.pos_embed"].shape[1] == tok.shape[1]:          # dinomodel.py:112-114: only when the grids happen to match
        tok = tok + sd["model.pos_embed"]
    D = tok.shape[-1]

    def to_map(t):
        return t[:, 1:, :].transpose(1, 2).reshape(B, D, Hp, Wp)

    def reduce(i, m):
        return F.conv2d(m, sd[f"reduce_layers.{i}.weight"], sd[f"reduce_layers.{i}.bias"])
    results, red = [reduce(0, to_map(tok))], 1
    depth = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("model.blocks."))
    for i in range(depth):
        p = f"model.blocks.{i}."
        h = F.layer_norm(tok, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-6)
        qkv = F.linear(h, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]).reshape(B, -1, 3, heads, D // heads).permute(2, 0, 3, 1, 4)
        a = ((qkv[0] * (D // heads) ** -0.5) @ qkv[1].transpose(-2, -1)).softmax(-1) @ qkv[2]
        a = F.linear(a.transpose(1, 2).reshape(B, -1, D), sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
        tok = tok + sd[p + "ls1.gamma"] * a
        h = F.layer_norm(tok, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-6)
        h = F.linear(F.gelu(F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])), sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
        tok = tok + sd[p + "ls2.gamma"] * h
        if i in tap_blocks:
            results.append(reduce(red, to_map(tok)))
            red += 1
    results[-1] = F.interpolate(results[-1], size=final_size, mode="bilinear", align_corners=False)
    return results
