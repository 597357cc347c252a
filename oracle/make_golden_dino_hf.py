"""Second pin of the DINOv2 BACKBONE restatement (TEST INFRASTRUCTURE; container-only:  python -m oracle.make_golden_dino_hf).

The hub module the reference loads (`torch.hub.load(<dinov2 checkout>, 'dinov2_vitl14', source='local')`, dinomodel.py:36) is
absent, but an INDEPENDENT implementation of the same published architecture is installed: `transformers.Dinov2Model`
(transformers 5.5; it loads the official DINOv2 checkpoints through the key conversion of its convert_dinov2_to_hf.py, so its
arithmetic under that key map IS the hub model's).  This script

  1. regenerates the seeded state of tests/golden/dino_spec.json (the same tensors the reference-wrapper golden was made with),
  2. maps its `model.*` entries into a Dinov2Model by that published conversion read backwards (qkv -> query | key | value,
     attn.proj -> attention.output.dense, ls{1,2}.gamma -> layer_scale{1,2}.lambda1, ...),
  3. runs the HF patch embedding + HF layers on the wrapper's production input (64x216 -> replicate-padded 70x224, 80 patch
     tokens + cls, NO positional embedding: the wrapper's fallback skips it when the grids differ, dinomodel.py:103-117), applies
     the wrapper's 1x1 reducers and compares with (a) oracle.dino_oracle.dino_encoder, (b) StandInViT's blocks, (c) the committed
     golden of the UNMODIFIED reference wrapper (tests/golden/dino.npz) - closing the triangle wrapper / restatement / HF,
  4. runs the whole HF model on a 518x518 input, where the 37x37 grid matches `pos_embed` and both sides add it un-interpolated
     (the positional branch and the cls-first token order),
and writes the HF token maps (sub-sampled) to tests/golden/dino_hf.npz so the CPU suite re-checks the oracle without importing
transformers."""
import json
import os

import numpy as np
import torch
import torch.nn.functional as F

from oracle import affgw_oracle as O
from oracle import dino_oracle as DO
from oracle import weights as W

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
POS_STRIDE = 16      # the 518x518 case keeps every 16th patch token of the last tap


def seeded_state():
    meta = json.load(open(os.path.join(OUT, "dino_spec.json")))
    sd = W.make_state(meta["spec"])
    for k in meta["spec"]:
        if k.endswith(("norm1.weight", "norm2.weight", "norm.weight", ".gamma")):
            sd[k] = sd[k] + 1.0
    return meta, sd


def to_hf(sd, depth):
    """`model.*` (hub names) -> transformers.Dinov2Model names: convert_dinov2_to_hf.py's rename table read backwards."""
    m = lambda k: sd["model." + k]      # noqa: E731
    out = {"embeddings.cls_token": m("cls_token"), "embeddings.mask_token": m("mask_token"),
           "embeddings.position_embeddings": m("pos_embed"),
           "embeddings.patch_embeddings.projection.weight": m("patch_embed.proj.weight"),
           "embeddings.patch_embeddings.projection.bias": m("patch_embed.proj.bias"),
           "layernorm.weight": m("norm.weight"), "layernorm.bias": m("norm.bias")}
    for i in range(depth):
        s, d = f"blocks.{i}.", f"encoder.layer.{i}."
        D = m(s + "attn.qkv.weight").shape[1]
        for j, n in enumerate(("query", "key", "value")):
            out[d + f"attention.attention.{n}.weight"] = m(s + "attn.qkv.weight")[j * D:(j + 1) * D]
            out[d + f"attention.attention.{n}.bias"] = m(s + "attn.qkv.bias")[j * D:(j + 1) * D]
        for a, b in (("attn.proj", "attention.output.dense"), ("norm1", "norm1"), ("norm2", "norm2"), ("mlp.fc1", "mlp.fc1"),
                     ("mlp.fc2", "mlp.fc2")):
            out[d + b + ".weight"], out[d + b + ".bias"] = m(s + a + ".weight"), m(s + a + ".bias")
        out[d + "layer_scale1.lambda1"], out[d + "layer_scale2.lambda1"] = m(s + "ls1.gamma"), m(s + "ls2.gamma")
    return out


def x_channels(meta):
    return meta["spec"]["model.patch_embed.proj.weight"][1]


def hf_model(meta, sd):
    from transformers import Dinov2Config, Dinov2Model
    a = meta["arch"]
    grid = int(round((meta["spec"]["model.pos_embed"][1] - 1) ** 0.5))
    cfg = Dinov2Config(hidden_size=a["embed_dim"], num_hidden_layers=a["depth"], num_attention_heads=a["num_heads"], patch_size=14,
                       image_size=14 * grid, num_channels=x_channels(meta), mlp_ratio=4,
                       layer_norm_eps=1e-6, hidden_act="gelu", qkv_bias=True, use_swiglu_ffn=False, drop_path_rate=0.0,
                       hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    hf = Dinov2Model(cfg).eval()
    missing, unexpected = hf.load_state_dict(to_hf(sd, a["depth"]), strict=True)
    assert not missing and not unexpected
    return hf


def hf_layers(hf, tok):
    """Token states after every HF layer (each layer is transformers' own Dinov2Layer.forward)."""
    states = []
    for layer in hf.encoder.layer:
        tok = layer(tok)
        states.append(tok)
    return states


def token_map(t, Hp, Wp):
    return t[:, 1:, :].transpose(1, 2).reshape(t.shape[0], t.shape[2], Hp, Wp)


def rel(a, b):
    return float((a - b).abs().max() / max(1.0, float(b.abs().max())))


def main():
    meta, sd = seeded_state()
    taps, heads, depth = meta["taps"], meta["arch"]["num_heads"], meta["arch"]["depth"]
    hf = hf_model(meta, sd)
    vit = DO.StandInViT(**meta["arch"]).eval()
    vit.patch_embed.proj = torch.nn.Conv2d(x_channels(meta), meta["arch"]["embed_dim"], 14, 14)     # the wrapper's 50-plane stem
    vit.load_state_dict({k[len("model."):]: v for k, v in sd.items() if k.startswith("model.")})
    gold = np.load(os.path.join(OUT, "dino.npz"))
    report, out = [], {}

    def note(name, e, tol):
        report.append({"name": name, "max_abs": e, "tol": tol})
        print(f"{name}: {e:.2e} (tol {tol:g})")
        assert e <= tol, name

    with torch.no_grad():
        # ---- production input: 64x216 -> 70x224, 5x16 patch tokens + cls, no positional embedding (wrapper fallback)
        x = O.synthetic_batch(2, 50)["tr_img"]
        xp = F.pad(x, (0, 8, 0, 6), mode="replicate")
        tok = torch.cat((hf.embeddings.cls_token.expand(2, -1, -1), hf.embeddings.patch_embeddings(xp)), dim=1)
        states = [tok] + hf_layers(hf, tok)                        # states[i + 1] = after block i
        mine = DO.dino_encoder(x, sd, heads, taps)
        t = tok
        for i, blk in enumerate(vit.blocks):
            t = blk(t)
            note(f"dino_hf.standin_block{i}", rel(t, states[i + 1]), 1e-5)
        for r, src in enumerate([0] + [b + 1 for b in taps]):
            m = F.conv2d(token_map(states[src], 5, 16), sd[f"reduce_layers.{r}.weight"], sd[f"reduce_layers.{r}.bias"])
            if r == len(taps):
                m = F.interpolate(m, size=(8, 27), mode="bilinear", align_corners=False)
            note(f"dino_hf.oracle_result{r}", rel(mine[r], m), 1e-5)
            if f"result{r}" in gold.files:                          # the UNMODIFIED reference wrapper's own output
                note(f"dino_hf.reference_wrapper_result{r}", rel(torch.from_numpy(gold[f"result{r}"]), m), 1e-5)
            out[f"tokens{r}"] = states[src].numpy()
        # ---- 37x37 grid: positional embedding added on both sides, whole HF model
        g = torch.Generator().manual_seed(3)
        xb = torch.rand(1, x.shape[1], 518, 518, generator=g) * 2 - 1
        emb = hf.embeddings(xb)                                      # cls | patches, + position_embeddings (no interpolation)
        last = hf_layers(hf, emb)[taps[-1]]
        mine_b = DO.dino_encoder(xb, sd, heads, taps, final_size=(37, 37))
        m = F.conv2d(token_map(last, 37, 37), sd[f"reduce_layers.{len(taps)}.weight"], sd[f"reduce_layers.{len(taps)}.bias"])
        note("dino_hf.oracle_pos_embed_result4", rel(mine_b[-1], m), 1e-5)
        m0 = F.conv2d(token_map(emb, 37, 37), sd["reduce_layers.0.weight"], sd["reduce_layers.0.bias"])
        note("dino_hf.oracle_pos_embed_result0", rel(mine_b[0], m0), 1e-5)
        out["pos_tokens_last"] = last[:, 1::POS_STRIDE, :].numpy()
        out["pos_tokens_stem"] = emb[:, 1::POS_STRIDE, :].numpy()
    import transformers
    np.savez_compressed(os.path.join(OUT, "dino_hf.npz"), **out)
    json.dump({"transformers": transformers.__version__, "pos_stride": POS_STRIDE, "pos_seed": 3, "report": report},
              open(os.path.join(OUT, "dino_hf_report.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
