"""TEST INFRASTRUCTURE - CPU restatement of the reference's handwriting recogniser as the GAN step calls it
(the top-ranked "next" row; no CUDA path exists for it yet - see DESIGN.md §9).

    RecModel.forward          GAN_word/modules_tro.py:610-638
    Encoder (VGG19-BN + BiGRU) recognizer/models/encoder_vgg.py:665-735, vgg_tro_channel3.py:56-82,196-212
    locationAttention          recognizer/models/attention.py:105-160
    Decoder                    recognizer/models/decoder.py:9-57
    Seq2Seq (beam search)      recognizer/models/seq2seqnew2.py:13-181

Functional: weights come from a state_dict (keys under `seq2seq.encoder.` / `seq2seq.decoder.`).  Parity pinned against the
UNMODIFIED reference run in the build container (oracle/make_golden_rec.py, tests/golden/rec.npz, tests/test_oracle_golden.py).

Two properties of the reference that any replacement has to reproduce, and that this restatement makes explicit:

  * RecModel.forward forces `seq2seq.train()` (modules_tro.py:633): BatchNorm uses batch statistics, `Dropout2d(0.5)` after the
    VGG features and the 0.5 inter-layer dropout of both 2-layer GRUs are ACTIVE on every call.  Random numbers are drawn in this
    order: one Dropout2d mask [B, 512, 1, 1]-broadcast over the feature map, the encoder GRU's layer-0 output mask, then one
    decoder-GRU mask per decoder step in the order the beam search visits (sample, step, hypothesis).  With the same
    torch.manual_seed the restatement consumes the generator identically (it calls the same ATen dropout / GRU entry points),
    which is what makes a bit-exact pin possible; a CUDA implementation has to take these masks as INPUTS.
  * The beam search scores hypotheses with log(step_out + 1e-12) where step_out are the decoder's raw LOGITS, not
    probabilities (seq2seqnew2.py:126): negative logits give NaN scores, and torch.topk ranks NaN above every number.  On
    random-init weights 46 % of the logits are negative, i.e. EVERY step has more than beam_size NaN scores and the hypotheses
    kept are decided by the order in which ATen's CPU topk returns NaNs (e.g. topk([.3, nan, 1, nan, -2, nan, nan], 3) ->
    indices [3, 6, 5]) and by how list.sort treats NaN keys.  The tokens chosen feed the next step, so the returned logits
    depend on it: a device implementation either emulates that order or does the 3 x 55 selection with the same CPU calls.
"""
import numpy as np
import torch
import torch.nn.functional as F
from torch import _VF
from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence

from oracle.affgw_oracle import batch_norm

VGG19 = (64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512)   # cfg 'E', no last pool
HIDDEN, EMBED, LAYERS, P_DROP = 512, 60, 2, 0.5


def vgg19_bn_features(x, sd, prefix, training=True, stats=None):
    """vgg_tro_channel3.py:56-69 with cfg 'E' (:78): conv3x3 pad 1 + BatchNorm2d + ReLU, 2x2 max pools, H/16 x W/16 out."""
    i = 0
    for v in VGG19:
        if v == "M":
            x = F.max_pool2d(x, 2, 2)
            i += 1
        else:
            x = F.conv2d(x, sd[f"{prefix}{i}.weight"], sd[f"{prefix}{i}.bias"], padding=1)
            x = torch.relu(batch_norm(x, sd, f"{prefix}{i + 1}.", training, stats))
            i += 3
    return x


def _gru_weights(sd, prefix, bidirectional):
    out = []
    for layer in range(LAYERS):
        for suffix in (("", "_reverse") if bidirectional else ("",)):
            out += [sd[f"{prefix}weight_ih_l{layer}{suffix}"], sd[f"{prefix}weight_hh_l{layer}{suffix}"],
                    sd[f"{prefix}bias_ih_l{layer}{suffix}"], sd[f"{prefix}bias_hh_l{layer}{suffix}"]]
    return out


def encoder(img, img_width, sd, prefix="seq2seq.encoder.", image_width=216, training=True, stats=None):
    """encoder_vgg.py:705-735 with bgru=True, step=None, flip=False -> (enc_out [T, B, 512], hidden [2, B, 512])."""
    B = img.shape[0]
    out = vgg19_bn_features(img, sd, prefix + "layer.features.", training, stats)            # B, 512, H/16, W/16
    if training:
        out = F.dropout2d(out, P_DROP, True)
    out = out.permute(3, 0, 2, 1).reshape(-1, B, out.shape[2] * out.shape[1])               # width, batch, height * channels
    width = out.shape[0]
    src_len = (np.asarray(img_width) * (width / image_width) + 0.999).astype("int")
    packed = pack_padded_sequence(out, src_len.tolist(), batch_first=False)
    h0 = torch.zeros(2 * LAYERS, B, HIDDEN, dtype=out.dtype)
    data, hidden = _VF.gru(packed.data, packed.batch_sizes, h0, _gru_weights(sd, prefix + "rnn.", True), True, LAYERS,
                           P_DROP, training, True)
    output, _ = pad_packed_sequence(type(packed)(data, packed.batch_sizes, packed.sorted_indices, packed.unsorted_indices))
    output = output[:, :, :HIDDEN] + output[:, :, HIDDEN:]                                   # SUM_UP: forward + backward halves
    return output, hidden[[1, 3]]                                                            # the backward direction of each layer


def attention(hidden, enc_out, enc_len, prev_attn, sd, prefix):
    """attention.py:132-160 (locationAttention): softmax over the valid steps of
    out(tanh(W_e enc + W_h mean_layers(hidden) + W_p conv1d_7(prev_attn)))."""
    enc = enc_out.transpose(0, 1)                                                            # b, t, f
    h = hidden.permute(1, 2, 0)                                                              # b, f, layers
    mask = torch.full((h.shape[0], LAYERS, 1), 1 / LAYERS)
    h = torch.bmm(h, mask).permute(0, 2, 1)                                                  # b, 1, f
    h = F.linear(h, sd[prefix + "hidden_proj.weight"], sd[prefix + "hidden_proj.bias"])
    loc = F.conv1d(prev_attn.unsqueeze(1), sd[prefix + "conv1d.weight"], sd[prefix + "conv1d.bias"], padding=3).permute(0, 2, 1)
    loc = F.linear(loc, sd[prefix + "prev_attn_proj.weight"], sd[prefix + "prev_attn_proj.bias"])
    e = F.linear(enc, sd[prefix + "encoder_output_proj.weight"], sd[prefix + "encoder_output_proj.bias"])
    energy = F.linear(torch.tanh(e + h + loc), sd[prefix + "out.weight"], sd[prefix + "out.bias"]).squeeze(2)   # b, t
    w = torch.zeros(energy.shape)
    for i, le in enumerate(enc_len):
        w[i, :le] = torch.softmax(energy[i, :le], dim=0)
    return w.unsqueeze(2)


def decoder_step(in_char, hidden, enc_out, src_len, prev_attn, sd, prefix="seq2seq.decoder.", training=True):
    """decoder.py:31-57 (tradeoff None): attention -> context, embedding of the arg-max of in_char, 2-layer GRU step, logits."""
    width = enc_out.shape[0]
    src = np.asarray(src_len)
    enc_len = (src * (width / src[0]) + 0.999).astype("int")
    attn = attention(hidden, enc_out, enc_len, prev_attn, sd, prefix + "attention.")         # b, t, 1
    context = torch.bmm(enc_out.permute(1, 2, 0), attn).squeeze(2)                           # b, f
    top1 = in_char.topk(1)[1]
    emb = sd[prefix + "embedding.weight"][top1].squeeze(1)
    x = torch.cat((emb, context), 1).unsqueeze(0)
    out, hid = _VF.gru(x, hidden, _gru_weights(sd, prefix + "gru.", False), True, LAYERS, P_DROP, training, False, False)
    logits = F.linear(out.squeeze(0), sd[prefix + "out.weight"], sd[prefix + "out.bias"])
    return logits, hid, attn.squeeze(2)


def seq2seq_beam(img, tar, img_width, sd, output_max_len=12, vocab_size=55, beam_size=3, training=True, stats=None):
    """seq2seqnew2.py:64-160 with train=False, eos_id=None: one independent beam search per sample; returns the logits of the best
    hypothesis at every step, [T - 1, B, V]."""
    B = img.shape[0]
    steps = output_max_len - 1
    enc_out, enc_hidden = encoder(img, img_width, sd, training=training, stats=stats)
    enc_T = enc_out.shape[0]
    eye = torch.eye(vocab_size)
    best_outputs = torch.zeros(steps, B, vocab_size)
    widths = torch.as_tensor(np.asarray(img_width))
    for b in range(B):
        enc_b = enc_out[:, b:b + 1, :]
        beams = [dict(logp=0.0, tokens=[int(tar[b, 0])], hidden=enc_hidden[:, b:b + 1, :].contiguous(),
                      attn=torch.zeros(1, enc_T), dists=[])]
        for _ in range(steps):
            new = []
            for beam in beams:
                dec_in = eye.index_select(0, torch.tensor([beam["tokens"][-1]]))
                out, hid, attn = decoder_step(dec_in, beam["hidden"], enc_b, widths[b:b + 1].numpy(), beam["attn"], sd,
                                              training=training)
                logp = torch.log(out + 1e-12).squeeze(0)                                      # logits, not probabilities
                top_lp, top_id = torch.topk(logp, k=beam_size, dim=-1)
                for k in range(beam_size):
                    new.append(dict(logp=beam["logp"] + float(top_lp[k]), tokens=beam["tokens"] + [int(top_id[k])], hidden=hid,
                                    attn=attn, dists=beam["dists"] + [out.squeeze(0)]))
            new.sort(key=lambda z: z["logp"], reverse=True)                                  # NaN compares False: order kept
            beams = new[:beam_size]
        best = max(beams, key=lambda z: z["logp"])
        n = min(len(best["dists"]), steps)
        if n:
            best_outputs[:n, b, :] = torch.stack(best["dists"][:n], dim=0)
    return best_outputs


def rec_forward(img, label, sd, img_width=None, training=True, stats=None):
    """RecModel.forward (modules_tro.py:631-636): grey image replicated to 3 channels, beam size 3 -> [B, T - 1, V] logits."""
    if img_width is None:
        img_width = [img.shape[-1]] * img.shape[0]
    out = seq2seq_beam(torch.cat([img, img, img], dim=1), label, img_width, sd, training=training, stats=stats)
    return out.permute(1, 0, 2)


def label_smoothing_loss(pred, target, vocab_size=55, padding_idx=2, smoothing=0.4):
    """network_tro.py:44-45 / :92-93 with loss_tro.py:8-35: `crit(log_softmax(pred.reshape(-1, V)), target.reshape(-1))`,
    KLDivLoss(reduction='sum') against the smoothed one-hot (confidence 0.6, 0.4 / (V - 2) elsewhere, PAD column and PAD rows 0).
    pred: [B, T - 1, V] logits from rec_forward, target: [B, T - 1] labels without <GO>."""
    x = torch.log_softmax(pred.reshape(-1, vocab_size), dim=-1)
    t = target.reshape(-1)
    dist = torch.full_like(x, smoothing / (vocab_size - 2))
    dist.scatter_(1, t.unsqueeze(1), 1.0 - smoothing)
    dist[:, padding_idx] = 0
    dist[t == padding_idx] = 0.0
    return F.kl_div(x, dist, reduction="sum")


# ----------------------------------------------------------------------------------------------------------------------
# Mask-injectable form (what a device implementation can be compared with): the same network with the recurrent layers
# written out and every random draw exposed.  `masks` maps a draw's name to the KEEP-mask already divided by (1 - p)
#   "enc.drop2d"            [B, 512, 1, 1]     Dropout2d after the VGG features (one value per (sample, channel))
#   "enc.gru"               [T, B, 1024]       dropout on the encoder GRU's layer-0 output (both directions), before layer 1
#   "dec.gru.{b}.{t}.{h}"   [1, 1, 512]        same for the decoder GRU, beam-search visit (sample b, step t, hypothesis h)
# With masks=None the draws are made with torch's generator in exactly this order, which reproduces `_VF.gru` / F.dropout2d bit for
# bit under the same seed (checked in tests/test_oracle_golden.py), and the masks drawn are returned in `record` if given.
# ----------------------------------------------------------------------------------------------------------------------
def _gru_cell(x, h, w_ih, w_hh, b_ih, b_hh):
    gi, gh = F.linear(x, w_ih, b_ih), F.linear(h, w_hh, b_hh)
    i_r, i_z, i_n = gi.chunk(3, -1)
    h_r, h_z, h_n = gh.chunk(3, -1)
    r, z = torch.sigmoid(i_r + h_r), torch.sigmoid(i_z + h_z)
    n = torch.tanh(i_n + r * h_n)
    return (1 - z) * n + z * h          # written as torch does: n + z * (h - n) differs in the last bit only


def _draw(name, shape, dtype, masks, record):
    if masks is not None:
        m = masks[name]
        assert tuple(m.shape) == tuple(shape), (name, tuple(m.shape), tuple(shape))
    else:
        m = torch.empty(shape, dtype=dtype).bernoulli_(1 - P_DROP).div_(1 - P_DROP)
    if record is not None:
        record[name] = m
    return m


def gru_layers(x, h0, weights, bidirectional, training, name, masks=None, record=None):
    """nn.GRU with 2 layers on a [T, B, F] input of full-length sequences: returns (output [T, B, H * dirs], h_n [2 * dirs, B, H]);
    inter-layer dropout (p = 0.5) multiplies the layer-0 output by the draw `name`."""
    dirs = 2 if bidirectional else 1
    T = x.shape[0]
    finals = []
    for layer in range(LAYERS):
        outs = []
        for d in range(dirs):
            w = weights[4 * (layer * dirs + d): 4 * (layer * dirs + d) + 4]
            h = h0[layer * dirs + d]
            steps = range(T - 1, -1, -1) if d == 1 else range(T)
            seq = [None] * T
            for t in steps:
                h = _gru_cell(x[t], h, *w)
                seq[t] = h
            outs.append(torch.stack(seq, 0))
            finals.append(h)
        x = torch.cat(outs, -1)
        if layer == 0 and training:
            x = x * _draw(name, x.shape, x.dtype, masks, record)
    return x, torch.stack(finals, 0)


def rec_forward_explicit(img, label, sd, masks=None, record=None, training=True, stats=None, beam_size=3, vocab_size=55,
                         output_max_len=12):
    """rec_forward with every random draw exposed (see above); all images must span the full width (the GAN step passes
    img_width = IMG_WIDTH for every sample, network_tro.py:43,88-89)."""
    x = torch.cat([img, img, img], dim=1)
    B = x.shape[0]
    ep, dp = "seq2seq.encoder.", "seq2seq.decoder."
    f = vgg19_bn_features(x, sd, ep + "layer.features.", training, stats)
    if training:
        f = f * _draw("enc.drop2d", (B, f.shape[1], 1, 1), f.dtype, masks, record)
    f = f.permute(3, 0, 2, 1).reshape(-1, B, f.shape[2] * f.shape[1])
    enc_T = f.shape[0]
    out, hid = gru_layers(f, torch.zeros(2 * LAYERS, B, HIDDEN, dtype=f.dtype), _gru_weights(sd, ep + "rnn.", True), True,
                          training, "enc.gru", masks, record)
    enc_out, enc_hidden = out[:, :, :HIDDEN] + out[:, :, HIDDEN:], hid[[1, 3]]
    steps = output_max_len - 1
    eye = torch.eye(vocab_size, dtype=f.dtype)
    best = torch.zeros(steps, B, vocab_size, dtype=f.dtype)
    dec_w = _gru_weights(sd, dp + "gru.", False)
    for b in range(B):
        enc_b = enc_out[:, b:b + 1, :]
        beams = [dict(logp=0.0, tokens=[int(label[b, 0])], hidden=enc_hidden[:, b:b + 1, :].contiguous(),
                      attn=torch.zeros(1, enc_T, dtype=f.dtype), dists=[])]
        for t in range(steps):
            new = []
            for hyp, beam in enumerate(beams):
                attn = attention(beam["hidden"], enc_b, [enc_T], beam["attn"], sd, dp + "attention.")
                context = torch.bmm(enc_b.permute(1, 2, 0), attn).squeeze(2)
                emb = sd[dp + "embedding.weight"][beam["tokens"][-1]].unsqueeze(0)
                o, hnew = gru_layers(torch.cat((emb, context), 1).unsqueeze(0), beam["hidden"], dec_w, False, training,
                                     f"dec.gru.{b}.{t}.{hyp}", masks, record)
                logits = F.linear(o.squeeze(0), sd[dp + "out.weight"], sd[dp + "out.bias"])
                logp = torch.log(logits + 1e-12).squeeze(0)
                top_lp, top_id = torch.topk(logp, k=beam_size, dim=-1)
                for k in range(beam_size):
                    new.append(dict(logp=beam["logp"] + float(top_lp[k]), tokens=beam["tokens"] + [int(top_id[k])], hidden=hnew,
                                    attn=attn.squeeze(2), dists=beam["dists"] + [logits.squeeze(0)]))
            new.sort(key=lambda z: z["logp"], reverse=True)
            beams = new[:beam_size]
        top = max(beams, key=lambda z: z["logp"])
        best[:len(top["dists"]), b, :] = torch.stack(top["dists"], dim=0)
    return best.permute(1, 0, 2)


def rec_forward_batched(img, label, sd, masks, stats=None, beam_size=3, vocab_size=55, output_max_len=12):
    """The shape a device implementation takes (prototype, test infrastructure): per decoding step ONE batched decoder call
    over every live hypothesis of every sample (B rows at step 0, 3 B afterwards) - attention, context, embedding, 2-layer GRU
    step with the per-visit masks gathered row by row, logits - and then, per sample, the reference's own selection calls
    (torch.topk / list.sort / max on 3 x 55 scores) on the host.  Same results as rec_forward_explicit with the same masks
    (tests/test_oracle_golden.py), so only the arithmetic has to move to the device."""
    x = torch.cat([img, img, img], dim=1)
    B = x.shape[0]
    ep, dp = "seq2seq.encoder.", "seq2seq.decoder."
    f = vgg19_bn_features(x, sd, ep + "layer.features.", True, stats) * masks["enc.drop2d"]
    f = f.permute(3, 0, 2, 1).reshape(-1, B, f.shape[2] * f.shape[1])
    T = f.shape[0]
    out, hid = gru_layers(f, torch.zeros(2 * LAYERS, B, HIDDEN, dtype=f.dtype), _gru_weights(sd, ep + "rnn.", True), True, True,
                          "enc.gru", masks)
    enc_out, enc_hidden = out[:, :, :HIDDEN] + out[:, :, HIDDEN:], hid[[1, 3]]
    steps = output_max_len - 1
    dec_w = _gru_weights(sd, dp + "gru.", False)
    beams = [[dict(logp=0.0, tokens=[int(label[b, 0])], hidden=enc_hidden[:, b, :], attn=torch.zeros(T, dtype=f.dtype), dists=[])]
             for b in range(B)]
    for t in range(steps):
        rows = [(b, h) for b in range(B) for h in range(len(beams[b]))]
        hidden = torch.stack([beams[b][h]["hidden"] for b, h in rows], 1)                    # 2, N, 512
        prev = torch.stack([beams[b][h]["attn"] for b, h in rows], 0)                         # N, T
        enc = enc_out[:, [b for b, _ in rows], :]                                             # T, N, 512
        tok = torch.tensor([beams[b][h]["tokens"][-1] for b, h in rows])
        attn = attention(hidden, enc, [T] * len(rows), prev, sd, dp + "attention.")           # N, T, 1
        context = torch.bmm(enc.permute(1, 2, 0), attn).squeeze(2)
        xin = torch.cat((sd[dp + "embedding.weight"][tok], context), 1)
        drop = torch.cat([masks[f"dec.gru.{b}.{t}.{h}"] for b, h in rows], 1)                 # 1, N, 512
        h0 = _gru_cell(xin, hidden[0], *dec_w[0:4])
        h1 = _gru_cell(h0 * drop[0], hidden[1], *dec_w[4:8])
        logits = F.linear(h1, sd[dp + "out.weight"], sd[dp + "out.bias"])                     # N, V
        new = [[] for _ in range(B)]
        for r, (b, h) in enumerate(rows):                                                     # host side: the reference's own calls
            beam = beams[b][h]
            top_lp, top_id = torch.topk(torch.log(logits[r] + 1e-12), k=beam_size, dim=-1)
            for k in range(beam_size):
                new[b].append(dict(logp=beam["logp"] + float(top_lp[k]), tokens=beam["tokens"] + [int(top_id[k])],
                                   hidden=torch.stack((h0[r], h1[r]), 0), attn=attn[r, :, 0], dists=beam["dists"] + [logits[r]]))
        for b in range(B):
            new[b].sort(key=lambda z: z["logp"], reverse=True)
            beams[b] = new[b][:beam_size]
    best = torch.zeros(steps, B, vocab_size, dtype=f.dtype)
    for b in range(B):
        top = max(beams[b], key=lambda z: z["logp"])
        best[:, b, :] = torch.stack(top["dists"], dim=0)
    return best.permute(1, 0, 2)
