"""Handwriting recogniser of the GAN step, native.  Mirrors the reference's

    RecModel                   GAN_word/modules_tro.py:610-638
    Encoder (VGG19-BN + BiGRU) recognizer/models/encoder_vgg.py:669-735, vgg_tro_channel3.py:56-82
    locationAttention          recognizer/models/attention.py:105-160
    Decoder                    recognizer/models/decoder.py:9-57
    Seq2Seq (beam search)      recognizer/models/seq2seqnew2.py:13-181

with the same module tree (so `state_dict` keys - `enc.*`, `dec.*`, `seq2seq.encoder.*`, `seq2seq.decoder.*` - and default
initialisation are the reference's; torch.nn modules are parameter containers only) and every forward on libaffgw kernels:
the 16 convolutions, all input / hidden / attention / output projections on the tcgen05 GEMM kernels, BatchNorm / max-pool on
the streaming kernels, GRU cells and location attention on `rec.cu`.

How the decode is organised (the reference loops over samples, beams and steps in Python with a `.item()` per candidate,
seq2seqnew2.py:87-139):
  * per decoding step ONE batched pass over every live hypothesis of every sample (B rows at step 0, 3 B afterwards);
  * the scores of a step ([rows, 55] logits) are copied to the host once and the hypotheses are selected there with the
    reference's own calls (`torch.topk(torch.log(x + 1e-12))`, `list.sort`, `max`): the reference scores raw logits, so
    negative logits give NaN scores and the surviving hypotheses depend on how those calls order NaNs
    (measured with the CPU oracle) - running the same calls on the same numbers is what keeps the tokens identical;
  * the states of the selected parents are gathered on the device for the next step.

`RecModel.forward` always runs in training mode like the reference (modules_tro.py:633 forces `seq2seq.train()`): BatchNorm
uses batch statistics and the three dropouts are active.  Their keep-masks can be injected (`masks=`, the layout of
CPU oracle's mask-injectable restatement) so that a run can be compared with the CPU oracle; otherwise they are drawn on the
device with torch's generator.
"""
import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import ops, rec_ops
from .load_data import IMG_HEIGHT, IMG_WIDTH, OUTPUT_MAX_LEN, vocab_size

_VGG19 = (64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512)   # cfg 'E', no last pool
P_DROP = 0.5


def _keep_mask(shape, device):
    return torch.empty(shape, dtype=torch.float32, device=device).bernoulli_(1 - P_DROP).div_(1 - P_DROP)


class _VGG(nn.Module):
    """vgg_tro_channel3.py:25-82 (`vgg19_bn`, 3 input planes, BatchNorm): container + initialisation."""

    def __init__(self):
        super().__init__()
        layers, cin = [], 3
        for v in _VGG19:
            if v == "M":
                layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
            else:
                layers += [nn.Conv2d(cin, v, kernel_size=3, padding=1), nn.BatchNorm2d(v), nn.ReLU()]
                cin = v
        self.features = nn.Sequential(*layers)
        for m in self.modules():                       # vgg_tro_channel3.py:38-50
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out")
                nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def run(self, x):
        """x: [B, 1, H, W] grey image; the reference feeds cat([img, img, img]) (modules_tro.py:634), i.e. the first
        convolution sees three identical planes: its weight is summed over the input planes instead."""
        feats = self.features
        i, first = 0, True
        while i < len(feats):
            m = feats[i]
            if isinstance(m, nn.Conv2d):
                w = m.weight.sum(dim=1, keepdim=True) if first else m.weight
                first = False
                x = ops.conv2d(x, w, m.bias, stride=1, pad=1, pad_mode="zero")
                x = ops.batch_norm(x, feats[i + 1], act="relu")
                i += 3
            else:
                x = ops.max_pool2(x)
                i += 1
        return x


class Encoder(nn.Module):
    def __init__(self, hidden_size, height, width, bgru, step, flip):
        super().__init__()
        assert bgru and step is None and not flip, "the GAN step builds Encoder(512, 64, 216, True, None, False)"
        self.hidden_size, self.height, self.width = hidden_size, height, width
        self.n_layers, self.dropout = 2, P_DROP
        self.layer = _VGG()
        self.layer_dropout = nn.Dropout2d(p=P_DROP)
        self.rnn = nn.GRU(self.height // 16 * 512, self.hidden_size, self.n_layers, dropout=self.dropout, bidirectional=True)

    def _direction(self, x_seq, layer, reverse, h0):
        sfx = f"_l{layer}" + ("_reverse" if reverse else "")
        w_ih, w_hh = getattr(self.rnn, "weight_ih" + sfx), getattr(self.rnn, "weight_hh" + sfx)
        b_ih, b_hh = getattr(self.rnn, "bias_ih" + sfx), getattr(self.rnn, "bias_hh" + sfx)
        T, B, Fin = x_seq.shape
        gi = ops.linear(x_seq.reshape(T * B, Fin), w_ih, b_ih).view(T, B, -1).unbind(0)     # one GEMM for all time steps
        h, outs = h0, [None] * T
        for t in (range(T - 1, -1, -1) if reverse else range(T)):
            h = rec_ops.gru_cell(gi[t], ops.linear(h, w_hh, b_hh), h)
            outs[t] = h
        return torch.stack(outs, 0), h

    def forward(self, img, masks=None):
        """-> (enc_out [T, B, 512], hidden [2, B, 512]); every image spans the full width (network_tro.py:43,88-89)."""
        x = ops.input_to_internal(img)
        B = x.shape[0]
        f = self.layer.run(x)                                                    # B, 512, H/16, W/16
        m2d = masks["enc.drop2d"].to(f.device) if masks is not None else _keep_mask((B, f.shape[1]), f.device)
        f = rec_ops.scale_nc(f, m2d.reshape(B, -1))
        seq = rec_ops.map_to_seq(f)                                              # T, B, H/16 * 512
        T = seq.shape[0]
        h0 = torch.zeros(B, self.hidden_size, dtype=torch.float32, device=seq.device)
        of, _ = self._direction(seq, 0, False, h0)
        ob, hb0 = self._direction(seq, 0, True, h0)
        x1 = torch.cat((of, ob), dim=-1)                                         # T, B, 1024
        mg = masks["enc.gru"].to(x1.device) if masks is not None else _keep_mask(x1.shape, x1.device)
        x1 = rec_ops.mul_mask(x1, mg)
        of1, _ = self._direction(x1, 1, False, h0)
        ob1, hb1 = self._direction(x1, 1, True, h0)
        enc_out = ops.add(of1.reshape(T * B, -1), ob1.reshape(T * B, -1)).view(T, B, -1)     # SUM_UP (encoder_vgg.py:727-728)
        return enc_out, torch.stack((hb0, hb1), 0)                               # hidden[[1, 3]]: the backward directions


class locationAttention(nn.Module):
    def __init__(self, hidden_size, decoder_layer):
        super().__init__()
        self.hidden_size, self.decoder_layer = hidden_size, decoder_layer
        self.proj = nn.Linear(hidden_size, hidden_size)                          # unused by forward (attention.py:114), kept for the keys
        self.tanh = nn.Tanh()
        self.hidden_proj = nn.Linear(hidden_size, hidden_size)
        self.encoder_output_proj = nn.Linear(hidden_size, hidden_size)
        self.out = nn.Linear(hidden_size, 1)
        self.conv1d = nn.Conv1d(1, 128, 7, padding=3)
        self.prev_attn_proj = nn.Linear(128, hidden_size)
        self.softmax = nn.Softmax(dim=0)
        self.sigmoid = nn.Sigmoid()

    def location_filter(self):
        """conv1d(1 -> 128, k = 7) followed by Linear(128 -> F) is one linear map of the 7-wide window of the previous
        attention: weight [F, 7], bias [F] (attention.py:151-154)."""
        w = self.prev_attn_proj.weight @ self.conv1d.weight[:, 0, :]
        b = self.prev_attn_proj.weight @ self.conv1d.bias + self.prev_attn_proj.bias
        return w, b

    def forward(self, h0, h1, e_proj, enc_bt, sample, prev_attn, loc_w, loc_b):
        n, t = prev_attn.shape
        hp = ops.linear(ops.add(h0, h1), self.hidden_proj.weight * (1.0 / self.decoder_layer), self.hidden_proj.bias)
        win = F.pad(prev_attn, (3, 3)).unfold(1, 7, 1).reshape(n * t, 7)
        loc = ops.linear(win, loc_w, loc_b).view(n, t, -1)
        energy = rec_ops.attn_energy(e_proj, sample, hp, loc, self.out.weight.reshape(-1), self.out.bias)
        return rec_ops.attn_softmax_context(energy, enc_bt, sample)              # (attn [n, t], context [n, F])


class Decoder(nn.Module):
    def __init__(self, hidden_size, embedding_size, vocab, attention, tradeoff_context_embed):
        super().__init__()
        assert tradeoff_context_embed is None
        self.hidden_size, self.embed_size, self.n_layers, self.dropout = hidden_size, embedding_size, 2, P_DROP
        self.embedding = nn.Embedding(vocab, self.embed_size)
        self.attention = attention(self.hidden_size, self.n_layers)
        self.gru = nn.GRU(self.embed_size + self.hidden_size, self.hidden_size, self.n_layers, dropout=self.dropout)
        self.out = nn.Linear(self.hidden_size, vocab)

    def step(self, tok, h0, h1, e_proj, enc_bt, sample, prev_attn, loc_w, loc_b, drop):
        """One decoder step for `n` hypothesis rows (decoder.py:31-57): -> logits [n, V], new h0, new h1, attention [n, T]."""
        attn, context = self.attention(h0, h1, e_proj, enc_bt, sample, prev_attn, loc_w, loc_b)
        x = torch.cat((ops.embedding(tok, self.embedding.weight), context), 1)
        g = self.gru
        n0 = rec_ops.gru_cell(ops.linear(x, g.weight_ih_l0, g.bias_ih_l0), ops.linear(h0, g.weight_hh_l0, g.bias_hh_l0), h0)
        n1 = rec_ops.gru_cell(ops.linear(rec_ops.mul_mask(n0, drop), g.weight_ih_l1, g.bias_ih_l1),
                              ops.linear(h1, g.weight_hh_l1, g.bias_hh_l1), h1)
        return ops.linear(n1, self.out.weight, self.out.bias), n0, n1, attn


def select_hypotheses(host, beams, beam_size, t):
    """Host side of one beam-search step (seq2seqnew2.py:118-149) for every sample at once.  `host` [rows, V]: the step's
    logits, rows ordered sample by sample, hypothesis by hypothesis; `beams[b]`: the live hypotheses of sample b as
    (score, tokens, row, [(step, row)] of the logits along its path).  -> (new beams, parent row of every kept hypothesis,
    its new token).
    The reference scores `torch.topk(torch.log(x + 1e-12), k)` per hypothesis (:126-127), sorts each sample's candidates with
    `list.sort` (:141) and keeps `beam_size`; negative logits give NaN scores, so WHICH calls are made matters.  topk runs
    the same per-row routine on a [rows, V] tensor as on each row alone (NaN ordering included) and log is elementwise, so ONE
    call over all rows returns what the per-hypothesis calls return (tests/test_rec_host_selection.py) - at 3 B = 192 rows
    that is 0.6 ms of host time per step instead of 9."""
    lp, idx = torch.topk(torch.log(host + 1e-12), k=beam_size, dim=-1)
    lp, idx = lp.tolist(), idx.tolist()
    parents, tokens, new_beams = [], [], []
    r = 0
    for bm in beams:
        cand = []
        for score, toks, _, dists in bm:
            path = dists + [(t, r)]
            for j in range(beam_size):
                cand.append((score + lp[r][j], toks + [idx[r][j]], r, path))
            r += 1
        cand.sort(key=lambda z: z[0], reverse=True)                              # NaN keys compare False: order kept (:141)
        nb = []
        for score, toks, parent, dists in cand[:beam_size]:
            nb.append((score, toks, len(parents), dists))
            parents.append(parent)
            tokens.append(toks[-1])
        new_beams.append(nb)
    return new_beams, parents, tokens


class Seq2Seq(nn.Module):
    def __init__(self, encoder, decoder, output_max_len, vocab):
        super().__init__()
        self.encoder, self.decoder = encoder, decoder
        self.output_max_len, self.vocab_size = output_max_len, vocab

    def forward(self, src, tar, src_len=None, teacher_rate=False, train=False, beam_size=3, masks=None):
        """seq2seqnew2.py:64-160 with train=False, eos_id=None: per-sample beam search, batched over samples and hypotheses;
        returns the logits of each sample's best hypothesis at every step, [T - 1, B, V]."""
        dev = src.device
        B, steps, V = src.shape[0], self.output_max_len - 1, self.vocab_size
        enc_out, hid = self.encoder(src, masks)                                  # [T, B, F], [2, B, F]
        T = enc_out.shape[0]
        enc_bt = enc_out.transpose(0, 1).contiguous()                            # B, T, F
        att = self.decoder.attention
        e_proj = ops.linear(enc_bt.reshape(B * T, -1), att.encoder_output_proj.weight, att.encoder_output_proj.bias).view(B, T, -1)
        loc_w, loc_b = att.location_filter()
        go = tar[:, 0].tolist()
        # host-side beam bookkeeping: per sample a list of (score, tokens, row in the current step, [(step, row)] of its logits)
        beams = [[(0.0, [int(go[b])], b, [])] for b in range(B)]
        sample = torch.arange(B, device=dev)
        h0, h1 = hid[0], hid[1]
        prev = torch.zeros(B, T, dtype=torch.float32, device=dev)
        tok = tar[:, 0].to(dev).contiguous()
        step_logits = []
        for t in range(steps):
            rows = [(b, k) for b in range(B) for k in range(len(beams[b]))]
            n = len(rows)
            if masks is not None:
                drop = torch.cat([masks[f"dec.gru.{b}.{t}.{k}"].reshape(1, -1) for b, k in rows], 0).to(dev)
            else:
                drop = _keep_mask((n, self.decoder.hidden_size), dev)
            logits, n0, n1, attn = self.decoder.step(tok, h0, h1, e_proj, enc_bt, sample, prev, loc_w, loc_b, drop)
            step_logits.append(logits)
            host = logits.detach().float().cpu()                                 # the one host synchronisation of the step
            new_beams, parents, tokens = select_hypotheses(host, beams, beam_size, t)
            beams = new_beams
            if t + 1 < steps:
                pidx = torch.tensor(parents, dtype=torch.int64, device=dev)
                h0, h1, prev = n0.index_select(0, pidx), n1.index_select(0, pidx), attn.index_select(0, pidx)
                tok = torch.tensor(tokens, dtype=torch.int64, device=dev)
                sample = torch.tensor([b for b in range(B) for _ in range(len(beams[b]))], dtype=torch.int64, device=dev)
        best = [max(bm, key=lambda z: z[0]) for bm in beams]                     # seq2seqnew2.py:151
        out = []
        for t in range(steps):
            idx = torch.tensor([best[b][3][t][1] for b in range(B)], dtype=torch.int64, device=dev)
            out.append(step_logits[t].index_select(0, idx))
        return torch.stack(out, 0), None


class RecModel(nn.Module):
    def __init__(self, pretrain=False):
        super().__init__()
        hidden_size_enc = hidden_size_dec = 512
        embed_size = 60
        self.enc = Encoder(hidden_size_enc, IMG_HEIGHT, IMG_WIDTH, True, None, False)
        self.dec = Decoder(hidden_size_dec, embed_size, vocab_size, locationAttention, None)
        self.seq2seq = Seq2Seq(self.enc, self.dec, OUTPUT_MAX_LEN, vocab_size)
        if pretrain:
            raise RuntimeError("pre-trained recogniser weights are not shipped; load a checkpoint with load_state_dict")

    def forward(self, img, label, img_width=None, masks=None):
        """img [B, 1, H, W], label [B, T] (only the <GO> column is read) -> logits [B, T - 1, vocab]  (modules_tro.py:631-636)."""
        self.seq2seq.train()
        if img_width is not None:
            w = np.asarray(img_width)
            if (w != img.shape[-1]).any():
                raise RuntimeError("RecModel: every image must span the full width (the GAN step passes IMG_WIDTH for all samples)")
        output, _ = self.seq2seq(img, label, img_width, teacher_rate=False, train=False, beam_size=3, masks=masks)
        return output.permute(1, 0, 2)
