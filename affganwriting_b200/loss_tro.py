"""Losses / metric of the training step (reference GAN_word/loss_tro.py).

  recon_criterion   loss_tro.py:5-6      L1 reconstruction (w_l1 = 0 in network_tro.py:12; only used with oov=False)
  crit, log_softmax loss_tro.py:8-35     label-smoothed KL (reduction 'sum') on the recogniser's logits - here ONE fused
                                         libaffgw kernel over the raw logits (`affgw_label_smooth_kl_*`): `log_softmax` is
                                         the identity marker the reference's call shape `crit(log_softmax(x), y)` composes with
  CER               loss_tro.py:43-72    character error rate accumulator (host side, as in the reference)
"""
import torch

from . import ops
from .load_data import index2letter, num_tokens, tokens, vocab_size


def recon_criterion(predict, target):
    return torch.mean(torch.abs(predict.float() - target.float()))


class _Logits:
    """What `log_softmax(x)` returns: the logits, tagged, so that `crit` can fuse soft-max, smoothing and KL in one kernel."""

    def __init__(self, x):
        self.x = x


def log_softmax(x):
    return _Logits(x)


class LabelSmoothing(torch.nn.Module):
    def __init__(self, size, padding_idx, smoothing=0.0):
        super().__init__()
        self.size, self.padding_idx, self.smoothing = size, padding_idx, smoothing
        self.confidence = 1.0 - smoothing

    def forward(self, x, target):
        if isinstance(x, _Logits):
            logits = x.x
        else:                       # already log-probabilities (the reference's call shape with torch's LogSoftmax): log_softmax is
            logits = x              # idempotent, so the fused kernel gives the same value
        assert logits.size(1) == self.size
        return ops.label_smoothing_kl(logits, target, self.padding_idx, self.smoothing)


crit = LabelSmoothing(vocab_size, tokens["PAD_TOKEN"], 0.4)


def _distance(a, b):
    """Levenshtein distance (the reference imports the `Levenshtein` package, loss_tro.py:2)."""
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


class CER:
    def __init__(self):
        self.ed = 0
        self.len = 0

    def add(self, pred, gt):
        pred_label = torch.topk(pred, 1, dim=-1)[1].squeeze(-1).cpu().numpy()      # b, t, V -> b, t
        gt = gt.cpu().numpy()
        for i in range(pred_label.shape[0]):
            pred_text = [c for c in pred_label[i].tolist() if c >= num_tokens]
            gt_text = [c for c in gt[i].tolist() if c >= num_tokens]
            self.ed += _distance("".join(index2letter[c - num_tokens] for c in pred_text),
                                 "".join(index2letter[c - num_tokens] for c in gt_text))
            self.len += len(gt_text)

    def fin(self):
        return 100 * (self.ed / self.len)
