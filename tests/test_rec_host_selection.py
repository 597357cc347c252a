"""CPU check of the recogniser's host-side beam selection (affganwriting_b200.recognizer.select_hypotheses): ONE
topk(log(x + 1e-12)) over all hypothesis rows of a step must select exactly what the reference's per-hypothesis calls select
(recognizer/models/seq2seqnew2.py:118-149: `torch.topk(torch.log(x + 1e-12), k)` per hypothesis, `list.sort`, keep beam_size).
The reference scores raw logits, so NaN scores are the rule, not the exception: the comparison runs on NaN-heavy, NaN-free,
mixed and tie-heavy logits, over whole 11-step decodes."""
import math

import pytest
import torch

from affganwriting_b200.recognizer import select_hypotheses


def _per_hypothesis(host, beams, beam_size, t):
    """The literal sequence of calls (one topk per hypothesis row)."""
    parents, tokens, new_beams = [], [], []
    r = 0
    for b in range(len(beams)):
        cand = []
        for k in range(len(beams[b])):
            score, toks, _, dists = beams[b][k]
            top_lp, top_id = torch.topk(torch.log(host[r] + 1e-12), k=beam_size, dim=-1)
            for j in range(beam_size):
                cand.append((score + float(top_lp[j]), toks + [int(top_id[j])], r, dists + [(t, r)]))
            r += 1
        cand.sort(key=lambda z: z[0], reverse=True)
        nb = []
        for score, toks, parent, dists in cand[:beam_size]:
            nb.append((score, toks, len(parents), dists))
            parents.append(parent)
            tokens.append(toks[-1])
        new_beams.append(nb)
    return new_beams, parents, tokens


def _same_score(a, b):
    return (math.isnan(a) and math.isnan(b)) or a == b


def _logits(kind, rows, vocab, g):
    if kind == "normal":                    # ~half of the logits negative -> NaN after the log
        return torch.randn(rows, vocab, generator=g)
    if kind == "positive":                  # no NaN at all: ordinary beam search
        return torch.rand(rows, vocab, generator=g) * 4 + 1e-3
    if kind == "few_negative":              # 0-4 NaNs per row: rows with fewer NaNs than beam_size mix NaN and finite scores
        x = torch.rand(rows, vocab, generator=g) * 4 + 1e-3
        for r in range(rows):
            n = int(torch.randint(0, 5, (1,), generator=g))
            x[r, torch.randperm(vocab, generator=g)[:n]] *= -1
        return x
    x = torch.randint(-2, 4, (rows, vocab), generator=g).float() * 0.5      # "ties": 6 distinct values, zeros (log(1e-12)) and NaNs
    return x


@pytest.mark.parametrize("kind", ["normal", "positive", "few_negative", "ties"])
@pytest.mark.parametrize("batch", [1, 5, 64])
def test_batched_selection_equals_the_per_hypothesis_calls(kind, batch):
    g = torch.Generator().manual_seed(17 + batch)
    vocab, beam, steps = 55, 3, 11
    a = [[(0.0, [0], b, [])] for b in range(batch)]
    b_ = [[(0.0, [0], b, [])] for b in range(batch)]
    for t in range(steps):
        rows = sum(len(bm) for bm in a)
        assert rows == (batch if t == 0 else beam * batch)
        host = _logits(kind, rows, vocab, g)
        a, pa, ta = select_hypotheses(host, a, beam, t)
        b_, pb, tb = _per_hypothesis(host, b_, beam, t)
        assert pa == pb and ta == tb
        for bm_a, bm_b in zip(a, b_):
            assert len(bm_a) == len(bm_b) == beam
            for (sa, toka, rowa, da), (sb, tokb, rowb, db) in zip(bm_a, bm_b):
                assert _same_score(sa, sb) and toka == tokb and rowa == rowb and da == db
    # the final pick (seq2seqnew2.py:151) sees the same scores in the same order
    for bm_a, bm_b in zip(a, b_):
        assert max(bm_a, key=lambda z: z[0])[1] == max(bm_b, key=lambda z: z[0])[1]
