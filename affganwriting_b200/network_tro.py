"""Training / evaluation step composition (reference GAN_word/network_tro.py:17-177): all four update modes and `eval`.

Generator, discriminator and writer classifier are this package's classes.  The recogniser is a constructor argument
(`rec=`): `True` builds this package's native RecModel (affganwriting_b200.recognizer); any other module with the reference
RecModel's call shape `rec(img [B,1,H,W], label [B,T], img_width=...) -> logits [B, T-1, vocab]` (modules_tro.py:631-636) works
too - the reference's own RecModel drops in unchanged (tests/test_gpu_rec_step.py).  With
a recogniser the step is the reference's complete objective: `rec_update` (network_tro.py:39-48) and
l_total = w_dis l_dis + w_cla l_cla + w_l1 l_l1 + w_rec l_rec in `gen_update` (:57-103).  With rec=None the l_rec term is
absent (BASELINE.json configs[1] names the three convolutional models only) and a warning says so once.

Loss weights follow network_tro.py:10-13 (w_dis = w_cla = w_rec = 1, w_l1 = 0).
"""
import warnings

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .load_data import IMG_WIDTH, OUTPUT_MAX_LEN, vocab_size
from .loss_tro import crit, log_softmax, recon_criterion
from .modules_tro import DisModel, GenModel_FC, WriterClaModel

import os as _os
_MERGED_DIS_PASS = _os.environ.get("AFFGW_MERGED_DIS_PASS", "1") != "0"
_RELAXED_DIS_FWD = _os.environ.get("AFFGW_RELAXED_DIS_FWD", "1") != "0"
_DIS_STREAM_PRIO = int(_os.environ.get("AFFGW_DIS_PRIO", "0"))
_dis_streams = {}
_text_streams = {}


def _text_stream(dev):
    """The text encoder's stream of ConTranModel.side_text_encoder (one per device)."""
    st = _text_streams.get(dev.index)
    if st is None:
        st = _text_streams[dev.index] = torch.cuda.Stream(device=dev)
    return st


def _dis_stream(dev):
    """The second stream of `shared={"early": True}` (one per device)."""
    st = _dis_streams.get(dev.index)
    if st is None:
        st = _dis_streams[dev.index] = torch.cuda.Stream(device=dev, priority=_DIS_STREAM_PRIO)
    return st

w_dis = 1.
w_cla = 1.
w_l1 = 0.
w_rec = 1.


class ConTranModel(nn.Module):
    def __init__(self, num_writers, show_iter_num=500, oov=True, rec=None, device=None, encoder=None):
        super().__init__()
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.gen = GenModel_FC(OUTPUT_MAX_LEN, encoder=encoder).to(dev)
        self.cla = WriterClaModel(num_writers).to(dev)
        self.dis = DisModel().to(dev)
        if rec is True or rec == "native":
            from .modules_tro import RecModel
            rec = RecModel(pretrain=False)          # network_tro.py:23
        self.rec = rec.to(dev) if rec is not None else None
        if rec is None:
            warnings.warn("ConTranModel built without a recogniser: gen_update optimises w_dis*l_dis + w_cla*l_cla only and "
                          "rec_update is unavailable (pass rec=<RecModel> for the reference's full objective, network_tro.py:88-101)",
                          stacklevel=2)
        self.iter_num = 0
        self.show_iter_num = show_iter_num
        self.oov = oov
        self.device_ = dev
        self.side_text_encoder = False      # set by trainer.Trainer: text encoder beside the image encoder (_generate_pair)

    def _to(self, t):
        return t.to(self.device_, non_blocking=True)

    def _generate_pair(self, tr_img, label_xt, label_xt_swap):
        g = self.gen
        if self.side_text_encoder:
            # the text encoder (a chain of ~15 latency-bound launches per label) depends on nothing the image encoder computes:
            # both calls run on their own stream, forked in front of the image encoder, and join before `mix` (their
            # BatchNorm1d statistics advance label first, swapped label second, as in the reference)
            main = torch.cuda.current_stream()
            aux = _text_stream(tr_img.device)
            fork = torch.cuda.Event()
            fork.record(main)
            f_xss = g.enc_image(tr_img)
            f_xs = f_xss[-1]
            aux.wait_event(fork)
            with torch.cuda.stream(aux):
                f_xt, f_embed = g.enc_text(label_xt, f_xs.shape)
                f_xt_s, f_embed_s = g.enc_text(label_xt_swap, f_xs.shape)
            main.wait_stream(aux)
            xg = g.decode(g.mix(f_xss, f_embed), f_xss, f_embed, f_xt)
            xg_swap = g.decode(g.mix(f_xss, f_embed_s), f_xss, f_embed_s, f_xt_s)
            return xg, xg_swap
        f_xss = g.enc_image(tr_img)
        f_xs = f_xss[-1]
        f_xt, f_embed = g.enc_text(label_xt, f_xs.shape)
        xg = g.decode(g.mix(f_xss, f_embed), f_xss, f_embed, f_xt)
        f_xt_s, f_embed_s = g.enc_text(label_xt_swap, f_xs.shape)
        xg_swap = g.decode(g.mix(f_xss, f_embed_s), f_xss, f_embed_s, f_xt_s)
        return xg, xg_swap

    # The reference evaluates the discriminator / classifier twice per loss, on xg and on xg_swap (or on two real samples),
    # and averages the two mean-reduced losses (network_tro.py:67-72,110-127).  Neither network has a batch-coupled layer
    # (ActFirstResBlocks with norm="none"), so one pass over the concatenated 2B batch gives the same number:
    # (mean_B(a) + mean_B(b)) / 2 == mean_2B(cat(a, b)).  It halves the launch count of the latency-bound tiny-map layers.
    @staticmethod
    def _pair(a, b):
        return torch.cat([a, b], dim=0)

    def _dis_losses(self, s1, s2, xg, xg_swap):
        """Discriminator forward + backward of dis_update on two real and two (detached) generated samples."""
        if not _MERGED_DIS_PASS:                                # the reference's literal sequence (A/B measurements)
            l_real = self.dis.calc_dis_real_loss(self._pair(s1, s2))
            l_real.backward(retain_graph=True)
            l_fake = self.dis.calc_dis_fake_loss(self._pair(xg, xg_swap))
            l_fake.backward()
            return l_real + l_fake
        n_real = 2 * s1.shape[0]
        logits = self.dis(torch.cat([s1, s2, xg, xg_swap], dim=0))
        l_real = ops.bce_with_logits_const(logits[:n_real], 1.0)
        l_fake = ops.bce_with_logits_const(logits[n_real:], 0.0)
        l_dis = l_real + l_fake
        l_dis.backward()
        return l_dis

    def _recognise(self, img, label):
        """RecModel call of network_tro.py:43,88-89: every image spans the full width."""
        widths = torch.from_numpy(np.array([IMG_WIDTH] * img.shape[0]))
        return self.rec(img, label, img_width=widths)

    @staticmethod
    def _rec_loss(pred, label):
        target = label[:, 1:]                                     # remove <GO>  (network_tro.py:44,90-91)
        return crit(log_softmax(pred.reshape(-1, vocab_size)), target.reshape(-1)), target

    def forward(self, train_data_list, epoch, mode, cer_func=None, shared=None):
        """`shared` (optional, not in the reference): a dict the caller passes to BOTH `dis_update` and the `gen_update` that
        follows it.  The reference generates the fake pair twice per iteration - under no_grad in dis_update
        (network_tro.py:117-118) and again, from the same generator weights and the same batch, in gen_update (:59-63).  With
        `shared`, dis_update runs that forward once WITH its autograd graph (BatchNorm running statistics advanced twice), trains
        the discriminator on the detached images, and gen_update back-propagates through the kept graph: same losses, same
        gradients, one generator forward less.

        `shared = {"early": True}` keeps BOTH of the reference's generator forwards and only moves the second one: dis_update
        generates the fake pair under no_grad as always, then issues gen_update's own generator forward (with its autograd graph,
        kept in `shared["pair"]`) on the launching stream while the discriminator's forward + backward over [real | fake] runs
        beside it on a second stream - the discriminator pass is HBM-bound 16/32-channel work, the generator forward is
        tensor-bound, and neither reads what the other writes.  Streams join before dis_update returns.
        `shared = {"heads": True, ...}`: in gen_update the writer classifier's pass over the generated pair runs on the second
        stream beside the discriminator's.  `"share": False` switches the single-forward behaviour off for a dict that only
        carries the stream flags."""
        tr_domain, tr_wid, tr_idx, tr_img, tr_img_width, tr_label, img_xt, label_xt, label_xt_swap = train_data_list
        tr_wid, tr_img = self._to(tr_wid), self._to(tr_img)
        img_xt, label_xt, label_xt_swap = self._to(img_xt), self._to(label_xt), self._to(label_xt_swap)

        if mode == "rec_update":                                  # network_tro.py:39-48
            if self.rec is None:
                raise ValueError("rec_update needs a recogniser: ConTranModel(..., rec=<RecModel>)")
            img = tr_img[:, 0:1, :, :]
            lab = self._to(tr_label)[:, 0, :]
            pred = self._recognise(img, lab)
            l_rec_tr, target = self._rec_loss(pred, lab)
            if cer_func is not None:
                cer_func.add(pred, target)
            l_rec_tr.backward()
            return l_rec_tr

        if mode == "cla_update":                                  # network_tro.py:50-55
            # the reference marks this slice requires_grad_ (network_tro.py:51) but never reads its gradient: the input
            # gradient of the stem convolution is dead work and is not computed here
            img = tr_img[:, 0:1, :, :]
            l_cla_tr = self.cla(img, tr_wid)
            l_cla_tr.backward()
            return l_cla_tr

        if mode == "gen_update":                                  # network_tro.py:57-103 without the l_rec term
            self.iter_num += 1
            if shared is not None and "pair" in shared:
                xg, xg_swap = shared.pop("pair")                  # generated by the dis_update of this iteration
            else:
                xg, xg_swap = self._generate_pair(tr_img, label_xt, label_xt_swap)
            both, wid2 = self._pair(xg, xg_swap), self._pair(tr_wid, tr_wid)
            if shared is not None and shared.get("heads", False):
                # the two critics of the generated pair are independent of each other: the classifier's forward (and, because
                # autograd runs a node's backward on the stream of its forward, its backward) on the second stream
                main = torch.cuda.current_stream()
                aux = _dis_stream(both.device)
                aux.wait_stream(main)
                with torch.cuda.stream(aux):
                    l_cla = self.cla(both, wid2)
                l_dis = self.dis.calc_gen_loss(both)
                main.wait_stream(aux)
            else:
                l_dis = self.dis.calc_gen_loss(both)
                l_cla = self.cla(both, wid2)
            l_l1 = torch.zeros((), device=xg.device) if self.oov else recon_criterion(xg, img_xt)   # network_tro.py:82-85
            if self.rec is not None:                              # network_tro.py:87-97
                pred_xt = self._recognise(xg, label_xt)
                pred_xt_swap = self._recognise(xg_swap, label_xt_swap)
                l_rec_ori, tgt = self._rec_loss(pred_xt, label_xt)
                l_rec_swap, tgt_swap = self._rec_loss(pred_xt_swap, label_xt_swap)
                if cer_func is not None:
                    cer_func[0].add(pred_xt, tgt)
                    cer_func[1].add(pred_xt_swap, tgt_swap)
                l_rec = (l_rec_ori + l_rec_swap) / 2.
                l_total = w_dis * l_dis + w_cla * l_cla + w_l1 * l_l1 + w_rec * l_rec
            else:
                l_rec = torch.zeros((), device=xg.device)
                l_total = w_dis * l_dis + w_cla * l_cla + w_l1 * l_l1
            l_total.backward()
            return l_total, l_dis, l_cla, l_l1, l_rec

        if mode == "dis_update":                                  # network_tro.py:105-138
            # as above: network_tro.py:108-109 sets requires_grad_ on both real samples, nobody reads .grad
            # One discriminator pass over [real pair | fake pair] (4B images) instead of the reference's two passes with two
            # backward calls (network_tro.py:110-129): the discriminator has no batch-coupled layer and both losses are means
            # over their own half, so losses and accumulated gradients are the same numbers - half the launches of the
            # latency-bound small-map layers, one weight-gradient GEMM per layer over twice the positions.
            s1 = tr_img[:, 0:1, :, :]
            s2 = tr_img[:, 1:2, :, :]
            early = shared is not None and shared.get("early", False)
            if early:
                with torch.no_grad(), ops.relaxed_forward(_RELAXED_DIS_FWD):
                    xg, xg_swap = self._generate_pair(tr_img, label_xt, label_xt_swap)
                main = torch.cuda.current_stream()
                aux = _dis_stream(xg.device)
                aux.wait_stream(main)                               # the fake pair (and everything queued before it)
                with torch.cuda.stream(aux):                        # autograd runs these nodes' backward on `aux` as well
                    l_dis = self._dis_losses(s1, s2, xg, xg_swap)
                # gen_update's generator forward, same weights and batch (network_tro.py:59-63), after the first one: the
                # BatchNorm running statistics advance in the reference's order
                shared["pair"] = self._generate_pair(tr_img, label_xt, label_xt_swap)
                main.wait_stream(aux)
                return l_dis
            if shared is not None and shared.get("share", True):
                with ops.bn_updates_twice():
                    xg, xg_swap = self._generate_pair(tr_img, label_xt, label_xt_swap)
                shared["pair"] = (xg, xg_swap)
                xg, xg_swap = xg.detach(), xg_swap.detach()
            else:
                # this pair only feeds the discriminator: the layers that tolerate it run one tensor-core pass (ops.relaxed_forward)
                with torch.no_grad(), ops.relaxed_forward(_RELAXED_DIS_FWD):
                    xg, xg_swap = self._generate_pair(tr_img, label_xt, label_xt_swap)
            return self._dis_losses(s1, s2, xg, xg_swap)

        if mode == "eval":                                        # network_tro.py:140-177 (the PNG dump of :151 is the caller's)
            with torch.no_grad():
                xg, xg_swap = self._generate_pair(tr_img, label_xt, label_xt_swap)
                self.iter_num += 1
                both = self._pair(xg, xg_swap)
                l_dis = self.dis.calc_gen_loss(both)
                l_rec = torch.zeros((), device=xg.device)
                if self.rec is not None:
                    pred_xt = self._recognise(xg, label_xt)
                    pred_xt_swap = self._recognise(xg_swap, label_xt_swap)
                    l_a, tgt = self._rec_loss(pred_xt, label_xt)
                    l_b, tgt_swap = self._rec_loss(pred_xt_swap, label_xt_swap)
                    if cer_func is not None:
                        cer_func[0].add(pred_xt, tgt)
                        cer_func[1].add(pred_xt_swap, tgt_swap)
                    l_rec = (l_a + l_b) / 2.
                l_cla = self.cla(both, self._pair(tr_wid, tr_wid))
            return l_dis, l_cla, l_rec

        raise ValueError(f"unsupported mode {mode!r}")
