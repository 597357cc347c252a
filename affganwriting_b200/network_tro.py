"""Training / evaluation step composition (reference GAN_word/network_tro.py:17-177) for the sub-networks on the
accelerated path: generator, discriminator, writer classifier.  The recogniser (`rec_update`, the l_rec term) is a
GRU seq2seq with host-side beam search and is out of scope; an optional `rec` module supplied
by the caller is used unchanged.

Loss weights follow network_tro.py:10-13 (w_dis = w_cla = w_rec = 1, w_l1 = 0).
"""
import torch
import torch.nn as nn

from .load_data import OUTPUT_MAX_LEN
from .modules_tro import DisModel, GenModel_FC, WriterClaModel

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
        if rec is not None:
            self.rec = rec
        self.iter_num = 0
        self.show_iter_num = show_iter_num
        self.oov = oov
        self.device_ = dev

    def _to(self, t):
        return t.to(self.device_, non_blocking=True)

    def _generate_pair(self, tr_img, label_xt, label_xt_swap):
        g = self.gen
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

    def forward(self, train_data_list, epoch, mode, cer_func=None):
        tr_domain, tr_wid, tr_idx, tr_img, tr_img_width, tr_label, img_xt, label_xt, label_xt_swap = train_data_list
        tr_wid, tr_img = self._to(tr_wid), self._to(tr_img)
        img_xt, label_xt, label_xt_swap = self._to(img_xt), self._to(label_xt), self._to(label_xt_swap)

        if mode == "cla_update":                                  # network_tro.py:50-55
            # the reference marks this slice requires_grad_ (network_tro.py:51) but never reads its gradient: the input
            # gradient of the stem convolution is dead work and is not computed here
            img = tr_img[:, 0:1, :, :]
            l_cla_tr = self.cla(img, tr_wid)
            l_cla_tr.backward()
            return l_cla_tr

        if mode == "gen_update":                                  # network_tro.py:57-103 without the l_rec term
            self.iter_num += 1
            xg, xg_swap = self._generate_pair(tr_img, label_xt, label_xt_swap)
            both, wid2 = self._pair(xg, xg_swap), self._pair(tr_wid, tr_wid)
            l_dis = self.dis.calc_gen_loss(both)
            l_cla = self.cla(both, wid2)
            l_l1 = torch.zeros((), device=xg.device)
            l_rec = torch.zeros((), device=xg.device)
            if getattr(self, "rec", None) is not None and cer_func is not None:
                raise NotImplementedError("recogniser term: pass rec=None (SURVEY.md §8(f).1)")
            l_total = w_dis * l_dis + w_cla * l_cla
            l_total.backward()
            return l_total, l_dis, l_cla, l_l1, l_rec

        if mode == "dis_update":                                  # network_tro.py:105-138
            # as above: network_tro.py:108-109 sets requires_grad_ on both real samples, nobody reads .grad
            s1 = tr_img[:, 0:1, :, :]
            s2 = tr_img[:, 1:2, :, :]
            l_real = self.dis.calc_dis_real_loss(self._pair(s1, s2))
            l_real.backward(retain_graph=True)
            with torch.no_grad():
                xg, xg_swap = self._generate_pair(tr_img, label_xt, label_xt_swap)
            l_fake = self.dis.calc_dis_fake_loss(self._pair(xg, xg_swap))
            l_fake.backward()
            return l_real + l_fake

        if mode == "eval":                                        # network_tro.py:140-177 without rec / image dump
            with torch.no_grad():
                xg, xg_swap = self._generate_pair(tr_img, label_xt, label_xt_swap)
                self.iter_num += 1
                both = self._pair(xg, xg_swap)
                l_dis = self.dis.calc_gen_loss(both)
                l_cla = self.cla(both, self._pair(tr_wid, tr_wid))
            return l_dis, l_cla, torch.zeros((), device=xg.device)

        raise ValueError(f"unsupported mode {mode!r} (rec_update needs the out-of-scope recogniser)")
