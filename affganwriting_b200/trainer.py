"""One training iteration as the reference's driver runs it (GAN_word/main_run.py:146-167, 275-278), minus the
recogniser update: cla_update -> dis_update -> gen_update, each followed by the data-parallel gradient exchange of
the sub-network that was just differentiated and by its Adam step.

torch.optim.Adam is used as-is (fused multi-tensor Adam is SURVEY.md §8(f).2, a "next" row).
"""
import torch

from .network_tro import ConTranModel
from .parallel import GradientReducer, broadcast_module


class Trainer:
    def __init__(self, num_writers=500, lr_gen=1e-4, lr_dis=1e-4, lr_cla=1e-5, device=None, skip_unused_wgrad=True,
                 bucket_bytes=None, encoder=None):
        self.model = ConTranModel(num_writers, oov=True, device=device, encoder=encoder)
        m = self.model
        # main_run.py:275-278: Adam over filter(requires_grad, parameters()) with default betas / eps
        self.cla_opt = torch.optim.Adam([p for p in m.cla.parameters() if p.requires_grad], lr=lr_cla)
        self.dis_opt = torch.optim.Adam([p for p in m.dis.parameters() if p.requires_grad], lr=lr_dis)
        self.gen_opt = torch.optim.Adam([p for p in m.gen.parameters() if p.requires_grad], lr=lr_gen)
        kw = {} if bucket_bytes is None else {"bucket_bytes": bucket_bytes}
        self.red = {"cla": GradientReducer(m.cla.parameters(), **kw), "dis": GradientReducer(m.dis.parameters(), **kw),
                    "gen": GradientReducer(m.gen.parameters(), **kw)}
        # the reference computes dis / cla weight gradients inside gen_update and throws them away at the next
        # zero_grad (main_run.py:148-163); skipping them changes nothing observable
        self.skip_unused_wgrad = skip_unused_wgrad
        broadcast_module(m)

    def train_step(self, batch, epoch=0):
        m = self.model
        self.cla_opt.zero_grad()
        l_cla = m(batch, epoch, "cla_update")
        self.red["cla"].reduce()
        self.cla_opt.step()

        self.dis_opt.zero_grad()
        l_dis = m(batch, epoch, "dis_update")
        self.red["dis"].reduce()
        self.dis_opt.step()

        self.gen_opt.zero_grad()
        frozen = []
        if self.skip_unused_wgrad:
            for p in list(m.dis.parameters()) + list(m.cla.parameters()):
                if p.requires_grad:
                    p.requires_grad_(False)
                    frozen.append(p)
        try:
            l_total, l_dis_g, l_cla_g, _, _ = m(batch, epoch, "gen_update")
        finally:
            for p in frozen:
                p.requires_grad_(True)
        self.red["gen"].reduce()
        self.gen_opt.step()
        return {"cla": l_cla.detach(), "dis": l_dis.detach(), "gen": l_total.detach(), "gen_dis": l_dis_g.detach(),
                "gen_cla": l_cla_g.detach()}
