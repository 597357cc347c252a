"""Adam as the reference's driver configures it (GAN_word/main_run.py:275-278: torch.optim.Adam over the parameters that
require gradients, default betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad), stepped by ONE libaffgw launch per call
(`affgw_adam_step`) instead of torch's per-dtype / per-device multi-tensor chunks.

Same observable behaviour as torch.optim.Adam for this configuration: parameters whose `.grad` is None are skipped and keep
their own step count (96 generator tensors never receive a gradient); state (`step`, `exp_avg`, `exp_avg_sq`)
is created lazily; `state_dict()` / `load_state_dict()` use torch's layout, so optimiser checkpoints are interchangeable.
`tests/test_gpu_blocks.py::test_adam_matches_torch` steps both side by side."""
import torch

from . import _lib as L  # noqa: N812
from .ops import weights_updated


class Adam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        params = list(params)
        if not params:
            raise ValueError("optimizer got an empty parameter list")
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1:
            raise ValueError("invalid Adam hyper-parameter")
        self.param_groups = [dict(params=params, lr=float(lr), betas=(float(betas[0]), float(betas[1])), eps=float(eps))]
        self.state = {}                 # parameter -> {"step": int, "exp_avg": tensor, "exp_avg_sq": tensor}
        self._tables = {}               # (pointer tuple) -> device tables

    # ------------------------------------------------------------------------------------------------ torch.optim API
    def zero_grad(self, set_to_none=True):
        for grp in self.param_groups:
            for p in grp["params"]:
                if p.grad is None:
                    continue
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.detach_()
                    p.grad.zero_()

    @torch.no_grad()
    def step(self, grad_scale=1.0):
        """One Adam step for every parameter that has a gradient; `grad_scale` multiplies the gradients first (the 1/world
        of a summed data-parallel exchange can ride here)."""
        for grp in self.param_groups:
            by_step = {}
            for p in grp["params"]:
                if p.grad is None:
                    continue
                L.require_cuda(p, p.grad)
                if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("affganwriting_b200.optim.Adam steps contiguous fp32 parameters")
                if not p.grad.is_contiguous():
                    p.grad = p.grad.contiguous()
                st = self.state.get(p)
                if st is None:
                    st = self.state[p] = {"step": 0, "exp_avg": torch.zeros_like(p, memory_format=torch.contiguous_format),
                                          "exp_avg_sq": torch.zeros_like(p, memory_format=torch.contiguous_format)}
                st["step"] += 1
                by_step.setdefault(st["step"], []).append(p)
            for step, ps in by_step.items():        # normally one group: every live tensor has been stepped equally often
                t = self._table(ps)
                L.call("affgw_adam_step", t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr(), t[4].data_ptr(),
                       t[5].data_ptr(), len(ps), grp["lr"], grp["betas"][0], grp["betas"][1], grp["eps"], int(step),
                       float(grad_scale), L.stream(), nbytes=28 * sum(p.numel() for p in ps))
                # the kernel writes through raw pointers, so autograd's version counters do not move: tell the packed
                # operand cache of the convolutions (ops._WeightCache) that these parameters changed
                weights_updated(ps)

    def _table(self, ps):
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in ps)
        hit = self._tables.get(key)
        if hit is not None:
            return hit
        sizes, offs, acc = [], [], 0
        for p in ps:
            sizes.append(p.numel())
            offs.append(acc)
            acc += p.numel()
        host = torch.tensor([[p.data_ptr() for p in ps], [p.grad.data_ptr() for p in ps],
                             [self.state[p]["exp_avg"].data_ptr() for p in ps],
                             [self.state[p]["exp_avg_sq"].data_ptr() for p in ps], sizes, offs], dtype=torch.int64).pin_memory()
        dev = host.to(ps[0].device, non_blocking=True)
        if len(self._tables) > 64:
            torch.cuda.synchronize()        # a side stream may still be reading the tables about to be freed
            self._tables.clear()
        self._tables[key] = tuple(dev[i] for i in range(6)) + (host,)       # keep the pinned source alive until the copy has run
        return self._tables[key]

    # ------------------------------------------------------------------------------------------------ checkpoints
    def state_dict(self):
        idx = {p: i for i, p in enumerate(self.param_groups[0]["params"])}
        state = {idx[p]: {"step": torch.tensor(float(s["step"])), "exp_avg": s["exp_avg"], "exp_avg_sq": s["exp_avg_sq"]}
                 for p, s in self.state.items()}
        g = self.param_groups[0]
        return {"state": state, "param_groups": [dict(lr=g["lr"], betas=g["betas"], eps=g["eps"], weight_decay=0, amsgrad=False,
                                                      params=list(range(len(g["params"]))))]}

    def load_state_dict(self, sd):
        params = self.param_groups[0]["params"]
        g = sd["param_groups"][0]
        if len(g["params"]) != len(params):
            raise ValueError("loaded state dict has a parameter group of a different size")
        if g.get("weight_decay", 0) or g.get("amsgrad", False):
            raise ValueError("affganwriting_b200.optim.Adam has no weight decay / amsgrad")
        self.param_groups[0].update(lr=float(g["lr"]), betas=tuple(float(b) for b in g["betas"]), eps=float(g["eps"]))
        self.state, self._tables = {}, {}
        for i, s in sd["state"].items():
            p = params[int(i)]
            self.state[p] = {"step": int(float(s["step"])),
                             "exp_avg": s["exp_avg"].to(device=p.device, dtype=torch.float32).contiguous().clone(),
                             "exp_avg_sq": s["exp_avg_sq"].to(device=p.device, dtype=torch.float32).contiguous().clone()}
