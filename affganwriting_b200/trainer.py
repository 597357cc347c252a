"""One training iteration as the reference's driver runs it (GAN_word/main_run.py:146-167, 275-278):
[rec_update ->] cla_update -> dis_update -> gen_update, each followed by the data-parallel gradient exchange of
the sub-network that was just differentiated and by its Adam step.  The recogniser sub-step (and the l_rec term of
gen_update) runs when a recogniser module is supplied (`rec=`, see network_tro.ConTranModel); it is stepped by
torch.optim.Adam at lr 1e-5 like main_run.py:277.  A recogniser decodes with a host-side beam search, so the sub-steps that
call it (rec_update, gen_update) are issued eagerly even in CUDA-graph mode; cla_update and dis_update are still replayed.

The optimiser is `optim.Adam`: torch.optim.Adam's update for the reference's configuration as ONE libaffgw launch per
sub-network; AFFGW_ADAM=torch selects torch.optim.Adam(fused=True) instead.

CUDA graphs.  One iteration is ~5000 small launches issued from Python (~27 us each, ~140 ms per iteration on the host),
which is of the same order as the GPU time of the kernels themselves.  With `cuda_graph=True` the forward + backward of
each of the three sub-steps is captured once (after a few eager iterations that run every lazy initialisation) and
replayed from then on; the gradient exchange (NCCL) and the optimiser steps stay eager between the replays, so nothing
that talks to another rank is ever inside a graph.  The batch is copied into static device tensors before each replay.

Overlapped exchange (`overlap_exchange=True`, graph mode only).  The three sub-networks have separate parameters and
cla_update touches only the classifier, dis_update only generator (forward) + discriminator, gen_update all three.  So the
gradient exchange + Adam of the GENERATOR (the largest: ~160 MB of gradients) can run on a side stream while the next
iteration's cla_update graph replays, and the classifier's while dis_update replays; the discriminator's cannot (gen_update
needs its new weights at once).  Stream order: the side stream waits for the graph that produced the gradients; the main
stream waits for the step's event before the first graph that reads the updated weights (dis_update after the generator's
step, gen_update after the classifier's) - the same points where the next replay would overwrite the gradient buffers.
Each graph then has its OWN memory pool: with a shared pool the classifier graph's scratch buffers alias the generator
graph's gradient buffers, which the side stream is still reading.  `join()` makes the main stream wait for everything
pending; train_step() leaves the generator's step in flight, so call join() before reading weights, evaluating,
checkpointing or switching to another path.
"""
import os

import torch

from . import ops
from .network_tro import ConTranModel
from .optim import Adam
from .parallel import GradientReducer, broadcast_module


class Trainer:
    GRAPH_WARMUP = 3      # eager iterations before the capture

    def __init__(self, num_writers=500, lr_gen=1e-4, lr_dis=1e-4, lr_cla=1e-5, device=None, skip_unused_wgrad=True,
                 bucket_bytes=None, encoder=None, cuda_graph=False, overlap_exchange=False, rec=None, lr_rec=1e-5,
                 wgrad_stream=True, concurrent_cla_dis=True, share_generator_forward=False, early_generator_forward=None,
                 concurrent_gen_heads=None):
        import warnings
        with warnings.catch_warnings():
            if rec is None:
                warnings.simplefilter("ignore")     # configs[1] of BASELINE.json: the three convolutional models, by design
            self.model = ConTranModel(num_writers, oov=True, device=device, encoder=encoder, rec=rec)
        m = self.model
        # main_run.py:275-278: Adam over filter(requires_grad, parameters()) with default betas / eps
        # (fused=True is torch's single-pass multi-tensor implementation of the same update)
        import os
        if os.environ.get("AFFGW_ADAM", "affgw") == "torch":
            fused = os.environ.get("AFFGW_FUSED_ADAM", "1") != "0"
            make = lambda ps, lr: torch.optim.Adam(ps, lr=lr, fused=fused)      # noqa: E731
        else:
            make = lambda ps, lr: Adam(ps, lr=lr)                               # noqa: E731
        self.cla_opt = make([p for p in m.cla.parameters() if p.requires_grad], lr_cla)
        self.dis_opt = make([p for p in m.dis.parameters() if p.requires_grad], lr_dis)
        self.gen_opt = make([p for p in m.gen.parameters() if p.requires_grad], lr_gen)
        kw = {} if bucket_bytes is None else {"bucket_bytes": bucket_bytes}
        self.red = {"cla": GradientReducer(m.cla.parameters(), **kw), "dis": GradientReducer(m.dis.parameters(), **kw),
                    "gen": GradientReducer(m.gen.parameters(), **kw)}
        self.opt = {"cla": self.cla_opt, "dis": self.dis_opt, "gen": self.gen_opt}
        self.names = ("cla", "dis", "gen")
        self.graphable = ("cla", "dis", "gen")
        if rec is not None:
            # main_run.py:277: Adam at lr 1e-5 (optim.Adam for the native recogniser, torch's own for a foreign torch module)
            from .recognizer import RecModel as _Native
            rec_params = [p for p in m.rec.parameters() if p.requires_grad]
            self.rec_opt = make(rec_params, lr_rec) if isinstance(m.rec, _Native) else torch.optim.Adam(rec_params, lr=lr_rec)
            self.opt["rec"] = self.rec_opt
            self.red["rec"] = GradientReducer(m.rec.parameters(), **kw)
            self.names = ("rec", "cla", "dis", "gen")       # main_run.py:148-167 order
            self.graphable = ("cla", "dis")                 # rec_update / gen_update decode on the host (beam search)
        self.cer = None                                     # optional (CER(), CER(), CER()) accumulators: rec, gen, gen-swap
        # the reference computes dis / cla weight gradients inside gen_update and throws them away at the next
        # zero_grad (main_run.py:148-163); skipping them changes nothing observable
        self.skip_unused_wgrad = skip_unused_wgrad
        self.cuda_graph = bool(cuda_graph)
        # weight-gradient GEMMs on a second stream beside the dgrad / normalisation chain (ops.wgrad_side_stream)
        self.wgrad_stream = bool(wgrad_stream) and os.environ.get("AFFGW_WGRAD_STREAM", "1") != "0"
        self._graphs = None           # {name: (CUDAGraph, static outputs)}
        self._packed = {}             # sub-network -> packed operand copies its graphs read (ops.packed_entries), set at capture
        self._static_in = None
        self._eager_steps = 0
        self._side = None
        self.graph_launches = 0       # libaffgw launches recorded in the three graphs (= launches per replayed iteration)
        self.overlap_exchange = bool(overlap_exchange) and self.cuda_graph
        self._comm = None             # side stream of the overlapped exchange
        # cla_update touches only the classifier and dis_update never reads it: with private graph pools (overlap_exchange)
        # the two captured sub-steps replay side by side on two streams and fill each other's idle SMs / launch gaps
        self.concurrent_cla_dis = bool(concurrent_cla_dis) and self.overlap_exchange and \
            os.environ.get("AFFGW_CONCURRENT_CLA_DIS", "1") != "0"
        self._aux = None              # stream of the concurrent cla_update replay
        # OFF by default (bench.py times the reference's iteration as the reference composes it): one generator forward per
        # iteration instead of the reference's two identical ones (network_tro.ConTranModel.forward, `shared`).  Not with a
        # recogniser in graph mode: its gen_update is issued eagerly and cannot walk an autograd graph recorded at capture.
        self.share_generator_forward = bool(share_generator_forward) and not (rec is not None and self.cuda_graph)
        self._shared = {} if self.share_generator_forward else None
        # Both generator forwards of the reference's iteration are executed, but gen_update's is issued inside dis_update, right
        # after the no_grad one, and the discriminator's forward + backward runs beside it on a second stream
        # (network_tro.ConTranModel.forward, shared={"early": True}): HBM-bound 16/32-channel work next to tensor-bound work.
        # Default: on in the replayed configuration (graphs + overlapped exchange), like concurrent_cla_dis.
        if early_generator_forward is None:
            early_generator_forward = self.overlap_exchange and os.environ.get("AFFGW_EARLY_GEN", "1") != "0"
        self.early_generator_forward = bool(early_generator_forward) and not self.share_generator_forward and \
            not (rec is not None and self.cuda_graph)
        if self.early_generator_forward:
            self._shared = {"early": True}
        # gen_update: the classifier's pass over the generated pair beside the discriminator's (two independent critics)
        if concurrent_gen_heads is None:
            concurrent_gen_heads = self.overlap_exchange and os.environ.get("AFFGW_GEN_HEADS", "1") != "0"
        self.concurrent_gen_heads = bool(concurrent_gen_heads) and rec is None
        if self.concurrent_gen_heads:
            if self._shared is None:
                self._shared = {"share": False}
            self._shared["heads"] = True
        self._pending = {}            # sub-network -> event of its exchange + Adam queued on the side stream
        # the generator's text encoder on its own stream beside the image encoder (ConTranModel._generate_pair)
        m.side_text_encoder = self.overlap_exchange and os.environ.get("AFFGW_SIDE_TEXT", "1") != "0"
        broadcast_module(m)

    # ------------------------------------------------------------------------------------------------ sub-steps
    def _fwd_bwd(self, name, batch, epoch):
        """zero_grad + forward + backward of one sub-step; returns its loss tensors."""
        with ops.scratch_scope(name):
            return self._fwd_bwd_streams(name, batch, epoch)

    def _fwd_bwd_streams(self, name, batch, epoch):
        if self.wgrad_stream:
            with ops.wgrad_side_stream():       # forked inside every backward pass, joined before the losses are returned
                return self._fwd_bwd_inner(name, batch, epoch)
        with ops.accumulate_into_grad(True):
            return self._fwd_bwd_inner(name, batch, epoch)

    def _fwd_bwd_inner(self, name, batch, epoch):
        m = self.model
        self.opt[name].zero_grad()
        if name == "rec":
            return (m(batch, epoch, "rec_update", self.cer[0] if self.cer else None),)
        if name == "cla":
            return (m(batch, epoch, "cla_update"),)
        if name == "dis":
            return (m(batch, epoch, "dis_update", shared=self._shared),)
        frozen = []
        if self.skip_unused_wgrad:
            others = list(m.dis.parameters()) + list(m.cla.parameters()) + (list(m.rec.parameters()) if m.rec is not None else [])
            for p in others:
                if p.requires_grad:
                    p.requires_grad_(False)
                    frozen.append(p)
        try:
            l_total, l_dis_g, l_cla_g, _, l_rec_g = m(batch, epoch, "gen_update", self.cer[1:] if self.cer else None,
                                                      shared=self._shared)
        finally:
            for p in frozen:
                p.requires_grad_(True)
        return (l_total, l_dis_g, l_cla_g, l_rec_g)

    def _finish(self, name):
        self.red[name].reduce()
        self.opt[name].step()           # optim.Adam invalidates the packed-weight cache itself (ops.weights_updated)
        if not isinstance(self.opt[name], Adam):
            ops.weights_updated(p for grp in self.opt[name].param_groups for p in grp["params"])   # torch fused Adam
        # re-pack the tensor-core operand copies of the stepped weights here, next to the step (in place: the CUDA graphs read
        # these buffers and contain no packing kernels).  With overlap_exchange this runs on the side stream under the next
        # graph replay instead of in front of the first convolution that needs each weight.
        self._refresh_packed(name)

    def _refresh_packed(self, name):
        ents = self._packed.get(name)
        if ents is None:
            ents = ops.packed_entries(p for grp in self.opt[name].param_groups for p in grp["params"])
        return ops.refresh_packed(ents)

    @staticmethod
    def _pack(outs):
        (l_cla,), (l_dis,), (l_total, l_dis_g, l_cla_g, l_rec_g) = outs["cla"], outs["dis"], outs["gen"]
        res = {"cla": l_cla.detach(), "dis": l_dis.detach(), "gen": l_total.detach(), "gen_dis": l_dis_g.detach(),
               "gen_cla": l_cla_g.detach()}
        if "rec" in outs:
            res["rec"] = outs["rec"][0].detach()
            res["gen_rec"] = l_rec_g.detach()
        return res

    # ------------------------------------------------------------------------------------------------ overlapped exchange
    def join(self, *names):
        """Main stream waits for the exchange + optimiser step of the named sub-networks (all pending ones when none is
        named) queued on the side stream; no host synchronisation."""
        for name in (names or tuple(self._pending)):
            ev = self._pending.pop(name, None)
            if ev is not None:
                torch.cuda.current_stream().wait_event(ev)

    def _finish_on_side_stream(self, name):
        if self._comm is None:
            self._comm = torch.cuda.Stream(device=self.model.device_)
        self._comm.wait_stream(torch.cuda.current_stream())     # the graph that produced these gradients
        with torch.cuda.stream(self._comm):
            self._finish(name)
            ev = torch.cuda.Event()
            ev.record(self._comm)
        self._pending[name] = ev

    def train_step_eager(self, batch, epoch=0):
        self.join()
        outs = {}
        for name in self.names:
            outs[name] = self._fwd_bwd(name, batch, epoch)
            self._finish(name)
        return self._pack(outs)

    # ------------------------------------------------------------------------------------------------ graph replay
    def train_step(self, batch, epoch=0):
        if not self.cuda_graph:
            return self.train_step_eager(batch, epoch)
        if self._graphs is None:
            self.join()
            if self._eager_steps < self.GRAPH_WARMUP:
                # eager iterations on the stream the graphs will be captured on, so that the autograd nodes that outlive an
                # iteration (AccumulateGrad) belong to that stream
                self._eager_steps += 1
                if self._side is None:
                    self._side = torch.cuda.Stream(device=self.model.device_, priority=int(os.environ.get("AFFGW_MAIN_PRIO", "-1")))
                self._side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(self._side):
                    out = self.train_step_eager(batch, epoch)
                torch.cuda.current_stream().wait_stream(self._side)
                return out
            return self._capture(batch, epoch)      # the capture pass itself runs this iteration
        # a replay computes on the captured shapes only: a batch of another shape (the reference's loaders keep the last,
        # shorter batch of an epoch, main_run.py:123-130) runs eagerly instead - copy_ would raise, or silently broadcast a
        # single sample over the captured batch
        if not self._matches_capture(batch):
            return self.train_step_eager(batch, epoch)
        for dst, src in zip(self._static_in, batch):
            if torch.is_tensor(dst) and dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        # weights changed behind the trainer's back (load_state_dict between iterations): the graphs do not re-pack.  Normally
        # nothing is stale here (every step re-packs its own sub-network) and this is a host-side version check.
        for name in self.names:
            if any(e[0] != ops._WeightCache.version(e[4]) for e in self._packed.get(name, ())):
                self.join(name)
                self._refresh_packed(name)
        outs = {}
        side_by_side = self.concurrent_cla_dis and "cla" in self._graphs and "dis" in self._graphs
        for name in self.names:
            if name == "cla" and side_by_side:
                # cla_update on its own stream, dis_update follows on the main stream without waiting for it; gen_update's
                # join() below waits for the classifier's step like it always does
                if self._aux is None:
                    self._aux = torch.cuda.Stream(device=self.model.device_)
                self._aux.wait_stream(torch.cuda.current_stream())          # the batch copy above
                with torch.cuda.stream(self._aux):
                    graph, static_out, grads = self._graphs[name]
                    graph.replay()
                    for p, g in grads:
                        p.grad = g
                    outs[name] = static_out
                    self._finish_on_side_stream(name)                        # exchange + Adam behind the replay, on _comm
                continue
            # dis_update reads the generator stepped by the previous iteration, gen_update also the classifier stepped by this
            # one; cla_update reads neither, so the generator's exchange + Adam overlap it (and the classifier's dis_update)
            if name == "dis":
                self.join("gen")
            elif name == "gen":
                self.join()
            if name not in self._graphs:                 # sub-steps that call the recogniser: issued eagerly
                if name == "rec":
                    self.join("gen")                     # (nothing of the generator is read, but .grad buffers are shared state)
                outs[name] = self._fwd_bwd(name, self._static_in, epoch)
                self._finish(name)
                continue
            graph, static_out, grads = self._graphs[name]
            graph.replay()
            for p, g in grads:                      # an eager iteration in between may have re-pointed .grad
                p.grad = g
            outs[name] = static_out
            if self.overlap_exchange and name != "dis":
                self._finish_on_side_stream(name)
            else:
                self._finish(name)
        if "gen" in self._graphs:
            self.model.iter_num += 1                # the captured gen_update's Python side does not run in a replay
        # the static loss tensors are overwritten by the next replay: hand out copies
        return {k: v.clone() for k, v in self._pack(outs).items()}

    def _matches_capture(self, batch):
        if len(batch) != len(self._static_in):
            return False
        for dst, src in zip(self._static_in, batch):
            if torch.is_tensor(dst) != torch.is_tensor(src):
                return False
            if torch.is_tensor(dst) and (dst.shape != src.shape or dst.dtype != src.dtype):
                return False
        return True

    # ------------------------------------------------------------------------------------------------ safe reads
    def state_dict(self):
        """model.state_dict() after every queued exchange / optimiser step has been ordered before the current stream."""
        self.join()
        return self.model.state_dict()

    def eval_step(self, batch, epoch=0):
        """network_tro.py:140-177 (`eval` mode) on the stepped weights."""
        self.join()
        return self.model(batch, epoch, "eval")

    def _capture(self, batch, epoch):
        from . import _lib
        dev = self.model.device_
        self._static_in = tuple(t.to(dev).clone() if torch.is_tensor(t) else t for t in batch)
        torch.cuda.synchronize()
        # the graphs read the packed operand copies of the weights in place; every copy is (re)built outside the graphs, by the
        # optimiser step of its sub-network (_finish) - make all of them current before the first capture
        for name in self.names:
            self._refresh_packed(name)
        # the three graphs always replay in capture order: one shared pool - unless the exchange overlaps the next replay
        pool = None if self.overlap_exchange else torch.cuda.graph_pool_handle()
        graphs, outs = {}, {}
        self.graph_launches = 0
        n0 = _lib.launch_count()
        for name in self.names:
            if name not in self.graphable:
                outs[name] = self._fwd_bwd(name, self._static_in, epoch)
                self._finish(name)
                n0 = _lib.launch_count()
                continue
            self.opt[name].zero_grad(set_to_none=True)      # captured backward WRITES fresh .grad tensors (no accumulation)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool, stream=self._side):
                out = self._fwd_bwd(name, self._static_in, epoch)
            n_captured = _lib.launch_count() - n0
            g.replay()                              # run the sub-step for real: this call is one full training iteration
            grads = [(p, p.grad) for grp in self.opt[name].param_groups for p in grp["params"] if p.grad is not None]
            graphs[name] = (g, out, grads)
            outs[name] = out
            self._finish(name)
            self.graph_launches += n_captured
            n0 = _lib.launch_count()
        torch.cuda.synchronize()
        self._packed = {name: ops.packed_entries(p for grp in self.opt[name].param_groups for p in grp["params"])
                        for name in self.names}
        self._graphs = graphs                       # (iter_num was advanced by the captured gen_update's Python side)
        # static graph outputs, overwritten by the first replay: hand out copies here too
        return {k: v.clone() for k, v in self._pack(outs).items()}
