"""Generation (the reference's tt.test_single_writer.* / helpers.generate_from_batch loops: enc_image -> enc_text -> mix ->
decode under no_grad, tt.test_single_writer.4_scenarios.py:150-158) replayed as a CUDA graph.

A generator forward is ~400 small kernels; issued from Python it is host-bound at batch 64.  GraphedGenerator captures
one forward for a fixed (batch, style planes) shape and replays it: inputs are copied into static tensors, the image
comes back in a static tensor (clone it if it has to outlive the next call).  Weights are read at replay time, so a
load_state_dict / optimiser step between calls is picked up as long as ops.weights_updated(gen) is called - the packed
operand copies are rebuilt inside the graph on every replay.
"""
import torch

from . import ops


class GraphedGenerator:
    """`relaxed=True` (mode 'f16', VGG encoder; off by default): VGG convolutions 9-16 and the decoder's ResBlock convolutions
    run one tensor-core pass instead of three (ops.relaxed_forward) - the image moves from 1.4e-3 to 5.4e-3 of the fp32
    reference's (bar 2e-2, tests/test_gpu_parity_c50.py) for 1.2x the images per second."""
    RELAXED_VGG_FROM = 26

    def __init__(self, gen, warmup=2, relaxed=False):
        self.gen = gen
        self.relaxed = bool(relaxed)
        self.warmup = int(warmup)
        self._graph = None
        self._shape = None
        self._calls = 0
        self._side = None

    @torch.no_grad()
    def __call__(self, tr_img, label):
        key = (tuple(tr_img.shape), tuple(label.shape), self.gen.training, ops.precision())     # a graph holds one mode's kernels
        if self._graph is None or key != self._shape:
            if self._shape != key:
                self._graph, self._calls, self._shape = None, 0, key
            if self._side is None:
                self._side = torch.cuda.Stream(device=tr_img.device)
            if self._calls < self.warmup:                 # eager calls on the capture stream: lazy initialisation, allocator
                self._calls += 1
                self._side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(self._side), ops.relaxed_forward(self.relaxed, self.RELAXED_VGG_FROM):
                    out = self.gen(tr_img, label)
                torch.cuda.current_stream().wait_stream(self._side)
                return out
            self._img, self._lab = tr_img.clone(), label.clone()
            torch.cuda.synchronize()
            ops.clear_weight_cache(self.gen)              # every packed weight the graph reads is packed inside it
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph, stream=self._side), ops.relaxed_forward(self.relaxed, self.RELAXED_VGG_FROM):
                self._out = self.gen(self._img, self._lab)
        self._img.copy_(tr_img, non_blocking=True)
        self._lab.copy_(label, non_blocking=True)
        self._graph.replay()
        return self._out


class GraphedForward:
    """Any no-grad forward `module(*tensors)` of this package replayed as a CUDA graph (the line-level generator,
    line_generation/generate.py's loop; GenModel_FC with the DINOv2 encoder): same contract as GraphedGenerator - fixed
    input shapes per captured graph, inputs copied into static tensors, the output lives in a static tensor, weights read
    at replay time.  Random draws inside the forward (the line generator's noise injections, pure_gen.py:199,205) come from
    torch's graph-safe generator: every replay draws fresh noise."""

    def __init__(self, module, warmup=2):
        self.module = module
        self.warmup = int(warmup)
        self._graph = None
        self._key = None
        self._calls = 0
        self._side = None

    @torch.no_grad()
    def __call__(self, *tensors):
        key = tuple((tuple(t.shape), t.dtype) for t in tensors) + (self.module.training, ops.precision())
        if self._graph is None or key != self._key:
            if self._key != key:
                self._graph, self._calls, self._key = None, 0, key
            if self._side is None:
                self._side = torch.cuda.Stream(device=tensors[0].device)
            if self._calls < self.warmup:
                self._calls += 1
                self._side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(self._side):
                    out = self.module(*tensors)
                torch.cuda.current_stream().wait_stream(self._side)
                return out
            self._in = tuple(t.clone() for t in tensors)
            torch.cuda.synchronize()
            ops.clear_weight_cache(self.module)
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph, stream=self._side):
                self._out = self.module(*self._in)
        for dst, src in zip(self._in, tensors):
            dst.copy_(src, non_blocking=True)
        self._graph.replay()
        return self._out
