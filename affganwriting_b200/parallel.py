"""Data-parallel training: one process per GPU, batch sharded on dim 0, bucketed NCCL all-reduce of gradients.

Replaces the reference's only multi-GPU mechanism - nn.DataParallel around the six VGG slices
(reference modules_tro.py:341-346: replicate / scatter / gather every forward, reduce-add every backward) - by the
DDP scheme of SURVEY.md §8(e):
  * every rank holds a full replica (weights broadcast from rank 0 once), sees batch/world samples;
  * after the backward of one sub-network (cla, dis, gen - each has its own optimiser, main_run.py:275-278) the
    gradients of that sub-network are packed into ~25 MB fp32 buckets by a libaffgw kernel, all-reduced (sum) by
    NCCL over NVLink / NVSwitch, and unpacked with the 1/world scale folded in;
  * tensors whose .grad is None are skipped exactly like torch.optim.Adam skips them (96 generator tensors never
    receive a gradient);
  * BatchNorm batch statistics stay per rank (each rank is an independent reference replica).
torch.distributed is used for rendezvous and the NCCL collective only.
"""
import torch
import torch.distributed as dist

from . import _lib as L  # noqa: N812

DEFAULT_BUCKET_BYTES = 25 * 1024 * 1024


def plan_buckets(sizes, bucket_elems):
    """Greedy, order-preserving split of tensor sizes into buckets of at most bucket_elems elements
    (a tensor larger than the cap gets a bucket of its own).  Returns a list of index lists."""
    buckets, cur, cur_n = [], [], 0
    for i, n in enumerate(sizes):
        if cur and cur_n + n > bucket_elems:
            buckets.append(cur)
            cur, cur_n = [], 0
        cur.append(i)
        cur_n += n
    if cur:
        buckets.append(cur)
    return buckets


class CudaPacker:
    """Gather / scatter a list of fp32 gradient tensors into one contiguous bucket with one kernel launch each."""

    def pack(self, grads, bucket):
        ptrs, sizes, offs = self._tables(grads, bucket.device)[:3]
        L.call("affgw_bucket_pack", ptrs.data_ptr(), sizes.data_ptr(), offs.data_ptr(), len(grads), bucket.data_ptr(),
               L.stream())

    def unpack(self, grads, bucket, scale):
        ptrs, sizes, offs = self._tables(grads, bucket.device)[:3]
        L.call("affgw_bucket_unpack", ptrs.data_ptr(), sizes.data_ptr(), offs.data_ptr(), len(grads), bucket.data_ptr(),
               float(scale), L.stream())

    def _tables(self, grads, device):
        # pointer / size / offset tables on the device; cached by the pointer tuple, which is constant from one iteration
        # to the next when the backward passes are replayed as CUDA graphs (static gradient buffers)
        ptrs = tuple(g.data_ptr() for g in grads)
        cache = self.__dict__.setdefault("_table_cache", {})
        hit = cache.get(ptrs)
        if hit is not None:
            return hit
        n = [g.numel() for g in grads]
        off, acc = [], 0
        for k in n:
            off.append(acc)
            acc += k
        host = torch.tensor([list(ptrs), n, off], dtype=torch.int64).pin_memory()
        dev = host.to(device, non_blocking=True)
        if len(cache) > 256:
            torch.cuda.synchronize()        # a side stream may still be reading the tables about to be freed
            cache.clear()
        cache[ptrs] = (dev[0], dev[1], dev[2], host)          # keep the pinned source alive until the copy has run
        return cache[ptrs]


class GradientReducer:
    """All-reduce (mean) the gradients of `params` across the process group, bucket by bucket."""

    def __init__(self, params, bucket_bytes=DEFAULT_BUCKET_BYTES, group=None, packer=None):
        self.params = [p for p in params if p.requires_grad]
        self.bucket_elems = max(1, bucket_bytes // 4)
        self.group = group
        self.packer = packer if packer is not None else CudaPacker()
        self.last_buckets = 0

    def reduce(self):
        if not dist.is_available() or not dist.is_initialized():
            return 0
        world = dist.get_world_size(self.group)
        if world == 1:
            return 0
        live = [p for p in self.params if p.grad is not None]
        if not live:
            return 0
        grads = []
        for p in live:
            if p.grad.dtype != torch.float32:
                raise RuntimeError("gradient buckets are fp32")
            if not p.grad.is_contiguous():
                p.grad = p.grad.contiguous()
            grads.append(p.grad)
        plans = plan_buckets([g.numel() for g in grads], self.bucket_elems)
        pending = []
        for idx in plans:
            gs = [grads[i] for i in idx]
            bucket = torch.empty(sum(g.numel() for g in gs), dtype=torch.float32, device=gs[0].device)
            self.packer.pack(gs, bucket)
            work = dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            pending.append((gs, bucket, work))
        for gs, bucket, work in pending:      # later buckets' packs overlap earlier buckets' collectives
            work.wait()
            self.packer.unpack(gs, bucket, 1.0 / world)
        self.last_buckets = len(plans)
        return len(plans)


def broadcast_module(module, src=0, group=None):
    """Identical initial weights and buffers on every rank (the reference has a single replica)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)


def shard_batch(batch, rank, world):
    """Equal contiguous shards of dim 0 (equal sizes keep mean-of-means == global mean)."""
    out = []
    for t in batch:
        if hasattr(t, "shape") and len(t.shape) > 0:
            n = t.shape[0]
            if n % world:
                raise ValueError(f"batch of {n} does not split evenly over {world} ranks")
            k = n // world
            out.append(t[rank * k:(rank + 1) * k])
        else:
            out.append(t)
    return tuple(out)
