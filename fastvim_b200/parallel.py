"""Batch-sharded training plumbing (images are independent, the only exchange is the gradient sum).

One process per GPU.  The reference trains under PyTorch-Lightning DDP (``imagenet_classification/train.py:34-43``):
gradients are bucketed and every bucket's NCCL all-reduce overlaps the rest of the backward pass.  ``GradExchange`` does
the same for a training step that is CAPTURED IN ONE CUDA GRAPH (forward + backward + exchange + optimizer):

    exch = GradExchange(params, world)      # buckets in reverse parameter order (~ the order gradients become ready)
    exch.attach()                           # flat fp32 bucket buffers allocated; hooks installed
    with torch.cuda.graph(g):
        exch.begin()                        # gradients dropped: autograd assigns fresh ones (no accumulate kernels)
        loss = fwd_bwd(...)                 # hook of a bucket's last gradient: ONE multi-tensor copy packs the bucket, the
                                            #   side stream waits for the backward so far and all-reduces (AVG) the bucket
                                            #   (NCCL over NVLink/NVSwitch)
        exch.finish()                       # main stream joins the side stream; .grad -> slices of the reduced buckets
        optimizer.step()

so FastVim-B's 392 MB of fp32 gradients travel in ~6 buckets while earlier blocks are still back-propagating; only the
last bucket (the first block + patch embedding) is exposed.  On CPU tensors (gloo, the world-size-2 tests) the same object
runs without streams.  ``flatten_grads`` / ``allreduce_sum_`` / ``scatter_mean_grads_`` are the round-1 un-overlapped
helpers, kept for the eager warm-up and the tests.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def flatten_grads(grads: Sequence[torch.Tensor]) -> torch.Tensor:
    """One contiguous 1-D buffer holding every gradient (a copy; capturable in a CUDA graph)."""
    return torch.cat([g.reshape(-1) for g in grads])


def allreduce_sum_(flat: torch.Tensor, group=None) -> torch.Tensor:
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def scatter_mean_grads_(grads: List[torch.Tensor], flat: torch.Tensor, world: int) -> None:
    """grads[i] <- flat[segment i] / world, in place (capturable)."""
    if world > 1:
        flat.mul_(1.0 / world)
    torch._foreach_copy_(list(grads), [c.view_as(g) for c, g in zip(flat.split([g.numel() for g in grads]), grads)])


def plan_buckets(numels: Sequence[int], cap_elems: int) -> List[List[int]]:
    """Indices of ``numels`` grouped into buckets of at most ``cap_elems`` elements, walking the list BACKWARDS (the last
    parameters receive their gradients first).  A single tensor larger than the cap gets its own bucket."""
    buckets, cur, cur_n = [], [], 0
    for i in range(len(numels) - 1, -1, -1):
        n = int(numels[i])
        if cur and cur_n + n > cap_elems:
            buckets.append(cur)
            cur, cur_n = [], 0
        cur.append(i)
        cur_n += n
    if cur:
        buckets.append(cur)
    return buckets


class GradExchange:
    """Bucketed gradient all-reduce (mean) overlapped with the backward pass; capturable in a CUDA graph."""

    def __init__(self, params: Sequence[torch.nn.Parameter], world: int, bucket_mb: float = 64.0, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.world, self.group = world, group
        self.plan = plan_buckets([p.numel() for p in self.params], max(1, int(bucket_mb * (1 << 20) / 4)))
        self.flats: List[torch.Tensor] = []
        self.pending: List[int] = []
        self.fired: List[bool] = []
        self.handles = []
        self.side: Optional[torch.cuda.Stream] = None
        self._bucket_of = {}

    # -- un-overlapped exchange for eager warm-up steps (before attach) ------------------------------------------------
    def eager_allreduce_(self):
        grads = [p.grad for p in self.params if p.grad is not None]
        if not grads or self.world <= 1:
            return
        flat = flatten_grads(grads)
        allreduce_sum_(flat, self.group)
        scatter_mean_grads_(grads, flat, self.world)

    # -- capture-time wiring --------------------------------------------------------------------------------------------
    def attach(self) -> str:
        """Allocates the flat fp32 buckets and installs the hooks.  Gradients are NOT made views of the buckets: with
        ``.grad`` pre-set autograd would run one accumulate kernel per parameter (~300 launches per step for FastVim-B,
        measured as ~1 ms inside the captured step) plus a zero pass; instead autograd assigns fresh gradients and the hook
        of a bucket's last parameter packs the whole bucket with ONE multi-tensor copy before the all-reduce."""
        dev = self.params[0].device
        self.views = [None] * len(self.params)
        for b, idxs in enumerate(self.plan):
            n = sum(self.params[i].numel() for i in idxs)
            flat = torch.zeros(n, device=dev, dtype=torch.float32)
            off = 0
            for i in idxs:
                p = self.params[i]
                if p.dtype != torch.float32:
                    raise TypeError("GradExchange expects fp32 master parameters")
                self.views[i] = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
                self._bucket_of[id(p)] = b
                self.handles.append(p.register_post_accumulate_grad_hook(self._hook))
            self.flats.append(flat)
        self.pending = [len(idxs) for idxs in self.plan]
        self.fired = [False] * len(self.plan)
        if dev.type == "cuda":
            self.side = torch.cuda.Stream(dev)
        mb = [round(f.numel() * 4 / 2**20, 1) for f in self.flats]
        return (f"GradExchange: {len(self.flats)} buckets ({mb} MB fp32) in reverse parameter order; a bucket is packed by one "
                f"multi-tensor copy when its last gradient arrives and all-reduced (NCCL AVG) on a side stream inside the "
                f"captured backward")

    def detach(self):
        for h in self.handles:
            h.remove()
        self.handles = []

    def begin(self):
        """Start of a step (inside the captured region): drop the previous gradients (autograd then ASSIGNS new ones
        instead of accumulating), re-arm the hooks."""
        for p in self.params:
            p.grad = None
        self.pending = [len(idxs) for idxs in self.plan]
        self.fired = [False] * len(self.plan)

    def _reduce(self, b: int):
        flat = self.flats[b]
        idxs = [i for i in self.plan[b]]
        have = [i for i in idxs if self.params[i].grad is not None]
        miss = [i for i in idxs if self.params[i].grad is None]
        if have:
            torch._foreach_copy_([self.views[i] for i in have], [self.params[i].grad for i in have])
        for i in miss:                      # unused parameter: contributes zeros
            self.views[i].zero_()
        if self.world > 1:
            if flat.is_cuda:
                self.side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(self.side):
                    dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
            else:   # gloo has no AVG
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
                flat.mul_(1.0 / self.world)
        self.fired[b] = True

    def _hook(self, p):
        b = self._bucket_of[id(p)]
        self.pending[b] -= 1
        if self.pending[b] == 0 and not self.fired[b]:
            self._reduce(b)

    def finish(self):
        """End of the backward: reduce buckets whose hooks never completed (unused parameters), make the main stream wait
        for every bucket, and point every ``.grad`` at its slice of the reduced buckets (no copy back: the optimizer reads
        the bucket memory directly)."""
        for b in range(len(self.flats)):
            if not self.fired[b]:
                self._reduce(b)
        if self.side is not None:
            torch.cuda.current_stream().wait_stream(self.side)
        for i, p in enumerate(self.params):
            p.grad = self.views[i]
