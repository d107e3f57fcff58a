"""Batch-sharded training plumbing (images are independent, the only exchange is the gradient sum).

One process per GPU.  The training step is captured in two CUDA graphs (forward + backward | optimizer); between them
the gradients of all parameters travel in ONE flat buffer through ONE collective:

    graph 1:  forward, backward, flat = flatten_grads(grads)
    eager  :  allreduce_sum_(flat)                       # NCCL over NVLink / NVSwitch (gloo in the CPU tests)
    graph 2:  scatter_mean_grads_(grads, flat, world)    # 1/world scaling folded into the copy back
              optimizer.step()

FastVim-B moves 392 MB of fp32 gradients per step: ~1 ms on NVLink 5 against a 45 ms step, so the exchange is not
overlapped with the backward pass; torch DDP (bucketed, overlapped) remains available as ``bench.py --no-graph``.
"""
from __future__ import annotations

from typing import List, Sequence

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
