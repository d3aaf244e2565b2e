"""Data parallelism for the hot path: gradients only (north_star), in place on the flat buffers.

Reference behaviour replaced: HF Trainer wrapping the model in DistributedDataParallel when
TrainerConfig.deepspeed is None (dexbotic/exp/trainer.py:110,121) — a bucketed gradient all-reduce; the
reference default (DeepSpeed ZeRO-3, base_exp.py:229) additionally all-gathers parameters per layer, which
this backend does not need (180 GB of HBM holds the full replica).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .params import ParamStore


def allreduce_gradients(store: ParamStore, group=None, bucket_elems: int = 1 << 28) -> None:
    """Average grad_a / grad_b across ranks in place.  NCCL: ReduceOp.AVG; gloo (CPU tests): SUM then scale."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    world = dist.get_world_size(group)
    nccl = dist.get_backend(group) == "nccl"
    for buf in (store.grad_a, store.grad_b):
        n = buf.numel()
        for a in range(0, n, bucket_elems):
            chunk = buf[a:min(n, a + bucket_elems)]
            if nccl:
                dist.all_reduce(chunk, op=dist.ReduceOp.AVG, group=group)
            else:
                t = chunk.float()
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
                chunk.copy_((t / world).to(chunk.dtype))


def broadcast_parameters(store: ParamStore, src: int = 0, group=None) -> None:
    """Rank `src`'s master weights to everyone (identical replicas at start); shadows are refreshed by the caller."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(store.master, src=src, group=group)
