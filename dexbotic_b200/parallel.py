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


class GradientOverlap:
    """Bucketed gradient all-reduce overlapped with backward.

    `TransformerBlockFn.backward` reports each block's gradient range in the flat bf16 buffer as soon as that
    block's wgrad kernels are enqueued (`ParamStore.grad_ready_hook`); consecutive ranges are merged into
    buckets of >= `bucket_bytes` and all-reduced asynchronously on NCCL's stream (which first waits for the
    compute stream), so NVLink traffic hides behind the backward of the earlier layers.  `finish()` reduces what
    is left (embeddings, projector, tower front end, fp32 action head) and joins the streams."""

    def __init__(self, store: ParamStore, group=None, bucket_bytes: int = 512 << 20, reserve_sms: int = 0):
        self.store, self.group, self.bucket_bytes = store, group, bucket_bytes
        self.enabled = dist.is_initialized() and dist.get_world_size(group) > 1
        # reserve_sms > 0: while collectives are in flight the persistent GEMM leaves that many SMs to NCCL (set
        # NCCL_MAX_CTAS to the same number before init_process_group) instead of queueing CTAs behind its kernels
        self.reserve_sms = reserve_sms
        self._limited = False
        self.works = []
        self.done = []          # list of (start, end) already reduced this step (element offsets in grad_a)
        self.pending = None     # (start, end) accumulated but not launched yet
        if self.enabled:
            store.grad_ready_hook = self.on_ready

    def _limit(self, on: bool) -> None:
        if self.reserve_sms > 0 and on != self._limited:
            from . import _lib
            n = torch.cuda.get_device_properties(self.store.device).multi_processor_count
            _lib.load().b200_set_gemm_sm_limit((n - self.reserve_sms) // 2 * 2 if on else 0)
            self._limited = on

    def _launch(self, a: int, b: int) -> None:
        self._limit(True)
        chunk = self.store.grad_a[a:b]
        self.works.append(dist.all_reduce(chunk, op=dist.ReduceOp.AVG, group=self.group, async_op=True))
        self.done.append((a, b))

    def on_ready(self, a: int, b: int) -> None:
        if self.pending is None:
            self.pending = (a, b)
        elif b == self.pending[0]:                 # backward walks the layers from the end: ranges grow downwards
            self.pending = (a, self.pending[1])
        elif a == self.pending[1]:
            self.pending = (self.pending[0], b)
        else:                                       # not adjacent: flush what we have
            self._launch(*self.pending)
            self.pending = (a, b)
        if (self.pending[1] - self.pending[0]) * 2 >= self.bucket_bytes:
            self._launch(*self.pending)
            self.pending = None

    def finish(self) -> None:
        if not self.enabled:
            return
        if self.pending is not None:
            self._launch(*self.pending)
            self.pending = None
        # everything not covered by a reported range
        covered = sorted(self.done)
        cur = 0
        for a, b in covered + [(self.store.n_a, self.store.n_a)]:
            if a > cur:
                self.works.append(dist.all_reduce(self.store.grad_a[cur:a], op=dist.ReduceOp.AVG, group=self.group,
                                                  async_op=True))
            cur = max(cur, b)
        if self.store.n_b:
            self.works.append(dist.all_reduce(self.store.grad_b, op=dist.ReduceOp.AVG, group=self.group, async_op=True))
        for w in self.works:
            w.wait()
        self._limit(False)
        self.works, self.done = [], []
