"""Data parallelism for the hot path, in place on the flat ParamStore buffers.

Two modes (both shard the episode batch, one process per GPU, NCCL over NVLink / NVSwitch):

* `GradientOverlap` — plain data parallel (the HF DDP path the reference takes when TrainerConfig.deepspeed is None,
  dexbotic/exp/trainer.py:110,121): bucketed all-reduce(AVG) of the bf16 gradient ranges as soon as a block's backward
  has written them, every rank then runs the whole AdamW.
* `ShardedDataParallel` — ZeRO-1 (the optimizer-state sharding of the reference's DeepSpeed configs,
  script/deepspeed/zero2.json / base_exp.py:229, without ZeRO-3's per-layer parameter traffic): each gradient chunk is
  REDUCE-SCATTERed as soon as it is final, every rank keeps fp32 master + Adam moments for 1/N of every chunk and
  updates only that, and the bf16 compute copy of the weights is ALL-GATHERed chunk by chunk under the next step's
  forward (block i waits for its own chunk only).  Same bytes on the wire as an all-reduce, but the half that
  competes with the backward GEMMs is gone, the 28 B/parameter AdamW sweep shrinks N-fold (39 ms -> 5 ms at N = 8 for
  the 7B model) and 8 B/parameter of optimizer state leave every GPU but one.

Transports of the sharded mode:
  "nccl"  reduce_scatter_tensor / all_gather_into_tensor in place on the flat buffers.  NCCL's kernels are resident on
          SMs while the persistent tcgen05 GEMM owns all 148 of them: both slow down (the 5-8 % weak-scaling loss
          measured at N = 2..8).
  "ce"    the same exchange over NVLink / NVSwitch PEER MEMORY, no NCCL kernel on the data path: the gradient and weight
          buffers are symmetric memory (ParamStore.SYMMETRIC).  Reduce-scatter = one small kernel per finished chunk and
          rank (csrc/exchange.cu: reduce_scatter_p2p_kernel) that LOADS the rank's piece out of every peer's gradient
          buffer (ld.global.cv on the peer-mapped addresses, several packs per peer in flight per thread), sums in fp32 in
          a fixed rank order and writes the average back in place — one pass, bracketed by stream-ordered cross-rank
          barriers on the symmetric-memory signal pads; 16 CTAs while backward runs (a block fits next to the GEMM's CTA
          on an SM), the whole GPU for the chunks left after backward.  All-gather = every rank PUSHES its piece of the
          new bf16 weights into the peers' weight buffers with peer copies on the COPY ENGINES.  Measured at N = 2 on one
          box: 155.8 samples/s against 153.4 with the NCCL transport (profiles/r2_bench_lines.md).
Gradients are produced by the tcgen05 wgrad GEMMs directly into the flat bf16 buffer the exchange reads — no packing,
no copies.
"""
from __future__ import annotations

import bisect
import os
from contextlib import contextmanager
from typing import Callable, Optional

import torch
import torch.distributed as dist

from .params import ALIGN, ParamStore


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
    is left (embeddings, projector, tower front end, fp32 action head) and joins the streams.

    Gradient accumulation: run the non-final micro-batches under `no_sync()` — a range is reduced ONCE per step, on
    the backward in which it becomes final (reducing it earlier would add later local gradients on top of an
    already averaged range)."""

    def __init__(self, store: ParamStore, group=None, bucket_bytes: int = 512 << 20, reserve_sms: int = 0):
        self.store, self.group, self.bucket_bytes = store, group, bucket_bytes
        self.enabled = dist.is_initialized() and dist.get_world_size(group) > 1
        # reserve_sms > 0: while collectives are in flight the persistent GEMM leaves that many SMs to NCCL (set
        # NCCL_MAX_CTAS to the same number before init_process_group) instead of queueing CTAs behind its kernels
        self.reserve_sms = reserve_sms
        self._limited = False
        self.sync = True        # False inside no_sync(): this backward is not the last one of the step
        self.works = []
        self.done = []          # list of (start, end) already reduced this step (element offsets in grad_a)
        self.pending = None     # (start, end) accumulated but not launched yet
        if self.enabled:
            store.grad_ready_hook = self.on_ready
            store.zero_grad_hook = self.reset

    @contextmanager
    def no_sync(self):
        """Micro-batches before the last one of an accumulation step: gradients stay local."""
        prev, self.sync = self.sync, False
        try:
            yield
        finally:
            self.sync = prev

    def reset(self) -> None:
        """A new step starts (ParamStore.zero_grad): nothing has been reduced yet."""
        for w in self.works:
            w.wait()
        self.works, self.done, self.pending = [], [], None

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
        if not self.sync:
            return
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
        if not self.enabled or not self.sync:
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


def partition_chunks(n_a: int, bounds: list, world: int) -> list[tuple[int, int]]:
    """Split region A [0, n_a) into the chunks the sharded optimizer exchanges: the given ranges (one per decoder block,
    forward order, None entries skipped) plus the gaps between / around them.  Every chunk is padded DOWN/UP to the
    ParamStore alignment already (multiples of ALIGN = 64 elements), so each splits into `world` equal 16-byte aligned
    pieces for world in {1, 2, 4, 8}.  Returns sorted (start, end) pairs covering [0, n_a) exactly."""
    assert ALIGN % (8 * world) == 0 or world == 1, f"world size {world} does not divide the {ALIGN}-element alignment"
    cuts = sorted({0, n_a} | {x for b in bounds if b is not None for x in b})
    chunks = [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
    for a, b in chunks:
        assert a % ALIGN == 0 and (b % ALIGN == 0 or b == n_a), (a, b)
    return chunks


class ShardedDataParallel:
    """ZeRO-1 over the flat buffers (see the module docstring).

    Usage (one process per GPU):
        dp = ShardedDataParallel(model.store)           # after the model (and its set_param_chunks) exists
        model.zero_grad(); loss.backward(); dp.finish(); model.optimizer_step(...)
    `ParamStore.adamw_step` delegates to `step()` while `store.sharder` is set.  state_dict() gathers the fp32 master
    shards first (`gather_master`).  With world size 1 nothing is sharded and the store's own path runs."""

    def __init__(self, store: ParamStore, group=None, adamw_fn: Optional[Callable] = None,
                 sumsq_fn: Optional[Callable] = None, clip_fn: Optional[Callable] = None, transport: str = "auto"):
        self.store, self.group = store, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.enabled = self.world > 1
        self.nccl = self.enabled and dist.get_backend(group) == "nccl"
        # kernels (CUDA) — injectable so that the world-size-2 gloo test can drive the host logic with torch stand-ins
        from . import ops
        self._ops = ops
        self._adamw = adamw_fn or ops.adamw_
        self._sumsq = sumsq_fn or ops.sumsq_
        self._clip = clip_fn or ops.clip_coef
        self.sync = True
        self.works: list = []
        self.reduced: set[int] = set()          # chunk ids reduce-scattered this step
        self.chunks = partition_chunks(store.n_a, store._chunk_bounds, self.world) if store.n_a else []
        self._starts = [c[0] for c in self.chunks]
        # chunk id of every decoder block (forward order) for wait_chunk(), and the remaining chunks ("rest")
        self.block_chunk = [None if b is None else self._starts.index(b[0]) for b in store._chunk_bounds]
        blk = {c for c in self.block_chunk if c is not None}
        self.rest = [i for i in range(len(self.chunks)) if i not in blk]
        # this rank's piece of every chunk and its offset in the local moment buffers
        self.piece: list[tuple[int, int]] = []
        self.local_off: list[int] = []
        off = 0
        for a, b in self.chunks:
            sz = (b - a) // self.world
            assert sz * self.world == b - a and sz % 8 == 0, f"chunk [{a},{b}) does not split into {self.world} pieces"
            self.piece.append((a + self.rank * sz, a + (self.rank + 1) * sz))
            self.local_off.append(off)
            off += sz
        self.n_local = off
        self.exp_avg = self.exp_avg_sq = None      # [n_local + n_b] fp32: this rank's pieces, then region B (replicated)
        self._side = None
        self._keep = None
        # copy-engine transport: both exchanged buffers must be symmetric memory
        want_ce = transport == "ce" or (transport == "auto" and getattr(store, "symmetric", False))
        self.ce = bool(self.enabled and self.nccl and want_ce)
        if self.ce:
            if not getattr(store, "symmetric", False):
                raise RuntimeError("transport='ce' needs ParamStore.SYMMETRIC = True before the model is built")
            import torch.distributed._symmetric_memory as symm_mem
            pg = group if group is not None else dist.group.WORLD
            self._h_grad = symm_mem.rendezvous(store.grad_a, pg)      # collective: maps every peer's buffer
            self._h_shadow = symm_mem.rendezvous(store.shadow, pg)
            self._comm = torch.cuda.Stream(device=store.device)
            # base addresses of every peer's gradient buffer as this process sees them (peer-mapped): the reduce-scatter
            # kernel loads its pieces straight from them.  Order = rank order after this rank: a fixed summation order.
            self._peer_grad = [self._h_grad.get_buffer((self.rank + k) % self.world, (store.grad_a.numel(),),
                                                       torch.bfloat16, 0) for k in range(1, self.world)]
            self._rs_ctas = int(os.environ.get("B200_RS_CTAS", "16"))
            # after backward (finish()) nothing else runs: the leftover chunks (embedding table, towers, projector, heads)
            # are exposed time, so their kernels take the whole GPU — NVLink loads are latency-bound, bandwidth scales
            # with the bytes in flight (770 GB/s x ~3 us = 2.3 MB = 16 B x 145 k threads at N = 2)
            self._rs_ctas_idle = int(os.environ.get("B200_RS_CTAS_IDLE", "592"))
            self._in_finish = False
        if self.enabled:
            store.sharder = self
            store.grad_ready_hook = self.on_ready
            store.zero_grad_hook = self.reset

    # ------------------------------------------------------------------ gradient exchange
    @contextmanager
    def no_sync(self):
        prev, self.sync = self.sync, False
        try:
            yield
        finally:
            self.sync = prev

    def reset(self) -> None:
        for w in self.works:
            w.wait()
        self.works = []
        self.reduced = set()

    def _reduce_scatter_ce(self, ci: int) -> None:
        """Reduce-scatter of one chunk over peer memory: after a cross-rank barrier (every rank has enqueued this
        chunk's wgrads before it) ONE small kernel (ops.reduce_scatter_p2p_, exchange.cu) loads piece `rank` out of every
        peer's gradient buffer over NVLink, sums in fp32 in a fixed order and writes the average back in place; a second
        barrier tells every rank that its chunk has been read.  Runs on the exchange stream; the compute stream goes on
        with the next block's backward."""
        st = self.store
        pa, pb = self.piece[ci]
        cs = self._comm
        cs.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cs):
            self._h_grad.barrier(channel=0)
            if pb > pa:
                self._ops.reduce_scatter_p2p_(st.grad_a[pa:pb], [t[pa:pb] for t in self._peer_grad], 1.0 / self.world,
                                              ctas=self._rs_ctas_idle if self._in_finish else self._rs_ctas)
            self._h_grad.barrier(channel=0)            # every peer has taken its piece of this rank's chunk
        self.reduced.add(ci)

    def _push_shadow_ce(self, ci: int) -> None:
        """Copy-engine all-gather of one chunk's new bf16 weights: push this rank's piece into every peer's weight
        buffer, then a barrier — when it completes on a rank, all N pieces of the chunk have landed there.  (No peer
        can still be reading the chunk: its reduce-scatter barrier of this step came after its last use.)"""
        st = self.store
        pa, pb = self.piece[ci]
        sz = pb - pa
        for step in range(1, self.world):
            peer = (self.rank + step) % self.world
            self._h_shadow.get_buffer(peer, (sz,), torch.bfloat16, pa).copy_(st.shadow[pa:pb])
        self._h_shadow.barrier(channel=0)

    def _reduce_scatter(self, ci: int) -> None:
        if self.ce:
            return self._reduce_scatter_ce(ci)
        a, b = self.chunks[ci]
        buf = self.store.grad_a[a:b]
        pa, pb = self.piece[ci]
        if self.nccl:        # in place: the output is this rank's slice of the input
            self.works.append(dist.reduce_scatter_tensor(self.store.grad_a[pa:pb], buf, op=dist.ReduceOp.AVG,
                                                         group=self.group, async_op=True))
        else:                # gloo has no reduce-scatter: all-reduce and keep the own piece (CPU tests)
            t = buf.float()
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            self.store.grad_a[pa:pb].copy_((t[pa - a:pb - a] / self.world).to(buf.dtype))
        self.reduced.add(ci)

    def on_ready(self, a: int, b: int) -> None:
        """grad_a[a:b] (a decoder block's range) is final: its reduce-scatter starts under the rest of backward."""
        if not (self.enabled and self.sync):
            return
        i = bisect.bisect_right(self._starts, a) - 1
        while i < len(self.chunks) and self.chunks[i][0] < b:
            if i not in self.reduced and self.chunks[i][0] >= a and self.chunks[i][1] <= b:
                self._reduce_scatter(i)
            i += 1

    def finish(self) -> None:
        """After backward: the chunks no hook reported (embeddings, towers, projector, heads) and the fp32 region."""
        if not (self.enabled and self.sync):
            return
        self._in_finish = True
        try:
            for ci in range(len(self.chunks)):
                if ci not in self.reduced:
                    self._reduce_scatter(ci)
        finally:
            self._in_finish = False
        st = self.store
        if st.n_b:
            if self.nccl:
                self.works.append(dist.all_reduce(st.grad_b, op=dist.ReduceOp.AVG, group=self.group, async_op=True))
            else:
                dist.all_reduce(st.grad_b, op=dist.ReduceOp.SUM, group=self.group)
                st.grad_b.div_(self.world)
        for w in self.works:
            w.wait()
        self.works = []
        if self.ce:
            torch.cuda.current_stream().wait_stream(self._comm)

    # ------------------------------------------------------------------ optimizer
    def _segments_of(self, segs: list, lo: int, hi: int) -> list:
        """Intersections of the (start, end, lr, wd, region) runs with the region-A range [lo, hi)."""
        out = []
        for a, b, lr, wd, region in segs:
            if region != "A":
                continue
            x, y = max(a, lo), min(b, hi)
            if x < y:
                out.append((x, y, lr, wd))
        return out

    def step(self, lrs: dict, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
             max_grad_norm: Optional[float] = 1.0):
        """Sharded AdamW: global-norm clip from the shard norms, update of this rank's pieces (and of the replicated
        fp32 region), all-gather of the bf16 shadow.  Returns the (device) gradient norm."""
        st = self.store
        dev = st.device
        if self.exp_avg is None:
            self.exp_avg = torch.zeros(self.n_local + st.n_b, device=dev, dtype=torch.float32)
            self.exp_avg_sq = torch.zeros(self.n_local + st.n_b, device=dev, dtype=torch.float32)
        st.step_count += 1
        st.finalize_grads()
        st.wait_all_params()
        clip = None
        norm = torch.zeros((), device=dev, dtype=torch.float32)
        if max_grad_norm is not None:
            ssq = torch.zeros((), device=dev, dtype=torch.float32)
            for pa, pb in self.piece:
                if pb > pa:
                    self._sumsq(st.grad_a[pa:pb], ssq)
            dist.all_reduce(ssq, op=dist.ReduceOp.SUM, group=self.group)      # 4 bytes
            if st.n_b:
                self._sumsq(st.grad_b, ssq)                                     # replicated: counted once
            clip = torch.empty((), device=dev, dtype=torch.float32)
            self._clip(ssq, max_grad_norm, clip, norm)
        segs = st.segments(lrs, weight_decay)

        def update_piece(ci: int) -> None:
            pa, pb = self.piece[ci]
            for x, y, lr, wd in self._segments_of(segs, pa, pb):
                o = self.local_off[ci] + (x - pa)
                self._adamw(st.master[x:y], st.grad_a[x:y], self.exp_avg[o:o + y - x], self.exp_avg_sq[o:o + y - x],
                            st.shadow[x:y], lr, betas[0], betas[1], eps, wd, st.step_count, clip)

        def gather(ci: int) -> None:
            if self.ce:
                return self._push_shadow_ce(ci)
            a, b = self.chunks[ci]
            pa, pb = self.piece[ci]
            if self.nccl:
                dist.all_gather_into_tensor(st.shadow[a:b], st.shadow[pa:pb], group=self.group)
            else:
                parts = [torch.empty(pb - pa, dtype=st.shadow.dtype, device=dev) for _ in range(self.world)]
                dist.all_gather(parts, st.shadow[pa:pb].contiguous(), group=self.group)
                st.shadow[a:b].copy_(torch.cat(parts))

        def region_b() -> None:
            for a, b, lr, wd, region in segs:
                if region != "B":
                    continue
                o = self.n_local + (a - st.n_a)
                self._adamw(st.master[a:b], st.grad_b[a - st.n_a:b - st.n_a], self.exp_avg[o:o + b - a],
                            self.exp_avg_sq[o:o + b - a], None, lr, betas[0], betas[1], eps, wd, st.step_count, clip)

        use_side = st.async_optimizer and dev.type == "cuda"
        if not use_side:
            region_b()
            for ci in range(len(self.chunks)):
                update_piece(ci)
                gather(ci)
            return norm
        # everything outside the decoder blocks first (the towers run first), then block by block in forward order on a
        # side stream; the caller's stream waits for the first part only, block i's forward for its own chunk
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream(device=dev)
        side = self._side
        ready = torch.cuda.Event()
        ready.record(main)
        side.wait_event(ready)
        self._keep = (clip, norm)
        with torch.cuda.stream(side):
            region_b()
            for ci in self.rest:
                update_piece(ci)
                gather(ci)
            ev_rest = torch.cuda.Event()
            ev_rest.record(side)
            for i, ci in enumerate(self.block_chunk):
                if ci is None:
                    continue
                update_piece(ci)
                gather(ci)
                ev = torch.cuda.Event()
                ev.record(side)
                st._chunk_events[i] = ev
        main.wait_event(ev_rest)
        return norm

    # ------------------------------------------------------------------ checkpoints
    def gather_master(self) -> None:
        """fp32 master weights: every rank owns the authoritative copy of its pieces only; collect them before a
        state_dict() / save_pretrained()."""
        if not self.enabled:
            return
        st = self.store
        st.wait_all_params()
        for (a, b), (pa, pb) in zip(self.chunks, self.piece):
            if self.nccl:
                dist.all_gather_into_tensor(st.master[a:b], st.master[pa:pb], group=self.group)
            else:
                parts = [torch.empty(pb - pa, dtype=torch.float32, device=st.device) for _ in range(self.world)]
                dist.all_gather(parts, st.master[pa:pb].contiguous(), group=self.group)
                st.master[a:b].copy_(torch.cat(parts))

    def optimizer_state_bytes(self) -> int:
        return 0 if self.exp_avg is None else 2 * self.exp_avg.numel() * 4
