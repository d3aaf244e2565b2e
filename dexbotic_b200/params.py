"""Flat parameter / gradient / optimizer-state storage for the B200 backend.

HBM layout: one fp32 master buffer (the nn.Parameters are views of it, so state_dict keys,
shapes and dtype match the reference's fp32 checkpoints), one bf16 shadow buffer that the tensor-core
kernels read, one bf16 gradient buffer (fp32 for the action head, which the reference computes in
fp32/TF32 — cogact_arch.py:133), and fp32 Adam moments.  Flat buffers make the optimizer one launch per
(lr, weight-decay) segment and the data-parallel all-reduce a handful of large NCCL calls with no copies.

Reference behaviour mirrored: torch.optim.AdamW via HF Trainer.create_optimizer (dexbotic/exp/trainer.py:
25-36), parameter groups from OptimizerConfig._get_optimizer_grouped_parameters (base_exp.py:95-203),
max_grad_norm=1.0 (trainer.py:122).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import torch

from . import ops

ALIGN = 64  # elements; keeps every tensor 128-byte aligned in the bf16 buffers (TMA needs 16 B)


@dataclass
class ParamSpec:
    name: str
    shape: tuple
    group: str = "llm"            # llm | vision | projector | action_head | lm_head
    compute: str = "bf16"         # dtype the kernels read: "bf16" (region A) or "fp32" (region B)
    fuse: Optional[str] = None    # consecutive specs with the same tag are packed without padding
    trainable: bool = True
    no_decay: Optional[bool] = None   # default: the reference's rule (LayerNorm children and *bias* names)

    @property
    def numel(self) -> int:
        n = 1
        for s in self.shape:
            n *= int(s)
        return n


@dataclass
class _Slot:
    spec: ParamSpec
    region: str
    offset: int       # element offset inside the region
    g_offset: int     # element offset in the global (A then B) fp32 master / moment buffers


# nn.LayerNorm children of the reference's models, by module name (RMSNorm modules — input_layernorm,
# post_attention_layernorm, norm — are NOT nn.LayerNorm and are therefore decayed by base_exp.py:102-103)
_LAYERNORM_PARENTS = {"layer_norm1", "layer_norm2", "pre_layrnorm", "post_layernorm", "layernorm", "attn_norm",
                      "ffn_norm", "norm3"}


def _is_layernorm_weight(name: str) -> bool:
    parts = name.split(".")
    if len(parts) < 2 or parts[-1] != "weight":
        return False
    if parts[-2] in _LAYERNORM_PARENTS:
        return True
    return len(parts) >= 4 and parts[-4] == "mlp_resnet_blocks" and parts[-2] == "0"   # ffn = Sequential(LayerNorm, ...)


def split_segments(segs: list, chunk_bounds: list):
    """Partition optimizer segments (a, b, lr, wd, region) at the chunk boundaries (region-A element ranges, disjoint,
    listed in FORWARD order, None = nothing trainable).  Returns (rest, per_chunk): every element of every segment
    lands in exactly one piece; region-B segments are never split."""
    rest, per_chunk = [], [[] for _ in chunk_bounds]
    for a, b, lr, wd, region in segs:
        if region != "A":
            rest.append((a, b, lr, wd, region))
            continue
        hits = sorted((max(a, c[0]), min(b, c[1]), i) for i, c in enumerate(chunk_bounds)
                      if c is not None and max(a, c[0]) < min(b, c[1]))
        cur = a
        for lo, hi, i in hits:
            if cur < lo:
                rest.append((cur, lo, lr, wd, region))
            per_chunk[i].append((lo, hi, lr, wd, region))
            cur = hi
        if cur < b:
            rest.append((cur, b, lr, wd, region))
    return rest, per_chunk


class ParamStore:
    """Regions: A = trainable, bf16 compute; B = trainable, fp32 compute; FA / FB = frozen counterparts
    (master + shadow only: no gradient, no optimizer state — e.g. lm_head in CogACT, or a frozen tower)."""

    # True (set by the launcher before the model is built, multi-GPU only): the bf16 gradient and weight buffers are
    # allocated as symmetric memory, so that peers can be mapped into them and the data-parallel exchange can run on
    # the copy engines over NVLink (parallel.ShardedDataParallel, transport "ce") instead of in SM-resident kernels.
    SYMMETRIC = False

    def _alloc_exchanged(self, n: int, dtype) -> torch.Tensor:
        if ParamStore.SYMMETRIC and self.device.type == "cuda" and n > 0:
            import torch.distributed._symmetric_memory as symm_mem
            t = symm_mem.empty(n, dtype=dtype, device=self.device)
            t.zero_()
            self.symmetric = True
            return t
        return torch.zeros(n, device=self.device, dtype=dtype)

    def __init__(self, specs: list[ParamSpec], device):
        self.device = torch.device(device)
        self.symmetric = False
        self.slots: dict[str, _Slot] = {}
        self.order: list[str] = []
        size = {"A": 0, "B": 0, "FA": 0, "FB": 0}
        prev_fuse = {k: None for k in size}
        placed = []
        for sp in specs:
            region = ("A" if sp.compute == "bf16" else "B") if sp.trainable else ("FA" if sp.compute == "bf16" else "FB")
            if not (sp.fuse is not None and sp.fuse == prev_fuse[region]):
                size[region] = (size[region] + ALIGN - 1) // ALIGN * ALIGN
            placed.append((sp, region, size[region]))
            size[region] += sp.numel
            prev_fuse[region] = sp.fuse
        al = lambda n: (n + ALIGN - 1) // ALIGN * ALIGN  # noqa: E731
        self.n_a, self.n_b, self.n_fa, self.n_fb = al(size["A"]), al(size["B"]), al(size["FA"]), al(size["FB"])
        base = {"A": 0, "B": self.n_a, "FA": self.n_a + self.n_b, "FB": self.n_a + self.n_b + self.n_fa}
        for sp, region, off in placed:
            assert sp.name not in self.slots, f"duplicate parameter {sp.name}"
            self.slots[sp.name] = _Slot(sp, region, off, base[region] + off)
            self.order.append(sp.name)
        dev = self.device
        n_train = self.n_a + self.n_b
        # one fp32 master for everything; trainable tensors first so moments / AdamW cover a prefix
        self.master = torch.zeros(n_train + self.n_fa + self.n_fb, device=dev, dtype=torch.float32)
        self.n_train = n_train
        self.shadow = self._alloc_exchanged(self.n_a, torch.bfloat16)
        self.shadow_f = torch.zeros(self.n_fa, device=dev, dtype=torch.bfloat16)
        self.grad_a = self._alloc_exchanged(self.n_a, torch.bfloat16)
        self.grad_b = torch.zeros(self.n_b, device=dev, dtype=torch.float32)
        self.exp_avg: Optional[torch.Tensor] = None
        self.exp_avg_sq: Optional[torch.Tensor] = None
        self.step_count = 0
        self._written: set[int] = set()       # grad tensors (by data_ptr) written since zero_grad
        self._written_ranges: list[tuple[int, int]] = []   # the same, as element ranges of grad_a
        # trainable region-A tensors as sorted element ranges: finalize_grads() zeroes the ones no kernel wrote
        self._a_slots = sorted((s.offset, s.offset + s.spec.numel) for s in self.slots.values() if s.region == "A")
        self._always_zero: list[tuple[int, int]] = []   # region-A ranges that need an explicit memset per step
        self.grad_ready_hook = None            # callable(start, end): grad_a[start:end] is final — data-parallel overlap
        self.zero_grad_hook = None             # callable(): a new step starts (the overlap resets its bookkeeping)
        self.sharder = None                    # parallel.ShardedDataParallel (ZeRO-1): owns the optimizer step if set
        # Optional overlap of the optimizer with the NEXT step's forward: AdamW is HBM-bound, the forward GEMMs are
        # tensor-bound, so the per-block updates run on a side stream in forward order and block i's forward waits
        # only for its own event (set_param_chunks / wait_chunk).  Off by default: every reader of the weights has to
        # go through wait_chunk / wait_all_params.
        self.async_optimizer = False
        self._opt_stream = None
        self._chunk_bounds: list[tuple[int, int]] = []     # [a, b) in region-A coordinates, forward order
        self._chunk_events: dict[int, "torch.cuda.Event"] = {}
        self._opt_keepalive = None

    # ------------------------------------------------------------------ views
    def _view(self, buf: torch.Tensor, off: int, shape) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= int(s)
        return buf[off:off + n].view(*shape)

    def master_view(self, name: str) -> torch.Tensor:
        s = self.slots[name]
        return self._view(self.master, s.g_offset, s.spec.shape)

    def w(self, name: str) -> torch.Tensor:
        """The tensor the kernels read (bf16 shadow for region A, fp32 master for region B)."""
        s = self.slots[name]
        if s.region == "A":
            return self._view(self.shadow, s.offset, s.spec.shape)
        if s.region == "FA":
            return self._view(self.shadow_f, s.offset, s.spec.shape)
        return self._view(self.master, s.g_offset, s.spec.shape)

    def g(self, name: str) -> Optional[torch.Tensor]:
        s = self.slots[name]
        if not s.spec.trainable:
            return None
        buf = self.grad_a if s.region == "A" else self.grad_b
        return self._view(buf, s.offset, s.spec.shape)

    def _fused(self, names: list[str], getter) -> Optional[torch.Tensor]:
        first = self.slots[names[0]]
        rows, off = 0, first.offset
        tail = first.spec.shape[1:]
        for n in names:
            s = self.slots[n]
            assert s.region == first.region and s.offset == off and s.spec.shape[1:] == tail, \
                f"{names} are not packed contiguously (fuse tag missing?)"
            off += s.spec.numel
            rows += s.spec.shape[0]
        if getter == "w":
            buf = {"A": self.shadow, "FA": self.shadow_f}.get(first.region, self.master)
            base = first.offset if first.region in ("A", "FA") else first.g_offset
        else:
            if not first.spec.trainable:
                return None
            buf = self.grad_a if first.region == "A" else self.grad_b
            base = first.offset
        return self._view(buf, base, (rows,) + tuple(tail))

    def grad_range(self, names: list[str]) -> Optional[tuple]:
        """[start, end) element range in grad_a spanned by the trainable region-A tensors in `names`."""
        sl = [self.slots[n] for n in names if n in self.slots and self.slots[n].region == "A"]
        if not sl:
            return None
        a = min(s.offset for s in sl)
        b = max(s.offset + s.spec.numel for s in sl)
        return (a // ALIGN * ALIGN, min(self.n_a, (b + ALIGN - 1) // ALIGN * ALIGN))

    def fused_w(self, names: list[str]) -> torch.Tensor:
        return self._fused(names, "w")

    def fused_g(self, names: list[str]) -> Optional[torch.Tensor]:
        return self._fused(names, "g")

    # ---------------------------------------------------------------- gradients
    def first_write(self, g: torch.Tensor) -> bool:
        """True the first time a gradient tensor is written after zero_grad (overwrite instead of accumulate).
        Region-B (fp32) gradients are memset by zero_grad and always accumulate."""
        if g.dtype == torch.float32:
            return False
        key = g.data_ptr()
        if key in self._written:
            return False
        self._written.add(key)
        a = (key - self.grad_a.data_ptr()) // 2
        self._written_ranges.append((a, a + g.numel()))
        return True

    def finalize_grads(self) -> int:
        """Region-A gradients are overwritten by the first wgrad of a step instead of being memset (15 GB at 7B), so a
        trainable tensor that no kernel reached this step (a branch not taken, a forward without its head) would keep
        the previous step's gradient.  The reference leaves such parameters at grad=None and AdamW skips them; here
        they are zeroed before the norm / optimizer read them.  Returns the number of tensors zeroed (normally 0)."""
        if not self._a_slots:
            return 0
        import bisect
        merged: list[list[int]] = []
        for a, b in sorted(self._written_ranges):
            if merged and a <= merged[-1][1]:
                merged[-1][1] = max(merged[-1][1], b)
            else:
                merged.append([a, b])
        starts = [m[0] for m in merged]
        zeroed = 0
        for a, b in self._a_slots:
            i = bisect.bisect_right(starts, a) - 1
            if i >= 0 and merged[i][1] >= b:
                continue
            self.grad_a[a:b].zero_()
            zeroed += 1
        return zeroed

    def mark_sparse_grad(self, name: str) -> None:
        """Gradient rows written by scatter (embedding table): the writer calls begin_sparse_write() first."""
        s = self.slots[name]
        assert s.region == "A"
        self._always_zero.append((s.offset, s.offset + s.spec.numel))

    def begin_sparse_write(self, g: Optional[torch.Tensor]) -> None:
        """Rows of a scatter-written gradient that receive nothing must read as zero: the first writer of a step memsets
        the tensor.  This happens in backward — not in zero_grad(): with `async_optimizer` the previous step's AdamW of
        the table may still be reading this gradient on the side stream when zero_grad() runs, whereas the forward has
        waited for that chunk (wait_chunk) long before the first backward kernel."""
        if g is not None and self.first_write(g):
            g.zero_()

    # ---------------------------------------------------------------- optimizer / forward overlap
    def set_param_chunks(self, bounds: list) -> None:
        """Region-A element ranges (e.g. one per decoder block, BlockW.grad_range) whose AdamW update may still be in
        flight when the next forward starts; the consumer calls wait_chunk(i) right before reading chunk i."""
        self._chunk_bounds = [None if b is None else tuple(b) for b in bounds]    # None: nothing trainable there

    def wait_chunk(self, i: int) -> None:
        ev = self._chunk_events.pop(i, None)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)

    def wait_all_params(self) -> None:
        for i in list(self._chunk_events):
            self.wait_chunk(i)

    def zero_grad(self) -> None:
        self._written.clear()
        self._written_ranges.clear()
        self.grad_b.zero_()
        if self.zero_grad_hook is not None:
            self.zero_grad_hook()

    def scratch_f32(self, n: int, tag: str = "") -> torch.Tensor:
        """Zeroed fp32 scratch (norm-weight / bias gradient accumulators)."""
        return torch.zeros(n, device=self.device, dtype=torch.float32)

    def accumulate_small(self, scratch_f32: torch.Tensor, g: Optional[torch.Tensor]) -> None:
        if g is None:
            return
        ops.cast_add_(scratch_f32, g.reshape(-1), accumulate=not self.first_write(g))

    # ---------------------------------------------------------------- optimizer
    def refresh_shadow(self) -> None:
        """bf16 shadow <- fp32 master (after loading weights)."""
        if self.n_a:
            ops.cast_(self.master[: self.n_a], self.shadow)
        if self.n_fa:
            a = self.n_a + self.n_b
            ops.cast_(self.master[a: a + self.n_fa], self.shadow_f)

    def segments(self, lrs: dict, weight_decay: float):
        """Contiguous (start, end, lr, wd) runs over the global master buffer, trainable tensors only."""
        runs = []
        for name in self.order:
            s = self.slots[name]
            if not s.spec.trainable:
                continue
            lr = lrs.get(s.spec.group, lrs["llm"])
            # base_exp.py:102-103: everything is decayed except nn.LayerNorm children and names containing "bias"
            # (RMSNorm weights, class_embedding, positional_embedding ... ARE decayed there); specs of LayerNorm
            # weights carry no_decay=True
            nd = s.spec.no_decay if s.spec.no_decay is not None else ("bias" in name or _is_layernorm_weight(name))
            wd = 0.0 if nd else weight_decay
            a, b = s.g_offset, s.g_offset + s.spec.numel
            b_al = (b + ALIGN - 1) // ALIGN * ALIGN
            if runs and runs[-1][2] == lr and runs[-1][3] == wd and runs[-1][4] == s.region and runs[-1][1] >= a:
                runs[-1][1] = b_al
            else:
                runs.append([a, b_al, lr, wd, s.region])
        return runs

    def grad_norm_sq(self, out: torch.Tensor) -> torch.Tensor:
        if self.n_a:
            ops.sumsq_(self.grad_a, out)
        if self.n_b:
            ops.sumsq_(self.grad_b, out)
        return out

    def adamw_step(self, lrs: dict, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                   max_grad_norm: Optional[float] = 1.0, grad_scale: float = 1.0):
        """One fused AdamW step over every trainable segment; returns the (device) gradient norm."""
        if self.sharder is not None and self.sharder.enabled:
            return self.sharder.step(lrs, betas, eps, weight_decay, max_grad_norm)
        dev = self.device
        if self.exp_avg is None:
            self.exp_avg = torch.zeros(self.n_train, device=dev, dtype=torch.float32)
            self.exp_avg_sq = torch.zeros(self.n_train, device=dev, dtype=torch.float32)
        self.step_count += 1
        self.finalize_grads()
        clip = None
        norm = torch.zeros((), device=dev, dtype=torch.float32)
        if max_grad_norm is not None:
            ssq = torch.zeros((), device=dev, dtype=torch.float32)
            self.grad_norm_sq(ssq)
            clip = torch.empty((), device=dev, dtype=torch.float32)
            ops.clip_coef(ssq, max_grad_norm, clip, norm)
        def update(a, b, lr, wd, region):
            p = self.master[a:b]
            if region == "A":
                g = self.grad_a[a:b]
                sh = self.shadow[a:b]
            else:
                g = self.grad_b[a - self.n_a:b - self.n_a]
                sh = None
            ops.adamw_(p, g, self.exp_avg[a:b], self.exp_avg_sq[a:b], sh, lr, betas[0], betas[1], eps, wd,
                       self.step_count, clip)

        segs = self.segments(lrs, weight_decay)
        if not (self.async_optimizer and any(b is not None for b in self._chunk_bounds)):
            self.wait_all_params()
            for seg in segs:
                update(*seg)
            return norm

        # split every segment at the chunk boundaries: `rest` (embeddings, towers, heads ...) first, then the chunks in
        # forward order, all on the side stream; the caller's stream waits for `rest` only
        rest, per_chunk = split_segments(segs, self._chunk_bounds)
        main = torch.cuda.current_stream()
        if self._opt_stream is None:
            self._opt_stream = torch.cuda.Stream(device=dev)
        side = self._opt_stream
        self.wait_all_params()
        ready = torch.cuda.Event()
        ready.record(main)
        side.wait_event(ready)
        self._opt_keepalive = (clip, norm)            # read by side-stream kernels after this function returns
        with torch.cuda.stream(side):
            for seg in rest:
                update(*seg)
            ev_rest = torch.cuda.Event()
            ev_rest.record(side)
            for i, pieces in enumerate(per_chunk):
                for seg in pieces:
                    update(*seg)
                ev = torch.cuda.Event()
                ev.record(side)
                self._chunk_events[i] = ev
        main.wait_event(ev_rest)
        return norm

    def bytes_allocated(self) -> int:
        n = (self.master.numel() * 4 + (self.shadow.numel() + self.shadow_f.numel()) * 2 + self.grad_a.numel() * 2 +
             self.grad_b.numel() * 4)
        if self.exp_avg is not None:
            n += 2 * self.exp_avg.numel() * 4
        return n
