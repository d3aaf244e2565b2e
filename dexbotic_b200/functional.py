"""autograd.Function layer over the CUDA kernels.

Every Function has a hand-written backward that (a) launches the dgrad kernels for its inputs and
(b) writes parameter gradients straight into the flat gradient buffers of the ParamStore (side effect;
the Function returns None for parameter handles).  Tiny glue between Functions (cat / where / repeat on
action-head tensors of a few MB) is left to torch autograd.

TransformerBlockFn keeps ONLY its input: the whole block is recomputed inside backward — the same
memory/compute trade the reference makes with gradient_checkpointing=True (dexbotic/exp/base_exp.py:245,
trainer.py:120), without torch.utils.checkpoint's replay overhead.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from . import ops
from .params import ParamStore


# ------------------------------------------------------------------------------- handles
@dataclass
class Lin:
    """nn.Linear weights as the kernels see them: w [out, in] (+ bias) and their gradient slots."""
    w: torch.Tensor
    b: Optional[torch.Tensor] = None
    gw: Optional[torch.Tensor] = None
    gb: Optional[torch.Tensor] = None

    @staticmethod
    def of(store: ParamStore, w_names, b_names=None) -> "Lin":
        w_names = [w_names] if isinstance(w_names, str) else list(w_names)
        b_names = None if b_names is None else ([b_names] if isinstance(b_names, str) else list(b_names))
        w = store.fused_w(w_names)
        gw = store.fused_g(w_names)
        b = gb = None
        if b_names is not None:
            b = store.fused_w(b_names)
            gb = store.fused_g(b_names)
        return Lin(w, b, gw, gb)


@dataclass
class Norm:
    kind: str                      # "rms" | "rms1p" (Gemma 1+w) | "ln" | "ln_noaffine"
    eps: float
    w: Optional[torch.Tensor] = None
    b: Optional[torch.Tensor] = None
    gw: Optional[torch.Tensor] = None
    gb: Optional[torch.Tensor] = None


@dataclass
class BlockCfg:
    d: int
    heads: int
    kv_heads: int
    head_dim: int
    inter: int
    mlp: str = "glu"               # "glu": down(act(gate x) * up x);  "mlp": fc2(act(fc1 x))
    act: str = "silu"
    rope: bool = True


@dataclass
class BlockW:
    cfg: BlockCfg
    norm1: Norm
    qkv: Lin
    o: Lin
    norm2: Norm
    gate: Optional[Lin] = None
    up: Optional[Lin] = None
    down: Optional[Lin] = None
    fc1: Optional[Lin] = None
    fc2: Optional[Lin] = None
    grad_range: Optional[tuple] = None   # [start, end) element range of this block's gradients in ParamStore.grad_a


@dataclass
class AttnEnv:
    """Per-forward attention environment shared by a stack of blocks."""
    B: int
    S: int
    keymask: Optional[torch.Tensor] = None   # uint8 [B,S]
    bid: Optional[torch.Tensor] = None       # int32 [B,S] (pi0 block ids: allowed iff bid[k] <= bid[q])
    causal: bool = False                     # plain causal order by index (decoder LLM)
    pos: Optional[torch.Tensor] = None       # int32 [B*S] RoPE positions
    cos: Optional[torch.Tensor] = None       # fp32 [n_pos, hd/2]
    sin: Optional[torch.Tensor] = None


# ------------------------------------------------------------------------- raw fwd / bwd
def linear_fwd(x2d, lin: Lin, act=None, residual=None, want_aux=False, out=None):
    aux = torch.empty((x2d.shape[0], lin.w.shape[0]), device=x2d.device, dtype=x2d.dtype) if want_aux else None
    y = ops.gemm(x2d, lin.w, bias=lin.b, residual=residual, act=act or 0, aux=aux, out=out)
    return y, aux


def linear_wgrad(store: ParamStore, dy2d, x2d, lin: Lin):
    """dW (+)= dy^T x ; db (+)= colsum(dy)."""
    if lin.gw is not None:
        first = store.first_write(lin.gw)
        ops.gemm(dy2d, x2d, a_mn=True, b_mn=True, out=lin.gw, residual=None if first else lin.gw)
    if lin.b is not None and lin.gb is not None:
        if lin.gb.dtype == torch.float32 and dy2d.shape[1] % 8 == 0:
            ops.colsum_(dy2d, lin.gb)
        elif dy2d.shape[1] % 8 == 0:
            sc = store.scratch_f32(lin.gb.numel())
            ops.colsum_(dy2d, sc)
            store.accumulate_small(sc, lin.gb)
        else:   # tiny odd widths (DiT final layer N=7, SE excite C/16): plain torch reduction on a few KB
            store.accumulate_small(dy2d.float().sum(0), lin.gb)


def linear_dgrad(dy2d, lin: Lin, out=None, residual=None):
    return ops.gemm(dy2d, lin.w, b_mn=True, out=out, residual=residual)


def norm_fwd(x, n: Norm, out=None):
    if n.kind in ("rms", "rms1p"):
        y, rstd = ops.rmsnorm_fwd(x, n.w, n.eps, n.kind == "rms1p", out=out)
        return y, (rstd,)
    y, mean, rstd = ops.layernorm_fwd(x, n.w, n.b, n.eps, out=out)
    return y, (mean, rstd)


def norm_bwd(store: ParamStore, dy, x, n: Norm, stats, dx=None, accumulate_dx=False):
    D = x.shape[-1]
    if n.kind in ("rms", "rms1p"):
        dw = store.scratch_f32(D) if n.gw is not None else None
        dx = ops.rmsnorm_bwd(dy, x, n.w, stats[0], n.kind == "rms1p", dx=dx, dw=dw, accumulate_dx=accumulate_dx)
        if dw is not None:
            store.accumulate_small(dw, n.gw)
        return dx
    dw = store.scratch_f32(D) if (n.w is not None and n.gw is not None) else None
    db = store.scratch_f32(D) if (n.b is not None and n.gb is not None) else None
    dx = ops.layernorm_bwd(dy, x, n.w, stats[0], stats[1], dx=dx, dw=dw, db=db, accumulate_dx=accumulate_dx)
    if dw is not None:
        store.accumulate_small(dw, n.gw)
    if db is not None:
        store.accumulate_small(db, n.gb)
    return dx


def block_forward(x2d: torch.Tensor, bw: BlockW, env: AttnEnv, mode: str):
    """One pre-norm transformer block on [B*S, d].

    mode "out":     forward only, nothing kept (the block is recomputed in backward);
    mode "saved":   backward-time recompute: every intermediate backward needs, no block output
                    (the down / fc2 GEMM is skipped);
    mode "both":    forward that also keeps the intermediates (no recompute later; +1.2 GB per 7B layer).
    The normed activations h / h2 are never stored: re-normalising is an HBM-bound 2-pass over [M, d]."""
    c = bw.cfg
    keep = mode != "out"
    sh = ops.AttnShape(env.B, env.S, c.heads, c.kv_heads, c.head_dim, x2d.dtype)
    h, st1 = norm_fwd(x2d, bw.norm1)
    qkv, _ = linear_fwd(h, bw.qkv)
    if c.rope:
        ops.rope_(qkv, env.pos, env.cos, env.sin, c.heads + c.kv_heads, c.head_dim)
    attn, probs = ops.attention_fwd(qkv.view(env.B, env.S, -1), sh, keymask=env.keymask, bid_q=env.bid, bid_k=env.bid,
                                    causal=env.causal)
    attn2d = attn.view(x2d.shape[0], -1)
    x1, _ = linear_fwd(attn2d, bw.o, residual=x2d)
    h2, st2 = norm_fwd(x1, bw.norm2)
    saved, y = None, None
    if c.mlp == "glu":
        fused = (bw.gate.b is None and bw.up.b is None
                 and ops.glu_fusable(h2, bw.gate.w, bw.up.w, bw.down.w, c.act))
        if fused:
            # ONE GEMM against gate and up (CTA pair: rank 0 stages gate columns, rank 1 up columns); its epilogue
            # writes act(g) * u and, for backward, the two pre-activations.  hm is kept: the backward's down-projection
            # dgrad produces dg / du in its epilogue and no longer recomputes it.
            g = torch.empty((h2.shape[0], bw.gate.w.shape[0]), device=h2.device, dtype=h2.dtype) if keep else None
            u = torch.empty_like(g) if keep else None
            hm = ops.gemm_dual(h2, bw.gate.w, bw.up.w, c.act, aux_gate=g, aux_up=u)
        else:
            g, _ = linear_fwd(h2, bw.gate)
            u, _ = linear_fwd(h2, bw.up)
            hm = ops.glu_fwd(g, u, c.act) if mode != "saved" else None
        if mode != "saved":
            y, _ = linear_fwd(hm, bw.down, residual=x1)
        if keep:
            saved = dict(st1=st1, qkv=qkv, probs=probs, attn=attn2d, x1=x1, st2=st2, g=g, u=u, sh=sh,
                         hm=hm if fused else None)
    else:
        hm, pre = linear_fwd(h2, bw.fc1, act=c.act, want_aux=keep)
        if mode != "saved":
            y, _ = linear_fwd(hm, bw.fc2, residual=x1)
        if keep:
            saved = dict(st1=st1, qkv=qkv, probs=probs, attn=attn2d, x1=x1, st2=st2, pre=pre, sh=sh)
    if mode == "saved":       # recompute pass: h / h2 are already here, hand them over instead of re-normalising
        saved["h"], saved["h2"] = h, h2
    return y, saved


def block_backward(store: ParamStore, dy: torch.Tensor, x2d: torch.Tensor, bw: BlockW, env: AttnEnv, s: dict):
    """Gradient of block_forward; dy is consumed (reused as the residual-stream gradient buffer)."""
    c = bw.cfg
    h2 = s.get("h2")
    if h2 is None:
        h2, _ = norm_fwd(s["x1"], bw.norm2)
    # ---- MLP
    if c.mlp == "glu":
        if s.get("hm") is not None:
            # fused: dgrad of the down projection with the GLU backward in its epilogue (dg / du in place over g / u)
            linear_wgrad(store, dy, s["hm"], bw.down)
            dg, du = ops.gemm_glu_bwd(dy, bw.down.w, s["g"], s["u"], c.act, dg=s["g"], du=s["u"])
        else:
            dhm = linear_dgrad(dy, bw.down)                                  # [M, inter]
            # one pass: dg, du (in place over g / u) and hm = act(g)*u (in place over dhm) for the down wgrad
            dg, du = ops.glu_bwd(dhm, s["g"], s["u"], c.act, dg=s["g"], du=s["u"], h_out=dhm)
            linear_wgrad(store, dy, dhm, bw.down)
        linear_wgrad(store, dg, h2, bw.gate)
        linear_wgrad(store, du, h2, bw.up)
        dh2 = linear_dgrad(dg, bw.gate)
        linear_dgrad(du, bw.up, out=dh2, residual=dh2)
    else:
        hm = ops.act_fwd(s["pre"], c.act)
        linear_wgrad(store, dy, hm, bw.fc2)
        dhm = linear_dgrad(dy, bw.fc2, out=hm)
        dpre = ops.act_bwd(dhm, s["pre"], c.act, out=dhm)
        linear_wgrad(store, dpre, h2, bw.fc1)
        dh2 = linear_dgrad(dpre, bw.fc1)
    dx1 = norm_bwd(store, dh2, s["x1"], bw.norm2, s["st2"], dx=dy, accumulate_dx=True)   # dy += norm2 bwd
    # ---- attention
    linear_wgrad(store, dx1, s["attn"], bw.o)
    dattn = linear_dgrad(dx1, bw.o)
    dqkv = ops.attention_bwd(dattn.view(env.B, env.S, -1), s["qkv"].view(env.B, env.S, -1), s["probs"], s["sh"],
                             causal=env.causal, out=s["attn"], keymask=env.keymask, bid_q=env.bid, bid_k=env.bid)
    dqkv2d = dqkv.view(x2d.shape[0], -1)
    if c.rope:
        ops.rope_(dqkv2d, env.pos, env.cos, env.sin, c.heads + c.kv_heads, c.head_dim, inverse=True)
    h = s.get("h")
    if h is None:
        h, _ = norm_fwd(x2d, bw.norm1)
    linear_wgrad(store, dqkv2d, h, bw.qkv)
    dh = linear_dgrad(dqkv2d, bw.qkv)
    return norm_bwd(store, dh, x2d, bw.norm1, s["st1"], dx=dx1, accumulate_dx=True)


# --------------------------------------------------------------------- autograd Functions
class TransformerBlockFn(torch.autograd.Function):
    """recompute=True: keep only the block input and recompute inside backward (the reference's
    gradient_checkpointing=True trade, base_exp.py:245); recompute=False: keep the intermediates."""

    @staticmethod
    def forward(ctx, x2d, bw: BlockW, env: AttnEnv, store: ParamStore, recompute: bool = True):
        y, saved = block_forward(x2d, bw, env, "out" if recompute else "both")
        ctx.save_for_backward(x2d)
        ctx.bw, ctx.env, ctx.store, ctx.kept = bw, env, store, saved
        return y

    @staticmethod
    def backward(ctx, dy):
        (x2d,) = ctx.saved_tensors
        saved = ctx.kept
        ctx.kept = None
        if saved is None:
            _, saved = block_forward(x2d, ctx.bw, ctx.env, "saved")          # recompute
        # block_backward accumulates the residual-stream gradient in place: work on a private buffer unless
        # the incoming gradient is a whole, contiguous tensor nobody else can be holding a view of
        dy = dy.contiguous() if dy._base is None and dy.is_contiguous() else dy.clone(memory_format=torch.contiguous_format)
        dx = block_backward(ctx.store, dy, x2d, ctx.bw, ctx.env, saved)
        hook = ctx.store.grad_ready_hook
        if hook is not None and ctx.bw.grad_range is not None:
            hook(*ctx.bw.grad_range)       # this block's gradients are final: data-parallel all-reduce may start
        return dx, None, None, None, None


class LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2d, lin: Lin, act, store: ParamStore, need_dx: bool, anchor=None):
        # `anchor` (a scalar that requires grad) makes autograd run this backward when x2d itself does not
        # require grad (first layer after a data tensor): parameter gradients are a side effect.
        y, aux = linear_fwd(x2d, lin, act=act, want_aux=act is not None and act != "none")
        ctx.save_for_backward(x2d, aux) if aux is not None else ctx.save_for_backward(x2d)
        ctx.lin, ctx.act, ctx.store, ctx.need_dx = lin, act, store, need_dx
        return y

    @staticmethod
    def backward(ctx, dy):
        saved = ctx.saved_tensors
        x2d = saved[0]
        dy = dy.contiguous()
        if len(saved) > 1:
            dy = ops.act_bwd(dy, saved[1], ctx.act)
        linear_wgrad(ctx.store, dy, x2d, ctx.lin)
        dx = linear_dgrad(dy, ctx.lin) if ctx.need_dx else None
        return dx, None, None, None, None, None


class NormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2d, n: Norm, store: ParamStore):
        y, stats = norm_fwd(x2d, n)
        ctx.save_for_backward(x2d, *stats)
        ctx.n, ctx.store = n, store
        return y

    @staticmethod
    def backward(ctx, dy):
        x2d, *stats = ctx.saved_tensors
        return norm_bwd(ctx.store, dy.contiguous(), x2d, ctx.n, tuple(stats)), None, None


class GatherRowsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2d, idx):
        ctx.save_for_backward(idx)
        ctx.shape = x2d.shape
        return ops.gather_rows(x2d, idx)

    @staticmethod
    def backward(ctx, dout):
        (idx,) = ctx.saved_tensors
        dx = torch.zeros(ctx.shape, device=dout.device, dtype=dout.dtype)
        ops.scatter_rows_add_(dout.contiguous(), idx, dx)
        return dx, None


class SpliceFn(torch.autograd.Function):
    """inputs_embeds = splice(embed_tokens.weight, image features)  (dexbotic_arch.py:182-373)."""

    @staticmethod
    def forward(ctx, feats2d, src, table, g_table, store: ParamStore):
        ctx.save_for_backward(src)
        ctx.g_table, ctx.store, ctx.fshape = g_table, store, feats2d.shape
        return ops.splice_gather(src, table, feats2d)

    @staticmethod
    def backward(ctx, dout):
        (src,) = ctx.saved_tensors
        d_feats = torch.zeros(ctx.fshape, device=dout.device, dtype=dout.dtype)
        ctx.store.begin_sparse_write(ctx.g_table)
        ops.splice_scatter(src, dout.contiguous(), ctx.g_table, d_feats)
        return d_feats, None, None, None, None


class MSELossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target):
        pred, target = pred.contiguous(), target.contiguous()
        out = torch.zeros((), device=pred.device, dtype=torch.float32)
        ops.mse_fwd(pred, target, out)
        ctx.save_for_backward(pred, target)
        return out

    @staticmethod
    def backward(ctx, g):
        pred, target = ctx.saved_tensors
        return ops.mse_bwd(pred, target, g.contiguous().float()), None


class CastFn(torch.autograd.Function):
    """dtype boundary between the bf16 trunk and the fp32 action head (cogact_arch.py:125-135)."""

    @staticmethod
    def forward(ctx, x, dtype):
        ctx.src_dtype = x.dtype
        out = torch.empty(x.shape, device=x.device, dtype=dtype)
        return ops.cast_(x.contiguous(), out)

    @staticmethod
    def backward(ctx, g):
        out = torch.empty(g.shape, device=g.device, dtype=ctx.src_dtype)
        return ops.cast_(g.contiguous(), out), None


class CrossEntropyFn(torch.autograd.Function):
    """mean cross entropy over rows with label != -100, fp32 math on bf16/fp32 logits
    (F.cross_entropy under fp32 autocast: oft_discrete_arch.py:169-191; dexbotic_arch.py:489)."""

    @staticmethod
    def forward(ctx, logits2d, labels):
        logits2d = logits2d.contiguous()
        loss_sum, n_valid, lse = ops.cross_entropy_fwd(logits2d, labels.contiguous())
        ctx.save_for_backward(logits2d, labels, lse, n_valid)
        return loss_sum / n_valid.clamp(min=1).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        logits2d, labels, lse, n_valid = ctx.saved_tensors
        return ops.cross_entropy_bwd(logits2d, labels.contiguous(), lse, n_valid, g.contiguous().float()), None


# ------------------------------------------------------------------ MemVLA memory path (memvla_arch.py)
class CrossAttnFn(torch.autograd.Function):
    """softmax(Q K^T / sqrt(hd)) V, no mask, separate query / key lengths, optional dropout on the attention weights
    (F.scaled_dot_product_attention(q, k, v, dropout_p) at memvla_arch.py:122-124; nn.MultiheadAttention's core at
    memvla/action_model/dit.py:181).  q [B*Sq, D], k / v [B*Sk, D]."""

    @staticmethod
    def forward(ctx, q, k, v, B, Sq, Sk, H, dropout_p=0.0, seed=0):
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        out, probs, pd = ops.cross_attention_fwd(q, k, v, B, Sq, Sk, H, dropout_p=dropout_p, seed=seed)
        ctx.save_for_backward(q, k, v, probs) if pd is None else ctx.save_for_backward(q, k, v, probs, pd)
        ctx.geom = (B, Sq, Sk, H, dropout_p, seed)
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, probs, *rest = ctx.saved_tensors
        B, Sq, Sk, H, p, seed = ctx.geom
        need_dq, need_dkv = ctx.needs_input_grad[0], ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        dq, dk, dv = ops.cross_attention_bwd(dout.contiguous(), q, k, v, probs, rest[0] if rest else None, B, Sq, Sk, H,
                                             dropout_p=p, seed=seed, need_dq=need_dq, need_dkv=need_dkv)
        return dq, dk, dv, None, None, None, None, None, None


class DropoutFn(torch.autograd.Function):
    """nn.Dropout (memvla_arch.py:100,102) with the library's counter-based generator; backward re-applies the mask."""

    @staticmethod
    def forward(ctx, x2d, p, seed):
        ctx.p, ctx.seed = p, seed
        return ops.dropout(x2d.contiguous(), p, seed)

    @staticmethod
    def backward(ctx, g):
        return ops.dropout(g.contiguous(), ctx.p, ctx.seed), None, None


class GateFuseFn(torch.autograd.Function):
    """GateFusion's elementwise tail (memvla_arch.py:183-192): sigmoid(z) * x1 + (1 - sigmoid(z)) * x2."""

    @staticmethod
    def forward(ctx, z, x1, x2):
        z, x1, x2 = z.contiguous(), x1.contiguous(), x2.contiguous()
        ctx.save_for_backward(z, x1, x2)
        return ops.gate_fuse_fwd(z, x1, x2)

    @staticmethod
    def backward(ctx, g):
        z, x1, x2 = ctx.saved_tensors
        return ops.gate_fuse_bwd(g.contiguous(), z, x1, x2)


class SEPoolFn(torch.autograd.Function):
    """AdaptiveAvgPool2d(1) over the token grid (memvla_arch.py:146): [B,P,C] -> [B,C] (same dtype)."""

    @staticmethod
    def forward(ctx, x3d):
        ctx.shape = x3d.shape
        return ops.se_reduce(x3d.contiguous(), None, 1.0 / x3d.shape[1]).to(x3d.dtype)

    @staticmethod
    def backward(ctx, g):
        B, P, Cc = ctx.shape
        ones = torch.ones((B, P, Cc), device=g.device, dtype=g.dtype)   # never built in SEScaleFn's fused path
        return ops.se_scale(ones, g.contiguous(), None, 0.0).mul_(1.0 / P)


class SEGateFn(torch.autograd.Function):
    """The squeeze-excite of BottleneckSE (memvla_arch.py:146-151,164-165) as ONE function of x:
        pool = mean_p x ; w = sigmoid(W2 relu(W1 pool + b1) + b2) ; out = x * w
    The two excite GEMMs have M = batch rows; their backward is chained by hand so the token map x is read twice in
    forward (pool, scale) and three times in backward (dw reduction, dx) — never materialising a broadcast."""

    @staticmethod
    def forward(ctx, x3d, l1: Lin, l2: Lin, store: ParamStore):
        x3d = x3d.contiguous()
        B, P, Cc = x3d.shape
        pool = ops.se_reduce(x3d, None, 1.0 / P).to(x3d.dtype)
        pre1, _ = linear_fwd(pool, l1)                                     # [B, C/16]: a few KB, any width
        h = torch.relu(pre1)
        z, _ = linear_fwd(h, l2)
        w = torch.sigmoid(z.float()).to(x3d.dtype)                       # [B, C]: a few KB
        out = ops.se_scale(x3d, w)
        ctx.save_for_backward(x3d, pool, pre1, h, w)
        ctx.misc = (l1, l2, store)
        return out

    @staticmethod
    def backward(ctx, dout):
        x3d, pool, pre1, h, w = ctx.saved_tensors
        l1, l2, store = ctx.misc
        dout = dout.contiguous()
        B, P, Cc = x3d.shape
        dw = ops.se_reduce(dout, x3d, 1.0)                                # [B, C] fp32 = sum_p dout * x
        wf = w.float()
        dz = (dw * wf * (1.0 - wf)).to(x3d.dtype)                         # sigmoid'
        linear_wgrad(store, dz, h, l2)
        dh = linear_dgrad(dz, l2)
        dpre = dh * (pre1 > 0).to(dh.dtype)
        linear_wgrad(store, dpre, pool, l1)
        dpool = linear_dgrad(dpre, l1)
        dx = ops.se_scale(dout, w, dpool, 1.0 / P)                        # dout * w + dpool / P
        return dx, None, None, None
