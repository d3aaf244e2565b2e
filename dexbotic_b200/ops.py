"""Tensor-level wrappers over the C-ABI (one Python call = one kernel launch on the current stream).

torch is used here for device memory and the current stream only; every arithmetic op below runs in
libdexbotic_b200.so.  Calling any of these with CPU tensors raises: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import os

import torch

from . import _lib
from ._lib import ACT_CODES, BF16, F32, GemmArgs

_INT32_MIN = -(2 ** 31)


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"dexbotic_b200 kernels take bf16 or fp32 tensors, got {t.dtype}")


def _cuda(*ts: Optional[torch.Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("dexbotic_b200 ops run on CUDA (sm_100a) tensors only; there is no CPU fallback")


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def act_code(act) -> int:
    return act if isinstance(act, int) else ACT_CODES[act]


# --------------------------------------------------------------------------- GEMM
def _tma_ok(*lds_and_dtypes) -> bool:
    for ld, dt in lds_and_dtypes:
        es = 4 if dt == F32 else 2
        if (ld * es) % 16 != 0:
            return False
    return True


def gemm(a: torch.Tensor, b: torch.Tensor, *, a_mn: bool = False, b_mn: bool = False,
         out: Optional[torch.Tensor] = None, out_dtype: Optional[torch.dtype] = None,
         bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None, act=0, alpha: float = 1.0,
         aux: Optional[torch.Tensor] = None, block_n: int = 0) -> torch.Tensor:
    """D[M,N] = act(alpha * A.B + bias) + residual.

    a: [M,K] (a_mn=False) or [K,M] (a_mn=True: M contiguous); b: [N,K] (b_mn=False, nn.Linear weight) or
    [K,N] (b_mn=True).  Inner stride must be 1.  aux (optional, like out) receives the pre-activation.
    """
    _cuda(a, b, out, bias, residual, aux)
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    M, K = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
    assert K == Kb, f"gemm: K mismatch {K} vs {Kb}"
    assert a.dtype == b.dtype
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=out_dtype or a.dtype)
    assert out.shape == (M, N) and out.stride(1) == 1
    if M == 0 or N == 0:
        return out
    g = GemmArgs()
    g.a, g.b, g.d = a.data_ptr(), b.data_ptr(), out.data_ptr()
    g.ab_dtype, g.d_dtype = _dt(a), _dt(out)
    g.a_mn_major, g.b_mn_major = int(a_mn), int(b_mn)
    g.m, g.n, g.k = M, N, K
    g.a_ld, g.b_ld, g.d_ld = a.stride(0), b.stride(0), out.stride(0)
    g.z_lo = g.z_hi = g.k_segs = 1
    g.a_z2 = g.b_z2 = 1
    g.a_div = g.a_mul = g.b_div = g.b_mul = 1
    g.alpha = float(alpha)
    if bias is not None:
        assert bias.numel() == N and bias.is_contiguous()
        g.bias, g.bias_dtype = bias.data_ptr(), _dt(bias)
    if residual is not None:
        assert residual.shape == (M, N) and residual.stride(1) == 1
        g.residual, g.res_dtype, g.res_ld = residual.data_ptr(), _dt(residual), residual.stride(0)
    if aux is not None:
        assert aux.shape == (M, N) and aux.stride(1) == 1 and aux.dtype == out.dtype
        g.aux, g.aux_ld = aux.data_ptr(), aux.stride(0)
    g.act = act_code(act)
    g.block_n = block_n
    lib = _lib.load()
    tma = _tma_ok((g.a_ld, g.ab_dtype), (g.b_ld, g.ab_dtype), (g.d_ld, g.d_dtype)) and all(
        t.data_ptr() % 16 == 0 for t in (a, b, out))
    if residual is not None:
        tma = tma and _tma_ok((g.res_ld, g.res_dtype)) and residual.data_ptr() % 16 == 0
    if bias is not None:
        tma = tma and bias.data_ptr() % 16 == 0 and N % 8 == 0
    if aux is not None:
        tma = tma and _tma_ok((g.aux_ld, g.d_dtype))
    if tma:
        _lib.check(lib.b200_gemm(C.byref(g), _stream()), "gemm")
    else:
        if aux is not None:
            # odd-stride operands (e.g. a 9-wide proprio state): pre-activation by the SIMT kernel, then the activation
            if residual is not None or not aux.is_contiguous() or not out.is_contiguous():
                raise RuntimeError("gemm: aux output on the SIMT path needs contiguous out/aux and no residual")
            g.d, g.d_ld, g.act, g.aux = aux.data_ptr(), aux.stride(0), 0, None
            _lib.check(lib.b200_gemm_simt(C.byref(g), _stream()), "gemm_simt")
            return act_fwd(aux, act, out=out)
        _lib.check(lib.b200_gemm_simt(C.byref(g), _stream()), "gemm_simt")
    return out


def gemm_dual(a: torch.Tensor, b_gate: torch.Tensor, b_up: torch.Tensor, act, *, out: Optional[torch.Tensor] = None,
              aux_gate: Optional[torch.Tensor] = None, aux_up: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = act(a @ b_gate^T) * (a @ b_up^T); optionally stores both pre-activations (for backward)."""
    _cuda(a, b_gate, b_up, out, aux_gate, aux_up)
    M, K = a.shape
    N = b_gate.shape[0]
    assert b_gate.shape == b_up.shape == (N, K) and b_gate.stride(0) == b_up.stride(0)
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=a.dtype)
    g = GemmArgs()
    g.a, g.b, g.b2, g.d = a.data_ptr(), b_gate.data_ptr(), b_up.data_ptr(), out.data_ptr()
    g.ab_dtype, g.d_dtype = _dt(a), _dt(out)
    g.m, g.n, g.k = M, N, K
    g.a_ld, g.b_ld, g.d_ld = a.stride(0), b_gate.stride(0), out.stride(0)
    g.z_lo = g.z_hi = g.k_segs = 1
    g.a_z2 = g.b_z2 = 1
    g.a_div = g.a_mul = g.b_div = g.b_mul = 1
    g.alpha = 1.0
    if aux_gate is not None:
        assert aux_up is not None and aux_gate.stride(0) == aux_up.stride(0) and aux_gate.dtype == out.dtype
        g.aux, g.aux2, g.aux_ld = aux_gate.data_ptr(), aux_up.data_ptr(), aux_gate.stride(0)
    g.act = act_code(act)
    g.dual_b = 1
    _lib.check(_lib.load().b200_gemm(C.byref(g), _stream()), "gemm_dual")
    return out


def glu_fusable(x2d: torch.Tensor, w_gate: torch.Tensor, w_up: torch.Tensor, w_down: torch.Tensor, act) -> bool:
    """True when the gate / up / down projections of a GLU MLP can run as the two fused GEMMs (gemm_dual with the
    pre-activations kept, gemm_glu_bwd): bf16, SiLU or tanh-GELU, TMA-describable operands (B200_FUSED_GLU=0: off)."""
    if os.environ.get("B200_FUSED_GLU", "1") == "0":
        return False
    if x2d.dtype != torch.bfloat16 or act_code(act) not in (2, 4):
        return False
    inter, d = w_gate.shape
    return (d % 64 == 0 and inter % 64 == 0 and w_gate.stride(0) == w_up.stride(0) and w_gate.stride(1) == 1
            and all(t.data_ptr() % 16 == 0 for t in (x2d, w_gate, w_up, w_down)) and x2d.stride(1) == 1
            and x2d.stride(0) % 8 == 0 and w_down.stride(0) % 8 == 0 and w_gate.stride(0) % 8 == 0)


def gemm_glu_bwd(dy: torch.Tensor, w_down: torch.Tensor, g: torch.Tensor, u: torch.Tensor, act, *,
                 dg: Optional[torch.Tensor] = None, du: Optional[torch.Tensor] = None):
    """(dg, du) of h = act(g) * u given dy of y = h @ w_down^T: the dgrad GEMM dh = dy @ w_down with the GLU backward in
    its epilogue (dh never reaches HBM).  dg / du may alias g / u (in-place)."""
    _cuda(dy, w_down, g, u, dg, du)
    M, d = dy.shape
    inter = w_down.shape[1]
    assert w_down.shape == (d, inter) and g.shape == u.shape == (M, inter) and g.dtype == u.dtype == dy.dtype == torch.bfloat16
    assert g.stride(1) == 1 and u.stride(1) == 1 and g.stride(0) == u.stride(0)
    dg = torch.empty_like(g) if dg is None else dg
    du = torch.empty_like(u) if du is None else du
    assert dg.stride() == g.stride() and du.stride() == g.stride()
    a = GemmArgs()
    a.a, a.b, a.d = dy.data_ptr(), w_down.data_ptr(), dg.data_ptr()
    a.ab_dtype, a.d_dtype = _dt(dy), _dt(dg)
    a.a_mn_major, a.b_mn_major = 0, 1                      # dh = dy [M, d] @ w_down [d, inter]: B is N-contiguous
    a.m, a.n, a.k = M, inter, d
    a.a_ld, a.b_ld, a.d_ld = dy.stride(0), w_down.stride(0), dg.stride(0)
    a.z_lo = a.z_hi = a.k_segs = 1
    a.a_z2 = a.b_z2 = 1
    a.a_div = a.a_mul = a.b_div = a.b_mul = 1
    a.alpha = 1.0
    a.act = act_code(act)
    a.glu_bwd, a.glu_g, a.glu_u, a.d2, a.glu_ld = 1, g.data_ptr(), u.data_ptr(), du.data_ptr(), g.stride(0)
    _lib.check(_lib.load().b200_gemm(C.byref(a), _stream()), "gemm_glu_bwd")
    return dg, du


def gemm_raw(**kw) -> None:
    """Fully general strided/batched GEMM: keyword fields of b200_gemm_args (pointers as ints)."""
    g = GemmArgs()
    for k, v in kw.items():
        setattr(g, k, v)
    _lib.check(_lib.load().b200_gemm(C.byref(g), _stream()), "gemm")


# --------------------------------------------------------------------- attention
class AttnShape:
    """Geometry of one attention call over a packed qkv buffer [B, S, (H + 2*KVH) * hd]."""

    def __init__(self, B: int, S: int, H: int, KVH: int, hd: int, dtype: torch.dtype):
        assert H % KVH == 0
        self.B, self.S, self.H, self.KVH, self.hd, self.dtype = B, S, H, KVH, hd, dtype
        self.G = H // KVH
        self.W = (H + 2 * KVH) * hd
        self.ld_s = (S + 3) // 4 * 4                       # fp32 score rows: 16-byte stride
        self.ld_p = (S + 7) // 8 * 8 if dtype == torch.bfloat16 else self.ld_s
        self.scale = hd ** -0.5


def _qkv_ptrs(qkv: torch.Tensor, sh: AttnShape):
    es = qkv.element_size()
    base = qkv.data_ptr()
    return base, base + sh.H * sh.hd * es, base + (sh.H + sh.KVH) * sh.hd * es


def _fused_scores_ok(qkv: torch.Tensor, sh: AttnShape) -> bool:
    import os
    # the persistent kernel keeps a per-CTA item list of <= 1024 (head, tile) entries: ~148 * 250 heads
    return (qkv.dtype == torch.bfloat16 and sh.hd <= 128 and sh.hd % 8 == 0 and sh.S <= 512 and sh.B * sh.H <= 30000 and
            os.environ.get("B200_FUSED_ATTN_SCORES", "1") != "0")


def attn_scores(a_ptr: int, a_ld: int, a_s_head: int, a_s_batch: int, b_ptr: int, b_ld: int, b_s_head: int,
                b_s_batch: int, sh: AttnShape, out: torch.Tensor, *, p_in: Optional[torch.Tensor] = None, keymask=None,
                bid_q=None, bid_k=None, causal: bool = False) -> torch.Tensor:
    """mode 0 (p_in None): out = softmax_mask(scale * A B^T); mode 1: out = scale * P * (A B^T - rowsum(P * A B^T))."""
    _lib.check(_lib.load().b200_attn_scores(a_ptr, b_ptr, _p(p_in), out.data_ptr(), sh.B, sh.H, sh.KVH, sh.S, sh.S,
                                            sh.hd, a_ld, a_s_head, a_s_batch, b_ld, b_s_head, b_s_batch, sh.ld_p,
                                            float(sh.scale), int(causal), _p(keymask), _p(bid_q), _p(bid_k),
                                            0 if p_in is None else 1, _stream()), "attn_scores")
    return out


def _flash_ok(qkv: torch.Tensor, sh: AttnShape) -> bool:
    import os
    return (qkv.dtype == torch.bfloat16 and sh.hd % 8 == 0 and 16 <= sh.hd <= 256 and qkv.is_contiguous() and
            os.environ.get("B200_FLASH_ATTN", "1") != "0")


def _flash_mask_ws(B: int, S: int, device, keymask, bid_k):
    """Workspace of the kernels' per-block mask summary (b200_flash_attn_fwd: mask_ws); None without key-side masks."""
    if keymask is None and bid_k is None:
        return None
    return torch.empty((B, (S + 63) // 64, 8), device=device, dtype=torch.int32)


def flash_attention_fwd(qkv: torch.Tensor, sh: AttnShape, *, keymask=None, bid_q=None, bid_k=None, causal: bool = False,
                        out: Optional[torch.Tensor] = None):
    """Flash attention over packed qkv: returns (out [B,S,H*hd], lse [B,H,S] fp32, log2 domain).  No P tensor."""
    _cuda(qkv, keymask, bid_q, bid_k)
    B, S, H, KVH, hd, W = sh.B, sh.S, sh.H, sh.KVH, sh.hd, sh.W
    if out is None:
        out = torch.empty((B, S, H * hd), device=qkv.device, dtype=qkv.dtype)
    lse = torch.empty((B, H, S), device=qkv.device, dtype=torch.float32)
    q, k, v = _qkv_ptrs(qkv, sh)
    ws = _flash_mask_ws(B, S, qkv.device, keymask, bid_k)
    _lib.check(_lib.load().b200_flash_attn_fwd(q, k, v, out.data_ptr(), lse.data_ptr(), B, H, KVH, S, hd, W, hd, S * W,
                                               H * hd, hd, S * H * hd, float(sh.scale), int(causal), _p(keymask),
                                               _p(bid_q), _p(bid_k), _p(ws), _stream()), "flash_attn_fwd")
    return out, lse


def flash_attention_bwd(dout: torch.Tensor, qkv: torch.Tensor, out: torch.Tensor, lse: torch.Tensor, sh: AttnShape, *,
                        keymask=None, bid_q=None, bid_k=None, causal: bool = False,
                        dqkv: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Gradient of flash_attention_fwd w.r.t. the packed qkv buffer (every element of dqkv is written)."""
    _cuda(dout, qkv, out, lse)
    B, S, H, KVH, hd, W = sh.B, sh.S, sh.H, sh.KVH, sh.hd, sh.W
    assert dout.is_contiguous() and dout.shape == (B, S, H * hd) and out.is_contiguous() and out.shape == dout.shape
    if dqkv is None:
        dqkv = torch.empty_like(qkv)
    delta = torch.empty((B, H, S), device=qkv.device, dtype=torch.float32)
    q, k, v = _qkv_ptrs(qkv, sh)
    dq, dk, dv = _qkv_ptrs(dqkv, sh)
    ws = _flash_mask_ws(B, S, qkv.device, keymask, bid_k)
    _lib.check(_lib.load().b200_flash_attn_bwd(q, k, v, out.data_ptr(), dout.data_ptr(), lse.data_ptr(),
                                               delta.data_ptr(), dq, dk, dv, B, H, KVH, S, hd, W, hd, S * W, H * hd, hd,
                                               S * H * hd, W, hd, S * W, float(sh.scale), int(causal), _p(keymask),
                                               _p(bid_q), _p(bid_k), _p(ws), _stream()), "flash_attn_bwd")
    return dqkv


def attention_fwd(qkv: torch.Tensor, sh: AttnShape, *, keymask: Optional[torch.Tensor] = None,
                  bid_q: Optional[torch.Tensor] = None, bid_k: Optional[torch.Tensor] = None, causal: bool = False,
                  scores: Optional[torch.Tensor] = None, probs: Optional[torch.Tensor] = None,
                  out: Optional[torch.Tensor] = None):
    """softmax(mask(Q K^T / sqrt(hd))) V over packed qkv (RoPE already applied).  Returns (out[B,S,H*hd], saved):
    saved is the fp32 log-sum-exp [B,H,S] on the flash path (bf16), the probabilities [B,H,S,ld_p] otherwise (fp32
    action heads); attention_bwd takes either."""
    _cuda(qkv)
    B, S, H, KVH, hd, G, W = sh.B, sh.S, sh.H, sh.KVH, sh.hd, sh.G, sh.W
    if _flash_ok(qkv, sh):
        return flash_attention_fwd(qkv, sh, keymask=keymask, bid_q=bid_q, bid_k=bid_k, causal=causal, out=out)
    dt = _dt(qkv)
    dev = qkv.device
    if probs is None:
        probs = torch.empty((B, H, S, sh.ld_p), device=dev, dtype=qkv.dtype)
    if out is None:
        out = torch.empty((B, S, H * hd), device=dev, dtype=qkv.dtype)
    q, k, v = _qkv_ptrs(qkv, sh)
    if _fused_scores_ok(qkv, sh):
        # scores stay in TMEM: P = softmax(scale * Q K^T) is written once, as bf16
        attn_scores(q, W, hd, S * W, k, W, hd, S * W, sh, probs, keymask=keymask, bid_q=bid_q, bid_k=bid_k,
                    causal=causal)
    else:
        if scores is None:
            scores = torch.empty((B, H, S, sh.ld_s), device=dev, dtype=torch.float32)
        gemm_raw(a=q, b=k, d=scores.data_ptr(), ab_dtype=dt, d_dtype=F32, a_mn_major=0, b_mn_major=0, m=S, n=S, k=hd,
                 a_ld=W, a_s2=hd, a_s3=S * W, a_z2=H, b_ld=W, b_s2=hd, b_s3=S * W, b_z2=KVH,
                 d_ld=sh.ld_s, d_s2=S * sh.ld_s, d_s3=H * S * sh.ld_s, z_lo=H, z_hi=B,
                 a_div=1, a_mul=1, a_seg=0, b_div=G, b_mul=1, b_seg=0, k_segs=1, alpha=sh.scale)
        softmax_fwd(scores, probs, S, S, heads=H, keymask=keymask, bid_q=bid_q, bid_k=bid_k, causal=causal)
    # out = P V
    gemm_raw(a=probs.data_ptr(), b=v, d=out.data_ptr(), ab_dtype=dt, d_dtype=dt, a_mn_major=0, b_mn_major=1,
             m=S, n=hd, k=S, a_ld=sh.ld_p, a_s2=S * sh.ld_p, a_s3=H * S * sh.ld_p, a_z2=H,
             b_ld=W, b_s2=hd, b_s3=S * W, b_z2=KVH, d_ld=H * hd, d_s2=hd, d_s3=S * H * hd, z_lo=H, z_hi=B,
             a_div=1, a_mul=1, a_seg=0, b_div=G, b_mul=1, b_seg=0, k_segs=1, alpha=1.0)
    return out, probs


def attention_cross(q_packed: torch.Tensor, kv: torch.Tensor, B: int, Sq: int, Sk: int, H: int, KVH: int, hd: int, *,
                    keymask: Optional[torch.Tensor] = None, bid_q: Optional[torch.Tensor] = None,
                    bid_k: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Inference attention of Sq new query rows over an Sk-row K/V cache (pi0_arch.py:172-192 with
    update_cache=False).  q_packed [B, Sq, (H+2KVH)*hd] (its q columns are read); kv [B, Sk, 2*KVH*hd] with K in the
    first KVH*hd columns and V in the rest.  Returns out [B, Sq, H*hd]."""
    _cuda(q_packed, kv)
    dt, dev, es = _dt(q_packed), q_packed.device, q_packed.element_size()
    G, W, Wkv = H // KVH, (H + 2 * KVH) * hd, 2 * KVH * hd
    assert q_packed.is_contiguous() and kv.is_contiguous() and kv.shape[-1] == Wkv and q_packed.shape[-1] == W
    lds = (Sk + 3) // 4 * 4
    ldp = (Sk + 7) // 8 * 8 if q_packed.dtype == torch.bfloat16 else lds
    scores = torch.empty((B, H, Sq, lds), device=dev, dtype=torch.float32)
    probs = torch.empty((B, H, Sq, ldp), device=dev, dtype=q_packed.dtype)
    if out is None:
        out = torch.empty((B, Sq, H * hd), device=dev, dtype=q_packed.dtype)
    k_ptr = kv.data_ptr()
    v_ptr = k_ptr + KVH * hd * es
    gemm_raw(a=q_packed.data_ptr(), b=k_ptr, d=scores.data_ptr(), ab_dtype=dt, d_dtype=F32, a_mn_major=0, b_mn_major=0,
             m=Sq, n=Sk, k=hd, a_ld=W, a_s2=hd, a_s3=Sq * W, a_z2=H, b_ld=Wkv, b_s2=hd, b_s3=Sk * Wkv, b_z2=KVH,
             d_ld=lds, d_s2=Sq * lds, d_s3=H * Sq * lds, z_lo=H, z_hi=B,
             a_div=1, a_mul=1, a_seg=0, b_div=G, b_mul=1, b_seg=0, k_segs=1, alpha=hd ** -0.5)
    softmax_fwd(scores, probs, Sq, Sk, heads=H, keymask=keymask, bid_q=bid_q, bid_k=bid_k)
    gemm_raw(a=probs.data_ptr(), b=v_ptr, d=out.data_ptr(), ab_dtype=dt, d_dtype=dt, a_mn_major=0, b_mn_major=1,
             m=Sq, n=hd, k=Sk, a_ld=ldp, a_s2=Sq * ldp, a_s3=H * Sq * ldp, a_z2=H,
             b_ld=Wkv, b_s2=hd, b_s3=Sk * Wkv, b_z2=KVH, d_ld=H * hd, d_s2=hd, d_s3=Sq * H * hd, z_lo=H, z_hi=B,
             a_div=1, a_mul=1, a_seg=0, b_div=G, b_mul=1, b_seg=0, k_segs=1, alpha=1.0)
    return out


def attention_bwd(dout: torch.Tensor, qkv: torch.Tensor, probs: torch.Tensor, sh: AttnShape, *, causal: bool = False,
                  dqkv: Optional[torch.Tensor] = None, scratch: Optional[torch.Tensor] = None,
                  dprobs: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, keymask=None,
                  bid_q=None, bid_k=None) -> torch.Tensor:
    """Gradient of attention_fwd w.r.t. the packed qkv buffer (before the inverse RoPE).  `probs` is what
    attention_fwd returned second: the log-sum-exp (flash path: `out` and the mask arguments are then required, P is
    recomputed) or the probabilities."""
    _cuda(dout, qkv, probs)
    if probs.dim() == 3:
        assert out is not None, "attention_bwd: the flash path needs the forward output"
        return flash_attention_bwd(dout, qkv, out.view(sh.B, sh.S, -1), probs, sh, keymask=keymask, bid_q=bid_q,
                                   bid_k=bid_k, causal=causal, dqkv=dqkv)
    B, S, H, KVH, hd, G, W = sh.B, sh.S, sh.H, sh.KVH, sh.hd, sh.G, sh.W
    dt = _dt(qkv)
    dev = qkv.device
    assert dout.is_contiguous() and dout.shape == (B, S, H * hd)
    if dqkv is None:
        dqkv = torch.empty_like(qkv)
    if dprobs is None:
        dprobs = torch.empty_like(probs)
    q, k, v = _qkv_ptrs(qkv, sh)
    dq, dk, dv = _qkv_ptrs(dqkv, sh)
    ldp, lds = sh.ld_p, sh.ld_s
    # dV = sum_g P_g^T dO_g
    gemm_raw(a=probs.data_ptr(), b=dout.data_ptr(), d=dv, ab_dtype=dt, d_dtype=dt, a_mn_major=1, b_mn_major=1,
             m=S, n=hd, k=S, a_ld=ldp, a_s2=S * ldp, a_s3=H * S * ldp, a_z2=H, b_ld=H * hd, b_s2=hd,
             b_s3=S * H * hd, b_z2=H, d_ld=W, d_s2=hd, d_s3=S * W, z_lo=KVH, z_hi=B, a_div=1, a_mul=G, a_seg=1,
             b_div=1, b_mul=G, b_seg=1, k_segs=G, alpha=1.0)
    if _fused_scores_ok(qkv, sh):
        # dS = scale * P * (dO V^T - rowsum(P * dO V^T)): dP never leaves TMEM
        attn_scores(dout.data_ptr(), H * hd, hd, S * H * hd, v, W, hd, S * W, sh, dprobs, p_in=probs, causal=causal)
    else:
        if scratch is None:
            scratch = torch.empty((B, H, S, sh.ld_s), device=dev, dtype=torch.float32)
        # dP = dO V^T   (fp32)
        gemm_raw(a=dout.data_ptr(), b=v, d=scratch.data_ptr(), ab_dtype=dt, d_dtype=F32, a_mn_major=0, b_mn_major=0,
                 m=S, n=S, k=hd, a_ld=H * hd, a_s2=hd, a_s3=S * H * hd, a_z2=H, b_ld=W, b_s2=hd, b_s3=S * W,
                 b_z2=KVH, d_ld=lds, d_s2=S * lds, d_s3=H * S * lds, z_lo=H, z_hi=B, a_div=1, a_mul=1, a_seg=0,
                 b_div=G, b_mul=1, b_seg=0, k_segs=1, alpha=1.0)
        # dS = scale * P * (dP - rowsum(P dP))
        softmax_bwd(probs, scratch, dprobs, B * H * S, S, ldp, lds, ldp, sh.scale)
    # dQ = dS K
    gemm_raw(a=dprobs.data_ptr(), b=k, d=dq, ab_dtype=dt, d_dtype=dt, a_mn_major=0, b_mn_major=1, m=S, n=hd, k=S,
             a_ld=ldp, a_s2=S * ldp, a_s3=H * S * ldp, a_z2=H, b_ld=W, b_s2=hd, b_s3=S * W, b_z2=KVH,
             d_ld=W, d_s2=hd, d_s3=S * W, z_lo=H, z_hi=B, a_div=1, a_mul=1, a_seg=0, b_div=G, b_mul=1, b_seg=0,
             k_segs=1, alpha=1.0)
    # dK = sum_g dS_g^T Q_g
    gemm_raw(a=dprobs.data_ptr(), b=q, d=dk, ab_dtype=dt, d_dtype=dt, a_mn_major=1, b_mn_major=1, m=S, n=hd, k=S,
             a_ld=ldp, a_s2=S * ldp, a_s3=H * S * ldp, a_z2=H, b_ld=W, b_s2=hd, b_s3=S * W, b_z2=H,
             d_ld=W, d_s2=hd, d_s3=S * W, z_lo=KVH, z_hi=B, a_div=1, a_mul=G, a_seg=1, b_div=1, b_mul=G, b_seg=1,
             k_segs=G, alpha=1.0)
    return dqkv


# ------------------------------------------------------------------- elementwise
def rmsnorm_fwd(x, w, eps: float, unit_offset: bool = False, out=None, want_rstd: bool = True):
    _cuda(x, w)
    D = x.shape[-1]
    M = x.numel() // D
    assert x.is_contiguous() and w.is_contiguous() and w.dtype == x.dtype
    y = torch.empty_like(x) if out is None else out
    rstd = torch.empty(M, device=x.device, dtype=torch.float32) if want_rstd else None
    _lib.check(_lib.load().b200_rmsnorm_fwd(x.data_ptr(), w.data_ptr(), y.data_ptr(), _p(rstd), M, D, float(eps),
                                            int(unit_offset), _dt(x), _stream()), "rmsnorm_fwd")
    return y, rstd


def rmsnorm_bwd(dy, x, w, rstd, unit_offset: bool = False, dx=None, dw=None, accumulate_dx: bool = False):
    _cuda(dy, x, w, rstd)
    D = x.shape[-1]
    M = x.numel() // D
    assert dy.is_contiguous() and x.is_contiguous()
    if dx is None:
        assert not accumulate_dx
        dx = torch.empty_like(x)
    ws = None
    if dw is not None:
        assert dw.dtype == torch.float32 and dw.numel() == D
        ws = torch.empty((int(_lib.load().b200_norm_bwd_workspace_rows(M, D)), D), device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().b200_rmsnorm_bwd(dy.data_ptr(), x.data_ptr(), w.data_ptr(), rstd.data_ptr(), dx.data_ptr(),
                                            _p(dw), _p(ws), M, D, int(unit_offset), int(accumulate_dx), _dt(x),
                                            _stream()), "rmsnorm_bwd")
    return dx


def layernorm_fwd(x, w, b, eps: float, out=None):
    _cuda(x, w, b)
    D = x.shape[-1]
    M = x.numel() // D
    assert x.is_contiguous()
    y = torch.empty_like(x) if out is None else out
    mean = torch.empty(M, device=x.device, dtype=torch.float32)
    rstd = torch.empty(M, device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().b200_layernorm_fwd(x.data_ptr(), _p(w), _p(b), y.data_ptr(), mean.data_ptr(),
                                              rstd.data_ptr(), M, D, float(eps), _dt(x), _stream()), "layernorm_fwd")
    return y, mean, rstd


def layernorm_bwd(dy, x, w, mean, rstd, dx=None, dw=None, db=None, accumulate_dx: bool = False):
    _cuda(dy, x, w, mean, rstd)
    D = x.shape[-1]
    M = x.numel() // D
    assert dy.is_contiguous() and x.is_contiguous()
    if dx is None:
        assert not accumulate_dx
        dx = torch.empty_like(x)
    ws = None
    if dw is not None and db is not None and D <= 8192:     # wider rows take the streaming kernel (no workspace)
        ws = torch.empty((int(_lib.load().b200_norm_bwd_workspace_rows(M, D)), 2 * D), device=x.device,
                         dtype=torch.float32)
    _lib.check(_lib.load().b200_layernorm_bwd(dy.data_ptr(), x.data_ptr(), _p(w), mean.data_ptr(), rstd.data_ptr(),
                                              dx.data_ptr(), _p(dw), _p(db), _p(ws), M, D, int(accumulate_dx), _dt(x),
                                              _stream()), "layernorm_bwd")
    return dx


def rope_(qkv: torch.Tensor, pos: torch.Tensor, cos_t: torch.Tensor, sin_t: torch.Tensor, n_rot_heads: int, hd: int,
          inverse: bool = False) -> torch.Tensor:
    """In-place rotate_half RoPE on the first n_rot_heads heads of each row of qkv[..., W]."""
    _cuda(qkv, pos, cos_t, sin_t)
    W = qkv.shape[-1]
    M = qkv.numel() // W
    assert qkv.is_contiguous() and pos.dtype == torch.int32 and pos.numel() == M
    assert cos_t.dtype == torch.float32 and cos_t.shape[-1] == hd // 2 and cos_t.is_contiguous()
    _lib.check(_lib.load().b200_rope(qkv.data_ptr(), pos.data_ptr(), cos_t.data_ptr(), sin_t.data_ptr(), M,
                                     n_rot_heads, hd, W, int(inverse), _dt(qkv), _stream()), "rope")
    return qkv


def act_fwd(x, act, out=None):
    _cuda(x)
    y = torch.empty_like(x) if out is None else out
    _lib.check(_lib.load().b200_act_fwd(x.data_ptr(), y.data_ptr(), x.numel(), act_code(act), _dt(x), _stream()),
               "act_fwd")
    return y


def act_bwd(dy, x, act, out=None):
    _cuda(dy, x)
    dx = torch.empty_like(x) if out is None else out
    _lib.check(_lib.load().b200_act_bwd(dy.data_ptr(), x.data_ptr(), dx.data_ptr(), x.numel(), act_code(act), _dt(x),
                                        _stream()), "act_bwd")
    return dx


def glu_fwd(g, u, act, out=None):
    _cuda(g, u)
    h = torch.empty_like(g) if out is None else out
    _lib.check(_lib.load().b200_glu_fwd(g.data_ptr(), u.data_ptr(), h.data_ptr(), g.numel(), act_code(act), _dt(g),
                                        _stream()), "glu_fwd")
    return h


def glu_bwd(dh, g, u, act, dg=None, du=None, h_out=None):
    _cuda(dh, g, u)
    dg = torch.empty_like(g) if dg is None else dg
    du = torch.empty_like(u) if du is None else du
    _lib.check(_lib.load().b200_glu_bwd(dh.data_ptr(), g.data_ptr(), u.data_ptr(), dg.data_ptr(), du.data_ptr(),
                                        _p(h_out), g.numel(), act_code(act), _dt(g), _stream()), "glu_bwd")
    return dg, du


def softmax_fwd(scores, probs, Sq: int, Sk: int, heads: int = 1, keymask=None, bid_q=None, bid_k=None,
                causal: bool = False):
    _cuda(scores, probs, keymask, bid_q, bid_k)
    assert scores.dtype == torch.float32
    Z = scores.numel() // (Sq * scores.shape[-1])
    if keymask is not None:
        assert keymask.dtype == torch.uint8 and keymask.is_contiguous()
    for t in (bid_q, bid_k):
        if t is not None:
            assert t.dtype == torch.int32 and t.is_contiguous()
    _lib.check(_lib.load().b200_softmax_fwd(scores.data_ptr(), probs.data_ptr(), Z, Sq, Sk, scores.shape[-1],
                                            probs.shape[-1], heads, _p(keymask), _p(bid_q), _p(bid_k), int(causal),
                                            _dt(probs), _stream()), "softmax_fwd")
    return probs


def softmax_bwd(probs, dprobs_f32, ds, rows: int, Sk: int, p_ld: int, dp_ld: int, ds_ld: int, scale: float):
    _cuda(probs, dprobs_f32, ds)
    _lib.check(_lib.load().b200_softmax_bwd(probs.data_ptr(), dprobs_f32.data_ptr(), ds.data_ptr(), rows, Sk, p_ld,
                                            dp_ld, ds_ld, float(scale), _dt(probs), _stream()), "softmax_bwd")
    return ds


def colsum_(x2d, out_f32):
    _cuda(x2d, out_f32)
    assert x2d.is_contiguous() and out_f32.dtype == torch.float32
    M, N = x2d.shape
    _lib.check(_lib.load().b200_colsum(x2d.data_ptr(), out_f32.data_ptr(), M, N, _dt(x2d), _stream()), "colsum")
    return out_f32


def sumsq_(x, out_f32):
    _cuda(x, out_f32)
    _lib.check(_lib.load().b200_sumsq(x.data_ptr(), x.numel(), out_f32.data_ptr(), _dt(x), _stream()), "sumsq")
    return out_f32


def clip_coef(sumsq, max_norm: float, clip, norm_out=None):
    _cuda(sumsq, clip)
    _lib.check(_lib.load().b200_clip_coef(sumsq.data_ptr(), float(max_norm), clip.data_ptr(), _p(norm_out),
                                          _stream()), "clip_coef")
    return clip


def mse_fwd(a, b, out_f32):
    _cuda(a, b, out_f32)
    assert a.is_contiguous() and b.is_contiguous() and a.dtype == b.dtype
    _lib.check(_lib.load().b200_mse_fwd(a.data_ptr(), b.data_ptr(), a.numel(), out_f32.data_ptr(), _dt(a), _stream()),
               "mse_fwd")
    return out_f32


def mse_bwd(a, b, gscale=None, out=None):
    _cuda(a, b, gscale)
    da = torch.empty_like(a) if out is None else out
    _lib.check(_lib.load().b200_mse_bwd(a.data_ptr(), b.data_ptr(), a.numel(), _p(gscale), da.data_ptr(), _dt(a),
                                        _stream()), "mse_bwd")
    return da


def adamw_(p32, g, m, v, shadow_bf16, lr, beta1, beta2, eps, wd, step: int, clip=None):
    _cuda(p32, g, m, v, shadow_bf16, clip)
    assert p32.dtype == m.dtype == v.dtype == torch.float32
    _lib.check(_lib.load().b200_adamw(p32.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), _p(shadow_bf16),
                                      p32.numel(), float(lr), float(beta1), float(beta2), float(eps), float(wd),
                                      int(step), _p(clip), _dt(g), _stream()), "adamw")


def adamw_hyper_(out8: torch.Tensor, lr, beta1, beta2, eps, wd, step: int) -> None:
    """Fill a HOST float32[8] block with the seven scalars b200_adamw derives from its doubles (step <= 0: identity)."""
    assert out8.device.type == "cpu" and out8.dtype == torch.float32 and out8.numel() >= 8 and out8.is_contiguous()
    _lib.check(_lib.load().b200_adamw_hyper(float(lr), float(beta1), float(beta2), float(eps), float(wd), int(step),
                                            out8.data_ptr()), "adamw_hyper")


def adamw_dev_(p32, g, m, v, shadow_bf16, hyper8: torch.Tensor, clip=None):
    """adamw_ with its scalars read from the device block `hyper8` (graph-replayable optimizer, graph.py)."""
    _cuda(p32, g, m, v, shadow_bf16, clip, hyper8)
    assert p32.dtype == m.dtype == v.dtype == hyper8.dtype == torch.float32 and hyper8.numel() >= 8
    _lib.check(_lib.load().b200_adamw_dev(p32.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), _p(shadow_bf16),
                                          p32.numel(), hyper8.data_ptr(), _p(clip), _dt(g), _stream()), "adamw_dev")


def reduce_scatter_p2p_(own: torch.Tensor, peers: list, scale: float, ctas: int = 16) -> None:
    """own = scale * (own + sum(peers)) in place, bf16 with fp32 accumulation in list order; `peers` are views of the
    other ranks' buffers (symmetric memory, peer-mapped).  One small kernel: see csrc/exchange.cu."""
    _cuda(own, *peers)
    assert own.dtype == torch.bfloat16 and own.is_contiguous() and all(t.dtype == own.dtype and t.numel() == own.numel()
                                                                        and t.is_contiguous() for t in peers)
    arr = (C.c_void_p * len(peers))(*[t.data_ptr() for t in peers])
    _lib.check(_lib.load().b200_reduce_scatter_p2p(own.data_ptr(), C.cast(arr, C.c_void_p), len(peers), own.numel(),
                                                   float(scale), int(ctas), _stream()), "reduce_scatter_p2p")


def cast_(src, dst):
    _cuda(src, dst)
    assert src.numel() == dst.numel() and src.is_contiguous() and dst.is_contiguous()
    _lib.check(_lib.load().b200_cast(src.data_ptr(), dst.data_ptr(), src.numel(), _dt(src), _dt(dst), _stream()),
               "cast")
    return dst


def add(a, b, out=None):
    _cuda(a, b)
    y = torch.empty_like(a) if out is None else out
    _lib.check(_lib.load().b200_add(a.data_ptr(), b.data_ptr(), y.data_ptr(), a.numel(), _dt(a), _stream()), "add")
    return y


# ------------------------------------------------------------------ index-driven
def splice_lengths(input_ids, attention_mask_u8, n_img_tokens: int, max_len: int):
    _cuda(input_ids, attention_mask_u8)
    B, L = input_ids.shape
    lengths = torch.empty(B, device=input_ids.device, dtype=torch.int32)
    _lib.check(_lib.load().b200_splice_lengths(input_ids.data_ptr(), _p(attention_mask_u8), B, L, n_img_tokens,
                                               int(max_len or 0), lengths.data_ptr(), _stream()), "splice_lengths")
    return lengths


def splice_plan(input_ids, attention_mask_u8, labels, n_img_tokens: int, max_len: int, S: int, left_pad: bool):
    _cuda(input_ids, attention_mask_u8, labels)
    B, L = input_ids.shape
    dev = input_ids.device
    src = torch.empty((B, S), device=dev, dtype=torch.int32)
    new_labels = torch.empty((B, S), device=dev, dtype=torch.int64)
    new_mask = torch.empty((B, S), device=dev, dtype=torch.uint8)
    pos = torch.empty((B, S), device=dev, dtype=torch.int32)
    _lib.check(_lib.load().b200_splice_plan(input_ids.data_ptr(), _p(attention_mask_u8), _p(labels), B, L,
                                            n_img_tokens, int(max_len or 0), S, int(left_pad), src.data_ptr(),
                                            new_labels.data_ptr(), new_mask.data_ptr(), pos.data_ptr(), _stream()),
               "splice_plan")
    return src, new_labels, new_mask, pos


def splice_gather(src, table, feats, out=None):
    _cuda(src, table, feats)
    D = table.shape[-1]
    rows = src.numel()
    if out is None:
        out = torch.empty(tuple(src.shape) + (D,), device=table.device, dtype=table.dtype)
    _lib.check(_lib.load().b200_splice_gather(src.data_ptr(), table.data_ptr(), _p(feats), out.data_ptr(), rows, D,
                                              _dt(table), _stream()), "splice_gather")
    return out


def splice_scatter(src, dout, d_table, d_feats):
    _cuda(src, dout, d_table, d_feats)
    D = dout.shape[-1]
    assert dout.is_contiguous()
    _lib.check(_lib.load().b200_splice_scatter(src.data_ptr(), dout.data_ptr(), _p(d_table), _p(d_feats), src.numel(),
                                               D, _dt(dout), _stream()), "splice_scatter")


def gather_rows(x2d, idx, out=None):
    _cuda(x2d, idx)
    D = x2d.shape[-1]
    assert idx.dtype == torch.int32 and x2d.is_contiguous()
    if out is None:
        out = torch.empty((idx.numel(), D), device=x2d.device, dtype=x2d.dtype)
    _lib.check(_lib.load().b200_gather_rows(x2d.data_ptr(), idx.data_ptr(), out.data_ptr(), idx.numel(), D, _dt(x2d),
                                            _stream()), "gather_rows")
    return out


def scatter_rows_add_(dout, idx, dx2d):
    _cuda(dout, idx, dx2d)
    D = dx2d.shape[-1]
    _lib.check(_lib.load().b200_scatter_rows_add(dout.data_ptr(), idx.data_ptr(), dx2d.data_ptr(), idx.numel(), D,
                                                 _dt(dx2d), _stream()), "scatter_rows_add")
    return dx2d


def last_valid_index(mask_u8):
    _cuda(mask_u8)
    B, S = mask_u8.shape
    idx = torch.empty(B, device=mask_u8.device, dtype=torch.int32)
    _lib.check(_lib.load().b200_last_valid_index(mask_u8.data_ptr(), B, S, idx.data_ptr(), _stream()),
               "last_valid_index")
    return idx


def q_sample(x, noise, t_i32, sqrt_ac, sqrt_1mac, out=None):
    _cuda(x, noise, t_i32, sqrt_ac, sqrt_1mac)
    B = x.shape[0]
    xt = torch.empty_like(x) if out is None else out
    _lib.check(_lib.load().b200_q_sample(x.data_ptr(), noise.data_ptr(), t_i32.data_ptr(), sqrt_ac.data_ptr(),
                                         sqrt_1mac.data_ptr(), xt.data_ptr(), B, x.numel() // B, _dt(x), _stream()),
               "q_sample")
    return xt


def timestep_embedding(t_f32, dim: int, dtype: torch.dtype, max_period: float = 10000.0):
    _cuda(t_f32)
    B = t_f32.numel()
    out = torch.empty((B, dim), device=t_f32.device, dtype=dtype)
    _lib.check(_lib.load().b200_timestep_embedding(t_f32.data_ptr(), out.data_ptr(), B, dim, float(max_period),
                                                   _dt(out), _stream()), "timestep_embedding")
    return out


def discretize_actions(actions_f32, n_bins: int):
    _cuda(actions_f32)
    assert actions_f32.dtype == torch.float32 and actions_f32.is_contiguous()
    bins = torch.empty(actions_f32.shape, device=actions_f32.device, dtype=torch.int64)
    _lib.check(_lib.load().b200_discretize_actions(actions_f32.data_ptr(), actions_f32.numel(), n_bins,
                                                   bins.data_ptr(), _stream()), "discretize_actions")
    return bins


def bins_to_continuous(bins_i64, n_bins: int):
    _cuda(bins_i64)
    out = torch.empty(bins_i64.shape, device=bins_i64.device, dtype=torch.float32)
    _lib.check(_lib.load().b200_bins_to_continuous(bins_i64.data_ptr(), bins_i64.numel(), n_bins, out.data_ptr(),
                                                   _stream()), "bins_to_continuous")
    return out


def argmax_last(logits2d, n_last: int):
    _cuda(logits2d)
    rows, V = logits2d.shape
    assert logits2d.is_contiguous()
    idx = torch.empty(rows, device=logits2d.device, dtype=torch.int64)
    _lib.check(_lib.load().b200_argmax_last(logits2d.data_ptr(), rows, V, n_last, idx.data_ptr(), _dt(logits2d),
                                            _stream()), "argmax_last")
    return idx


def sample_last(logits2d, n_last: int, temperature: float, u: torch.Tensor):
    """One draw per row from softmax(logits[:, -n_last:] / temperature): inverse CDF at the uniforms u [rows] (fp32)."""
    _cuda(logits2d, u)
    rows, V = logits2d.shape
    assert logits2d.is_contiguous() and u.dtype == torch.float32 and u.numel() == rows and u.is_contiguous()
    idx = torch.empty(rows, device=logits2d.device, dtype=torch.int64)
    _lib.check(_lib.load().b200_sample_last(logits2d.data_ptr(), rows, V, n_last, float(temperature), u.data_ptr(),
                                            idx.data_ptr(), _dt(logits2d), _stream()), "sample_last")
    return idx


def cross_entropy_fwd(logits2d, labels):
    _cuda(logits2d, labels)
    rows, V = logits2d.shape
    assert logits2d.is_contiguous() and labels.dtype == torch.int64 and labels.numel() == rows
    dev = logits2d.device
    lse = torch.empty(rows, device=dev, dtype=torch.float32)
    loss_sum = torch.zeros((), device=dev, dtype=torch.float32)
    n_valid = torch.zeros((), device=dev, dtype=torch.int32)
    _lib.check(_lib.load().b200_cross_entropy_fwd(logits2d.data_ptr(), labels.data_ptr(), rows, V, lse.data_ptr(),
                                                  loss_sum.data_ptr(), n_valid.data_ptr(), _dt(logits2d), _stream()),
               "cross_entropy_fwd")
    return loss_sum, n_valid, lse


def cross_entropy_bwd(logits2d, labels, lse, n_valid, gscale=None, out=None):
    _cuda(logits2d, labels, lse, n_valid, gscale)
    rows, V = logits2d.shape
    d = torch.empty_like(logits2d) if out is None else out
    _lib.check(_lib.load().b200_cross_entropy_bwd(logits2d.data_ptr(), labels.data_ptr(), lse.data_ptr(),
                                                  n_valid.data_ptr(), _p(gscale), d.data_ptr(), rows, V,
                                                  _dt(logits2d), _stream()), "cross_entropy_bwd")
    return d


# ------------------------------------------------------------------ ViT front end
def im2col_patches(images, patch: int, k_pad: int, out_dtype=torch.bfloat16):
    _cuda(images)
    B, Cc, H, W = images.shape
    assert images.is_contiguous()
    P = (H // patch) * (W // patch)
    out = torch.empty((B * P, k_pad), device=images.device, dtype=out_dtype)
    _lib.check(_lib.load().b200_im2col_patches(images.data_ptr(), out.data_ptr(), B, Cc, H, W, patch, k_pad,
                                               _dt(images), _dt(out), _stream()), "im2col_patches")
    return out


def vit_embed_fwd(patches, cls, pos, B: int, P: int):
    _cuda(patches, cls, pos)
    D = patches.shape[-1]
    out = torch.empty((B, P + 1, D), device=patches.device, dtype=patches.dtype)
    _lib.check(_lib.load().b200_vit_embed_fwd(patches.data_ptr(), cls.data_ptr(), pos.data_ptr(), out.data_ptr(), B, P,
                                              D, _dt(patches), _stream()), "vit_embed_fwd")
    return out


def vit_embed_bwd(dout, d_patches, d_cls_f32, d_pos_f32, B: int, P: int):
    _cuda(dout, d_patches, d_cls_f32, d_pos_f32)
    D = dout.shape[-1]
    assert dout.is_contiguous()
    _lib.check(_lib.load().b200_vit_embed_bwd(dout.data_ptr(), _p(d_patches), _p(d_cls_f32), _p(d_pos_f32), B, P, D,
                                              _dt(dout), _stream()), "vit_embed_bwd")


def cast_add_(src_f32, dst, accumulate: bool):
    _cuda(src_f32, dst)
    assert src_f32.dtype == torch.float32 and src_f32.numel() == dst.numel()
    _lib.check(_lib.load().b200_cast_add(src_f32.data_ptr(), dst.data_ptr(), dst.numel(), _dt(dst), int(accumulate),
                                         _stream()), "cast_add")
    return dst


def copy2d_(src, dst, rows: int, cols: int, accumulate: bool = False):
    _cuda(src, dst)
    _lib.check(_lib.load().b200_copy2d(src.data_ptr(), dst.data_ptr(), rows, cols, src.stride(0), dst.stride(0),
                                       _dt(src), _dt(dst), int(accumulate), _stream()), "copy2d")
    return dst


def copy3d_(src, dst, B: int, rows: int, cols: int, src_bs: int, src_ld: int, dst_bs: int, dst_ld: int,
            alpha: float = 1.0, accumulate: bool = False, src_off: int = 0, dst_off: int = 0):
    """dst[b, r, :cols] (+)= alpha * src[b, r, :cols]; offsets / strides in elements over the flat storage."""
    _cuda(src, dst)
    assert src.dtype == dst.dtype
    es = src.element_size()
    _lib.check(_lib.load().b200_copy3d(src.data_ptr() + src_off * es, dst.data_ptr() + dst_off * es, B, rows, cols,
                                       src_bs, src_ld, dst_bs, dst_ld, float(alpha), int(accumulate), _dt(src),
                                       _stream()), "copy3d")
    return dst


def add_pos_fwd(x, pos, B: int, P: int):
    _cuda(x, pos)
    out = torch.empty_like(x)
    _lib.check(_lib.load().b200_add_pos_fwd(x.data_ptr(), pos.data_ptr(), out.data_ptr(), B, P, x.shape[-1], _dt(x),
                                            _stream()), "add_pos_fwd")
    return out


def add_pos_bwd(dout, d_pos_f32, B: int, P: int):
    _cuda(dout, d_pos_f32)
    _lib.check(_lib.load().b200_add_pos_bwd(dout.data_ptr(), d_pos_f32.data_ptr(), B, P, dout.shape[-1], _dt(dout),
                                            _stream()), "add_pos_bwd")


# ------------------------------------------------------------------- MemVLA memory path
def dropout(x2d: torch.Tensor, p: float, seed: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Counter-based dropout over the logical [rows, cols] region of a (possibly row-padded) 2-D tensor; the same
    (p, seed, shape) on the gradient is the backward pass."""
    _cuda(x2d)
    assert x2d.dim() == 2 and x2d.stride(1) == 1
    if out is None:
        out = torch.empty_like(x2d)
    _lib.check(_lib.load().b200_dropout(x2d.data_ptr(), out.data_ptr(), x2d.shape[0], x2d.shape[1], x2d.stride(0),
                                        out.stride(0), float(p), int(seed) & (2 ** 64 - 1), _dt(x2d), _stream()),
               "dropout")
    return out


def se_reduce(x3d: torch.Tensor, y3d: Optional[torch.Tensor], scale: float) -> torch.Tensor:
    """out_f32[b, c] = scale * sum_p x[b,p,c] * (y[b,p,c] if y is given else 1)."""
    _cuda(x3d, y3d)
    B, P, Cc = x3d.shape
    out = torch.zeros((B, Cc), device=x3d.device, dtype=torch.float32)
    _lib.check(_lib.load().b200_se_reduce(x3d.data_ptr(), _p(y3d), out.data_ptr(), B, P, Cc, float(scale), _dt(x3d),
                                          _stream()), "se_reduce")
    return out


def se_scale(x3d: torch.Tensor, w2d: torch.Tensor, add2d: Optional[torch.Tensor] = None, add_scale: float = 0.0,
             out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[b,p,c] = x[b,p,c] * w[b,c] (+ add[b,c] * add_scale)."""
    _cuda(x3d, w2d, add2d)
    B, P, Cc = x3d.shape
    assert x3d.is_contiguous() and w2d.is_contiguous() and w2d.dtype == x3d.dtype
    if out is None:
        out = torch.empty_like(x3d)
    _lib.check(_lib.load().b200_se_scale(x3d.data_ptr(), w2d.data_ptr(), _p(add2d), out.data_ptr(), B, P, Cc,
                                         float(add_scale), _dt(x3d), _stream()), "se_scale")
    return out


def gate_fuse_fwd(z, x1, x2):
    _cuda(z, x1, x2)
    out = torch.empty_like(x1)
    _lib.check(_lib.load().b200_gate_fuse_fwd(z.data_ptr(), x1.data_ptr(), x2.data_ptr(), out.data_ptr(), x1.numel(),
                                              _dt(x1), _stream()), "gate_fuse_fwd")
    return out


def gate_fuse_bwd(dout, z, x1, x2):
    _cuda(dout, z, x1, x2)
    dz, dx1, dx2 = torch.empty_like(z), torch.empty_like(x1), torch.empty_like(x2)
    _lib.check(_lib.load().b200_gate_fuse_bwd(dout.data_ptr(), z.data_ptr(), x1.data_ptr(), x2.data_ptr(), dz.data_ptr(),
                                              dx1.data_ptr(), dx2.data_ptr(), x1.numel(), _dt(x1), _stream()),
               "gate_fuse_bwd")
    return dz, dx1, dx2


def cross_attention_fwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, B: int, Sq: int, Sk: int, H: int, *,
                        dropout_p: float = 0.0, seed: int = 0):
    """softmax(Q K^T / sqrt(hd)) V with separate query / key lengths and no mask: q [B*Sq, H*hd], k, v [B*Sk, H*hd]
    (row-major, heads side by side).  Returns (out [B*Sq, H*hd], probs [B,H,Sq,ld], probs_dropped or None)."""
    _cuda(q, k, v)
    dt, dev = _dt(q), q.device
    D = q.shape[1]
    hd = D // H
    if (hd * q.element_size()) % 16 != 0:
        raise RuntimeError(f"cross_attention: head_dim {hd} x {q.element_size()} B must be a multiple of 16 B (TMA)")
    lds = (Sk + 3) // 4 * 4
    ldp = (Sk + 7) // 8 * 8 if q.dtype == torch.bfloat16 else lds
    scores = torch.empty((B, H, Sq, lds), device=dev, dtype=torch.float32)
    probs = torch.empty((B, H, Sq, ldp), device=dev, dtype=q.dtype)
    gemm_raw(a=q.data_ptr(), b=k.data_ptr(), d=scores.data_ptr(), ab_dtype=dt, d_dtype=F32, a_mn_major=0, b_mn_major=0,
             m=Sq, n=Sk, k=hd, a_ld=D, a_s2=hd, a_s3=Sq * D, a_z2=H, b_ld=D, b_s2=hd, b_s3=Sk * D, b_z2=H,
             d_ld=lds, d_s2=Sq * lds, d_s3=H * Sq * lds, z_lo=H, z_hi=B,
             a_div=1, a_mul=1, a_seg=0, b_div=1, b_mul=1, b_seg=0, k_segs=1, alpha=hd ** -0.5)
    softmax_fwd(scores, probs, Sq, Sk, heads=H)
    pd = None
    if dropout_p > 0.0:
        pd = torch.zeros_like(probs)       # padding columns stay zero (they are read as K by the PV GEMM)
        dropout(probs.view(-1, ldp)[:, :Sk], dropout_p, seed, out=pd.view(-1, ldp)[:, :Sk])
    pa = probs if pd is None else pd
    out = torch.empty((B * Sq, D), device=dev, dtype=q.dtype)
    gemm_raw(a=pa.data_ptr(), b=v.data_ptr(), d=out.data_ptr(), ab_dtype=dt, d_dtype=dt, a_mn_major=0, b_mn_major=1,
             m=Sq, n=hd, k=Sk, a_ld=ldp, a_s2=Sq * ldp, a_s3=H * Sq * ldp, a_z2=H,
             b_ld=D, b_s2=hd, b_s3=Sk * D, b_z2=H, d_ld=D, d_s2=hd, d_s3=Sq * D, z_lo=H, z_hi=B,
             a_div=1, a_mul=1, a_seg=0, b_div=1, b_mul=1, b_seg=0, k_segs=1, alpha=1.0)
    return out, probs, pd


def cross_attention_bwd(dout: torch.Tensor, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, probs: torch.Tensor,
                        probs_dropped: Optional[torch.Tensor], B: int, Sq: int, Sk: int, H: int, *,
                        dropout_p: float = 0.0, seed: int = 0, need_dq: bool = True, need_dkv: bool = True):
    """Gradients (dq, dk, dv) of cross_attention_fwd."""
    _cuda(dout, q, k, v, probs)
    dt, dev = _dt(q), q.device
    D = q.shape[1]
    hd = D // H
    ldp = probs.shape[-1]
    lds = (Sk + 3) // 4 * 4
    scale = hd ** -0.5
    pa = probs if probs_dropped is None else probs_dropped
    dq = dk = dv = None
    if need_dkv:
        dv = torch.empty_like(v)
        gemm_raw(a=pa.data_ptr(), b=dout.data_ptr(), d=dv.data_ptr(), ab_dtype=dt, d_dtype=dt, a_mn_major=1,
                 b_mn_major=1, m=Sk, n=hd, k=Sq, a_ld=ldp, a_s2=Sq * ldp, a_s3=H * Sq * ldp, a_z2=H, b_ld=D, b_s2=hd,
                 b_s3=Sq * D, b_z2=H, d_ld=D, d_s2=hd, d_s3=Sk * D, z_lo=H, z_hi=B, a_div=1, a_mul=1, a_seg=0,
                 b_div=1, b_mul=1, b_seg=0, k_segs=1, alpha=1.0)
    dp = torch.empty((B, H, Sq, lds), device=dev, dtype=torch.float32)
    gemm_raw(a=dout.data_ptr(), b=v.data_ptr(), d=dp.data_ptr(), ab_dtype=dt, d_dtype=F32, a_mn_major=0, b_mn_major=0,
             m=Sq, n=Sk, k=hd, a_ld=D, a_s2=hd, a_s3=Sq * D, a_z2=H, b_ld=D, b_s2=hd, b_s3=Sk * D, b_z2=H,
             d_ld=lds, d_s2=Sq * lds, d_s3=H * Sq * lds, z_lo=H, z_hi=B, a_div=1, a_mul=1, a_seg=0,
             b_div=1, b_mul=1, b_seg=0, k_segs=1, alpha=1.0)
    if dropout_p > 0.0:
        dropout(dp.view(-1, lds)[:, :Sk], dropout_p, seed, out=dp.view(-1, lds)[:, :Sk])
    ds = torch.zeros_like(probs) if ldp != Sk else torch.empty_like(probs)
    softmax_bwd(probs, dp, ds, B * H * Sq, Sk, ldp, lds, ldp, scale)
    if need_dq:
        dq = torch.empty_like(q)
        gemm_raw(a=ds.data_ptr(), b=k.data_ptr(), d=dq.data_ptr(), ab_dtype=dt, d_dtype=dt, a_mn_major=0, b_mn_major=1,
                 m=Sq, n=hd, k=Sk, a_ld=ldp, a_s2=Sq * ldp, a_s3=H * Sq * ldp, a_z2=H, b_ld=D, b_s2=hd, b_s3=Sk * D,
                 b_z2=H, d_ld=D, d_s2=hd, d_s3=Sq * D, z_lo=H, z_hi=B, a_div=1, a_mul=1, a_seg=0, b_div=1, b_mul=1,
                 b_seg=0, k_segs=1, alpha=1.0)
    if need_dkv:
        dk = torch.empty_like(k)
        gemm_raw(a=ds.data_ptr(), b=q.data_ptr(), d=dk.data_ptr(), ab_dtype=dt, d_dtype=dt, a_mn_major=1, b_mn_major=1,
                 m=Sk, n=hd, k=Sq, a_ld=ldp, a_s2=Sq * ldp, a_s3=H * Sq * ldp, a_z2=H, b_ld=D, b_s2=hd, b_s3=Sq * D,
                 b_z2=H, d_ld=D, d_s2=hd, d_s3=Sk * D, z_lo=H, z_hi=B, a_div=1, a_mul=1, a_seg=0, b_div=1, b_mul=1,
                 b_seg=0, k_segs=1, alpha=1.0)
    return dq, dk, dv
