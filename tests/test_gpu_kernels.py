"""GPU parity tests of the individual kernels (through the C-ABI) against plain fp32 PyTorch on the
same inputs.  Tolerances: bf16 outputs 2^-7 relative (+ sqrt(K)-scaled absolute), tf32 2^-9."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"


@pytest.fixture(scope="module", autouse=True)
def _strict_fp32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


def ops():
    from dexbotic_b200 import ops as o
    return o


def _rand(shape, dtype, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g, dtype=torch.float32) * scale).to(dtype).to(DEV)


def _check(out, ref, K, dtype, what=""):
    out = out.float()
    ref = ref.float()
    rel = 2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -9
    atol = rel * math.sqrt(max(K, 1)) * 2.0
    err = (out - ref).abs()
    tol = atol + rel * ref.abs()
    bad = (err > tol).sum().item()
    assert bad == 0, (f"{what}: {bad}/{err.numel()} mismatches, max_err={err.max().item():.4g} "
                      f"ref_absmax={ref.abs().max().item():.4g} at {torch.nonzero(err > tol)[:4].tolist()}")


GEMM_SHAPES = [(128, 256, 64), (256, 512, 256), (300, 200, 136), (1000, 3584, 1024), (77, 72, 96), (129, 257, 65 * 8),
               (2176, 1152, 384), (1300, 768, 2048 + 64)]   # >= 1024 rows: CTA-pair (cta_group::2) path


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_gemm_layouts(M, N, K, a_mn, b_mn, dtype):
    o = ops()
    # pad leading dims to 16-byte multiples the way real activations are laid out
    es = 2 if dtype == torch.bfloat16 else 4
    al = 16 // es

    def mk(rows, cols, seed):
        ld = (cols + al - 1) // al * al
        return _rand((rows, ld), dtype, seed)[:, :cols]

    a = mk(K, M, 1) if a_mn else mk(M, K, 1)
    b = mk(K, N, 2) if b_mn else mk(N, K, 2)
    A = a.float().t() if a_mn else a.float()
    Bm = b.float() if b_mn else b.float().t()
    if dtype == torch.float32:  # tf32 truncates inputs: compare against fp64-free fp32 matmul with loose tol
        pass
    ref = A @ Bm
    out = o.gemm(a, b, a_mn=a_mn, b_mn=b_mn)
    _check(out, ref, K, dtype, f"gemm {M}x{N}x{K} a_mn={a_mn} b_mn={b_mn} {dtype}")


@pytest.mark.parametrize("block_n", [64, 128, 256])
def test_gemm_block_n(block_n):
    o = ops()
    a, b = _rand((520, 320), torch.bfloat16, 3), _rand((384, 320), torch.bfloat16, 4)
    out = o.gemm(a, b, block_n=block_n)
    _check(out, a.float() @ b.float().t(), 320, torch.bfloat16, f"block_n={block_n}")


@pytest.mark.parametrize("act", ["none", "gelu", "gelu_tanh", "quick_gelu", "silu", "relu"])
@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("M", [333, 1290])
def test_gemm_epilogue(act, out_dtype, M):
    o = ops()
    N, K = 264, 192
    a, b = _rand((M, K), torch.bfloat16, 5), _rand((N, K), torch.bfloat16, 6, 0.2)
    bias = _rand((N,), torch.bfloat16, 7)
    res = _rand((M, N), torch.bfloat16, 8)
    aux = torch.empty((M, N), device=DEV, dtype=out_dtype)
    out = o.gemm(a, b, bias=bias, residual=res, act=act, alpha=0.5, aux=aux, out_dtype=out_dtype)
    pre = 0.5 * (a.float() @ b.float().t()) + bias.float()
    f = {"none": lambda x: x, "gelu": torch.nn.functional.gelu,
         "gelu_tanh": lambda x: torch.nn.functional.gelu(x, approximate="tanh"),
         "quick_gelu": lambda x: x * torch.sigmoid(1.702 * x), "silu": torch.nn.functional.silu,
         "relu": torch.relu}[act]
    _check(aux, pre, K, out_dtype, f"aux {act}")
    _check(out, f(pre) + res.float(), K, out_dtype, f"epilogue {act}")


def test_gemm_accumulate_fp32():
    o = ops()
    M, N, K = 256, 384, 520
    a, b = _rand((K, M), torch.bfloat16, 9), _rand((K, N), torch.bfloat16, 10)
    acc = _rand((M, N), torch.float32, 11)
    ref = acc + a.float().t() @ b.float()
    o.gemm(a, b, a_mn=True, b_mn=True, out=acc, residual=acc)
    _check(acc, ref, K, torch.bfloat16, "wgrad accumulate")


@pytest.mark.parametrize("act", ["silu", "gelu_tanh"])
@pytest.mark.parametrize("M", [300, 1411])
def test_gemm_dual(act, M):
    o = ops()
    N, K = 648, 256
    a = _rand((M, K), torch.bfloat16, 12)
    wg, wu = _rand((N, K), torch.bfloat16, 13, 0.1), _rand((N, K), torch.bfloat16, 14, 0.1)
    ag = torch.empty((M, N), device=DEV, dtype=torch.bfloat16)
    au = torch.empty_like(ag)
    out = o.gemm_dual(a, wg, wu, act, aux_gate=ag, aux_up=au)
    g, u = a.float() @ wg.float().t(), a.float() @ wu.float().t()
    f = torch.nn.functional.silu if act == "silu" else (lambda x: torch.nn.functional.gelu(x, approximate="tanh"))
    _check(ag, g, K, torch.bfloat16, "dual aux gate")
    _check(au, u, K, torch.bfloat16, "dual aux up")
    _check(out, f(g) * u, K, torch.bfloat16, "dual out")


@pytest.mark.parametrize("act", ["silu", "gelu_tanh"])
@pytest.mark.parametrize("M", [300, 1411])
def test_gemm_glu_bwd(act, M):
    """Down-projection dgrad with the GLU backward in its epilogue (dh = dy @ W_down never stored): dg = dh * u * act'(g),
    du = dh * act(g), against torch autograd in fp32; out of place and in place over g / u."""
    o = ops()
    inter, d = 648 + 120, 256
    dy = _rand((M, d), torch.bfloat16, 15)
    wd = _rand((d, inter), torch.bfloat16, 16, 0.1)
    g, u = _rand((M, inter), torch.bfloat16, 17), _rand((M, inter), torch.bfloat16, 18)
    gr, ur = g.float().requires_grad_(True), u.float().requires_grad_(True)
    f = torch.nn.functional.silu if act == "silu" else (lambda x: torch.nn.functional.gelu(x, approximate="tanh"))
    dh = dy.float() @ wd.float()
    (f(gr) * ur).backward(dh)
    dg, du = o.gemm_glu_bwd(dy, wd, g, u, act)
    _check(dg, gr.grad, d, torch.bfloat16, "glu_bwd dg")
    _check(du, ur.grad, d, torch.bfloat16, "glu_bwd du")
    g2, u2 = g.clone(), u.clone()
    o.gemm_glu_bwd(dy, wd, g2, u2, act, dg=g2, du=u2)
    assert torch.equal(g2, dg) and torch.equal(u2, du)


@pytest.mark.parametrize("M,N,K", [(33, 7, 384), (2176, 384, 7), (5, 768, 256)])
def test_gemm_simt_odd_shapes(M, N, K):
    o = ops()
    a, b = _rand((M, K), torch.float32, 15), _rand((N, K), torch.float32, 16)
    bias = _rand((N,), torch.float32, 17)
    out = o.gemm(a, b, bias=bias)
    _check(out, a @ b.t() + bias, K, torch.float32, "simt")


# ------------------------------------------------------------------ attention
def _ref_attention(qkv, B, S, H, KVH, hd, keymask, bid):
    x = qkv.float().view(B, S, H + 2 * KVH, hd)
    q, k, v = x[:, :, :H], x[:, :, H:H + KVH], x[:, :, H + KVH:]
    G = H // KVH
    k = k.repeat_interleave(G, dim=2)
    v = v.repeat_interleave(G, dim=2)
    s = torch.einsum("bqhd,bkhd->bhqk", q, k) * hd ** -0.5
    allow = torch.ones(B, 1, S, S, dtype=torch.bool, device=qkv.device)
    if keymask is not None:
        allow = allow & keymask.bool()[:, None, None, :]
    if bid is not None:
        allow = allow & (bid[:, None, None, :] <= bid[:, None, :, None])
    s = s.masked_fill(~allow, float("-inf"))
    p = torch.softmax(s, dim=-1)
    p = torch.nan_to_num(p, nan=0.0)
    o = torch.einsum("bhqk,bkhd->bqhd", p, v).reshape(B, S, H * hd)
    return o


@pytest.mark.parametrize("B,S,H,KVH,hd,causal,dtype", [
    (2, 77, 4, 2, 128, True, torch.bfloat16),
    (2, 308, 28, 4, 128, True, torch.bfloat16),
    (3, 257, 16, 16, 64, False, torch.bfloat16),
    (4, 17, 4, 4, 96, False, torch.float32),
    (2, 50, 8, 1, 256, True, torch.bfloat16),
    (1, 1100, 4, 2, 64, True, torch.bfloat16),     # S >= 1024: batched GEMMs take the CTA-pair path
    (2, 365, 28, 4, 128, True, torch.bfloat16),    # OFT length: 384 + 128 TMEM columns
    (2, 512, 4, 4, 72, False, torch.bfloat16),     # SigLIP head_dim 72, Sk = 512 (TMEM full)
    (2, 130, 2, 1, 96, True, torch.bfloat16),
    (2, 867, 8, 1, 256, True, torch.bfloat16),     # pi0 joint sequence: Gemma head_dim 256, 8 q heads on 1 kv head
    (2, 256, 16, 16, 72, False, torch.bfloat16),   # SigLIP-So400m tower
])
def test_attention_fwd_bwd(B, S, H, KVH, hd, causal, dtype):
    o = ops()
    W = (H + 2 * KVH) * hd
    qkv = _rand((B, S, W), dtype, 20, 0.5)
    lens = torch.tensor([S - (3 * i) % max(S // 2, 1) for i in range(B)], device=DEV)
    keymask = (torch.arange(S, device=DEV)[None, :] < lens[:, None]).to(torch.uint8) if causal else None
    bid = torch.arange(S, device=DEV, dtype=torch.int32)[None, :].expand(B, S).contiguous() if causal else None
    sh = o.AttnShape(B, S, H, KVH, hd, dtype)
    if causal and S % 2 == 1:      # exercise the index-causal fast path as well as the block-id rule
        out, probs = o.attention_fwd(qkv, sh, keymask=keymask, causal=True)
    else:
        out, probs = o.attention_fwd(qkv, sh, keymask=keymask, bid_q=bid, bid_k=bid)

    qkv_ref = qkv.float().requires_grad_(True)
    ref = _ref_attention(qkv_ref, B, S, H, KVH, hd, keymask, bid)
    rowmask = keymask.bool()[:, :, None] if keymask is not None else torch.ones(B, S, 1, dtype=torch.bool, device=DEV)
    _check(out * rowmask, ref * rowmask, S, dtype, "attention fwd")

    dout = _rand((B, S, H * hd), dtype, 21) * rowmask.to(dtype)
    idx_causal = bool(causal and S % 2 == 1)
    dqkv = o.attention_bwd(dout, qkv, probs, sh, causal=idx_causal, out=out, keymask=keymask,
                           bid_q=None if idx_causal else bid, bid_k=None if idx_causal else bid)
    (ref * rowmask).backward(dout.float())
    _check(dqkv, qkv_ref.grad, S * 4, dtype, "attention bwd")


# ---------------------------------------------------------------- elementwise
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("unit_offset", [False, True])
def test_rmsnorm(dtype, unit_offset):
    o = ops()
    M, D = 517, 3584
    x, w = _rand((M, D), dtype, 30), _rand((D,), dtype, 31, 0.3)
    y, rstd = o.rmsnorm_fwd(x, w, 1e-6, unit_offset)
    xr = x.float().requires_grad_(True)
    wr = w.float().requires_grad_(True)
    n = xr * torch.rsqrt(xr.pow(2).mean(-1, keepdim=True) + 1e-6)
    ref = n * (1 + wr) if unit_offset else wr * n
    _check(y, ref, 1, dtype, "rmsnorm fwd")
    dy = _rand((M, D), dtype, 32)
    dw = torch.zeros(D, device=DEV, dtype=torch.float32)
    dx = o.rmsnorm_bwd(dy, x, w, rstd, unit_offset, dw=dw)
    ref.backward(dy.float())
    _check(dx, xr.grad, 4, dtype, "rmsnorm dx")
    _check(dw, wr.grad, M, dtype, "rmsnorm dw")


@pytest.mark.parametrize("D", [256, 1024, 2048, 3584])
@pytest.mark.parametrize("accumulate", [False, True])
def test_rmsnorm_staged_ring_matches_register_kernels(D, accumulate):
    """The norm kernels that stage rows in shared memory through the bulk-copy engine (default) against the
    register-prefetch kernels and fp32 torch, on enough rows that every block walks its mbarrier ring several times
    (both parities of every stage), with and without the accumulate-into-dx variant."""
    o = ops()
    from dexbotic_b200 import _lib
    lib = _lib.load()
    M = 6007
    dtype = torch.bfloat16
    x, w = _rand((M, D), dtype, 40), _rand((D,), dtype, 41, 0.3)
    dy = _rand((M, D), dtype, 42)
    dx0 = _rand((M, D), dtype, 43)
    res = {}
    try:
        for staged in (3, 0):          # bit 0: backward staged, bit 1: forward staged
            lib.b200_set_norm_staged(staged)
            y, rstd = o.rmsnorm_fwd(x, w, 1e-6)
            dw = torch.zeros(D, device=DEV, dtype=torch.float32)
            dx = dx0.clone() if accumulate else torch.empty_like(x)
            o.rmsnorm_bwd(dy, x, w, rstd, dx=dx, dw=dw, accumulate_dx=accumulate)
            torch.cuda.synchronize()
            res[staged] = (y, rstd, dx, dw)
    finally:
        lib.b200_set_norm_staged(1)
    res[1] = res[3]
    for a, b, name in zip(res[1][:3], res[0][:3], ("y", "rstd", "dx")):
        assert torch.equal(a, b), f"staged vs register kernel: {name} differs"      # same per-thread arithmetic
    torch.testing.assert_close(res[1][3], res[0][3], rtol=2e-4, atol=2e-3)         # dw: different partial grouping
    xr = x.float().requires_grad_(True)
    wr = w.float().requires_grad_(True)
    ref = wr * (xr * torch.rsqrt(xr.pow(2).mean(-1, keepdim=True) + 1e-6))
    _check(res[1][0], ref, 1, dtype, "staged rmsnorm fwd")
    ref.backward(dy.float())
    want_dx = xr.grad + (dx0.float() if accumulate else 0)
    _check(res[1][2], want_dx, 4, dtype, "staged rmsnorm dx")
    _check(res[1][3], wr.grad, M, dtype, "staged rmsnorm dw")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("affine", [False, True])
def test_layernorm(dtype, affine):
    o = ops()
    M, D = 301, 1024
    x = _rand((M, D), dtype, 33, 2.0) + 0.5
    w = _rand((D,), dtype, 34, 0.3) + 1 if affine else None
    b = _rand((D,), dtype, 35, 0.3) if affine else None
    y, mean, rstd = o.layernorm_fwd(x, w, b, 1e-5)
    xr = x.float().requires_grad_(True)
    wr = w.float().requires_grad_(True) if affine else None
    br = b.float().requires_grad_(True) if affine else None
    ref = torch.nn.functional.layer_norm(xr, (D,), wr, br, 1e-5)
    _check(y, ref, 1, dtype, "layernorm fwd")
    dy = _rand((M, D), dtype, 36)
    dw = torch.zeros(D, device=DEV) if affine else None
    db = torch.zeros(D, device=DEV) if affine else None
    dx = o.layernorm_bwd(dy, x, w, mean, rstd, dw=dw, db=db)
    ref.backward(dy.float())
    _check(dx, xr.grad, 8, dtype, "layernorm dx")
    if affine:
        _check(dw, wr.grad, M, dtype, "layernorm dw")
        _check(db, br.grad, M, dtype, "layernorm db")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("D", [384, 768, 1024, 1152, 1280, 1536])
@pytest.mark.parametrize("accumulate", [False, True])
def test_layernorm_warp_and_block_kernels(dtype, D, accumulate):
    """LayerNorm with one warp per row + the dw / db column kernel (`b200_set_norm_staged` bit 2, opt-in; D <= 1280 — 1536
    takes the block-per-row kernels either way) and the default block-per-row kernels, on more rows than warps are
    resident so that every warp walks several rows: both against fp32 torch, and against each other."""
    o = ops()
    from dexbotic_b200 import _lib
    lib = _lib.load()
    M = 9001 if dtype == torch.bfloat16 else 2177
    x = _rand((M, D), dtype, 50, 2.0) + 0.5
    w, b = _rand((D,), dtype, 51, 0.3) + 1, _rand((D,), dtype, 52, 0.3)
    dy = _rand((M, D), dtype, 53)
    dx0 = _rand((M, D), dtype, 54)
    res = {}
    try:
        for mode in (5, 1):
            lib.b200_set_norm_staged(mode)
            y, mean, rstd = o.layernorm_fwd(x, w, b, 1e-5)
            dw, db = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
            dx = dx0.clone() if accumulate else None
            dx = o.layernorm_bwd(dy, x, w, mean, rstd, dx=dx, dw=dw, db=db, accumulate_dx=accumulate)
            torch.cuda.synchronize()
            res[mode] = (y, mean, rstd, dx, dw, db)
    finally:
        lib.b200_set_norm_staged(1)          # the library default (staged rmsnorm backward only)
    xr = x.float().requires_grad_(True)
    wr, br = w.float().requires_grad_(True), b.float().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (D,), wr, br, 1e-5)
    ref.backward(dy.float())
    want_dx = xr.grad + (dx0.float() if accumulate else 0)
    for mode, (y, mean, rstd, dx, dw, db) in res.items():
        _check(y, ref, 1, dtype, f"layernorm fwd mode {mode}")
        torch.testing.assert_close(mean, x.float().mean(-1), rtol=1e-4, atol=1e-4)
        _check(dx, want_dx, 8, dtype, f"layernorm dx mode {mode}")
        _check(dw, wr.grad, M, dtype, f"layernorm dw mode {mode}")
        _check(db, br.grad, M, dtype, f"layernorm db mode {mode}")
    torch.testing.assert_close(res[5][2], res[1][2], rtol=1e-5, atol=1e-6)         # rstd: same statistics
    torch.testing.assert_close(res[5][4], res[1][4], rtol=1e-3, atol=2e-2)         # dw: different summation grouping


def _rope_tables(n_pos, hd, theta=1e6):
    inv = 1.0 / (theta ** (torch.arange(0, hd, 2, dtype=torch.float32) / hd))
    f = torch.arange(n_pos, dtype=torch.float32)[:, None] * inv[None, :]
    return f.cos().to(DEV).contiguous(), f.sin().to(DEV).contiguous()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_rope_roundtrip_and_ref(dtype):
    o = ops()
    B, S, H, KVH, hd = 2, 40, 4, 2, 128
    W = (H + 2 * KVH) * hd
    qkv = _rand((B, S, W), dtype, 40)
    pos = torch.arange(S, device=DEV, dtype=torch.int32).repeat(B)
    cos, sin = _rope_tables(64, hd)
    x = qkv.clone()
    o.rope_(x, pos, cos, sin, H + KVH, hd)
    xr = qkv.float().view(B, S, H + 2 * KVH, hd)
    c = cos[pos.long()].to(dtype).float().view(B, S, 1, hd // 2)
    s = sin[pos.long()].to(dtype).float().view(B, S, 1, hd // 2)
    lo, hi = xr[..., :hd // 2], xr[..., hd // 2:]
    rot = torch.cat([lo * c - hi * s, hi * c + lo * s], -1)
    ref = torch.cat([rot[:, :, :H + KVH], xr[:, :, H + KVH:]], 2).reshape(B, S, W)
    _check(x, ref, 1, dtype, "rope fwd")
    o.rope_(x, pos, cos, sin, H + KVH, hd, inverse=True)
    _check(x, qkv, 4, dtype, "rope inverse round trip")


@pytest.mark.parametrize("act", ["gelu", "gelu_tanh", "quick_gelu", "silu"])
def test_act_and_glu(act):
    o = ops()
    n = 8 * 1001
    f = {"gelu": torch.nn.functional.gelu, "gelu_tanh": lambda x: torch.nn.functional.gelu(x, approximate="tanh"),
         "quick_gelu": lambda x: x * torch.sigmoid(1.702 * x), "silu": torch.nn.functional.silu}[act]
    x, u, dy = _rand((n,), torch.bfloat16, 50, 2.0), _rand((n,), torch.bfloat16, 51), _rand((n,), torch.bfloat16, 52)
    xr, ur = x.float().requires_grad_(True), u.float().requires_grad_(True)
    _check(o.act_fwd(x, act), f(xr), 1, torch.bfloat16, "act fwd")
    f(xr).backward(dy.float())
    _check(o.act_bwd(dy, x, act), xr.grad, 1, torch.bfloat16, "act bwd")
    xr.grad = None
    h = f(xr) * ur
    _check(o.glu_fwd(x, u, act), h, 1, torch.bfloat16, "glu fwd")
    h.backward(dy.float())
    hout = torch.empty_like(x)
    dg, du = o.glu_bwd(dy, x, u, act, h_out=hout)
    _check(dg, xr.grad, 2, torch.bfloat16, "glu dg")
    _check(du, ur.grad, 2, torch.bfloat16, "glu du")
    _check(hout, h, 1, torch.bfloat16, "glu h recompute")


def test_colsum_sumsq_mse_cast():
    o = ops()
    x = _rand((777, 264), torch.bfloat16, 60)
    out = torch.zeros(264, device=DEV)
    o.colsum_(x, out)
    _check(out, x.float().sum(0), 777, torch.bfloat16, "colsum")
    ss = torch.zeros((), device=DEV)
    flat = _rand((100003,), torch.bfloat16, 61)
    o.sumsq_(flat, ss)
    assert abs(ss.item() - flat.float().pow(2).sum().item()) < 1e-3 * ss.item()
    a, b = _rand((128, 16, 7), torch.float32, 62), _rand((128, 16, 7), torch.float32, 63)
    loss = torch.zeros((), device=DEV)
    o.mse_fwd(a, b, loss)
    assert abs(loss.item() - ((a - b) ** 2).mean().item()) < 1e-5
    g = torch.full((), 0.5, device=DEV)
    da = o.mse_bwd(a, b, g)
    _check(da, 0.5 * 2 * (a - b) / a.numel(), 1, torch.float32, "mse bwd")
    f32 = _rand((12347,), torch.float32, 64)
    bf = torch.empty(12347, device=DEV, dtype=torch.bfloat16)
    o.cast_(f32, bf)
    assert torch.equal(bf, f32.to(torch.bfloat16))


def test_adamw_matches_torch():
    o = ops()
    n = 100003
    p = _rand((n,), torch.float32, 70)
    ref_p = torch.nn.Parameter(p.clone())
    opt = torch.optim.AdamW([ref_p], lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    shadow = torch.empty(n, device=DEV, dtype=torch.bfloat16)
    clip = torch.full((), 0.5, device=DEV)
    for step in range(1, 4):
        g = _rand((n,), torch.bfloat16, 70 + step)
        ref_p.grad = g.float() * 0.5
        opt.step()
        o.adamw_(p, g, m, v, shadow, 1e-3, 0.9, 0.95, 1e-8, 0.1, step, clip)
    assert (p - ref_p.data).abs().max().item() < 1e-5
    assert torch.equal(shadow, p.to(torch.bfloat16))


def test_adamw_device_scalars_equal_host_scalars():
    """b200_adamw_dev reads the seven scalars from device memory (a captured optimizer is replayed with new learning
    rates): bit-identical to b200_adamw; step 0 is the identity update."""
    o = ops()
    n = 50021
    p1 = _rand((n,), torch.float32, 80)
    p2 = p1.clone()
    m1, v1, m2, v2 = (torch.zeros_like(p1) for _ in range(4))
    s1 = torch.empty(n, device=DEV, dtype=torch.bfloat16)
    s2 = torch.empty_like(s1)
    clip = torch.full((), 0.7, device=DEV)
    host = torch.zeros(8)
    for step in range(1, 4):
        g = _rand((n,), torch.bfloat16, 80 + step)
        lr = 1e-3 / step
        o.adamw_(p1, g, m1, v1, s1, lr, 0.9, 0.95, 1e-8, 0.1, step, clip)
        o.adamw_hyper_(host, lr, 0.9, 0.95, 1e-8, 0.1, step)
        o.adamw_dev_(p2, g, m2, v2, s2, host.cuda(), clip)
        assert torch.equal(p1, p2) and torch.equal(m1, m2) and torch.equal(v1, v2) and torch.equal(s1, s2)
    o.adamw_hyper_(host, 1e-3, 0.9, 0.95, 1e-8, 0.1, 0)
    before = (p2.clone(), m2.clone(), v2.clone())
    o.adamw_dev_(p2, g, m2, v2, s2, host.cuda(), clip)
    assert torch.equal(p2, before[0]) and torch.equal(m2, before[1]) and torch.equal(v2, before[2])


# --------------------------------------------------------------- index kernels
def test_splice_plan_gather_scatter():
    o = ops()
    B, L, P, D, V = 4, 12, 5, 64, 50
    ids = torch.randint(1, V, (B, L), generator=torch.Generator().manual_seed(80))
    ids[0, 1] = -200
    ids[1, 0] = -200
    ids[1, 5] = -200
    ids[3, 11] = -200  # sample 2 has no image
    mask = torch.ones(B, L, dtype=torch.uint8)
    mask[0, 9:] = 0
    mask[2, 6:] = 0
    labels = torch.randint(0, V, (B, L), generator=torch.Generator().manual_seed(81))
    ids_d, mask_d, labels_d = ids.to(DEV), mask.to(DEV), labels.to(DEV)
    lengths = o.splice_lengths(ids_d, mask_d, P, 18).cpu().tolist()
    exp = []
    for b in range(B):
        n = sum((P if ids[b, i] == -200 else 1) for i in range(L) if mask[b, i])
        exp.append(min(n, 18))
    assert lengths == exp
    S = max(exp)
    for left in (False, True):
        src, nl, nm, pos = o.splice_plan(ids_d, mask_d, labels_d, P, 18, S, left)
        table = _rand((V, D), torch.bfloat16, 82)
        feats = _rand((6 * P, D), torch.bfloat16, 83)
        emb = o.splice_gather(src, table, feats)
        img = 0
        for b in range(B):
            rows, labs = [], []
            n_img = 0
            for i in range(L):
                if not mask[b, i]:
                    continue
                if ids[b, i] == -200:
                    rows += [feats[(img + n_img) * P + t] for t in range(P)]
                    labs += [-100] * P
                    n_img += 1
                else:
                    rows.append(table[ids[b, i]])
                    labs.append(int(labels[b, i]))
            img += max(n_img, 1)
            rows, labs = rows[:18], labs[:18]
            n = len(rows)
            off = S - n if left else 0
            got = emb[b, off:off + n]
            assert torch.equal(got, torch.stack(rows)), f"sample {b} left={left}"
            assert nl[b, off:off + n].cpu().tolist() == labs
            assert nm[b].cpu().tolist() == [0] * off + [1] * n + [0] * (S - n - off)
            assert pos[b, off:off + n].cpu().tolist() == list(range(n))
            pad = torch.cat([emb[b, :off], emb[b, off + n:]])
            assert pad.abs().sum().item() == 0
        # backward
        dout = _rand((B, S, D), torch.bfloat16, 84)
        d_table = torch.zeros_like(table)
        d_feats = torch.zeros_like(feats)
        o.splice_scatter(src, dout, d_table, d_feats)
        ref_t = torch.zeros(V, D, device=DEV)
        ref_f = torch.zeros(6 * P, D, device=DEV)
        sflat, dflat = src.view(-1).cpu().tolist(), dout.view(-1, D).float()
        for r, s in enumerate(sflat):
            if s >= 0:
                ref_t[s] += dflat[r]
            elif s != -(2 ** 31):
                ref_f[-1 - s] += dflat[r]
        _check(d_table, ref_t, 4, torch.bfloat16, "splice d_table")
        _check(d_feats, ref_f, 1, torch.bfloat16, "splice d_feats")


@pytest.mark.parametrize("rows,vocab", [(3000, 7), (4096, 300), (257, 1)])
def test_splice_scatter_duplicate_tokens(rows, vocab):
    """Token rows that repeat: the first occurrence sums all occurrences in ascending row order (fp32) and adds once.
    vocab 7 / 1 give more than 1024 repeats of a token (the serial fallback), vocab 300 about a dozen (the sorted
    shared-memory list).  Checked against an fp64 index_add and for run-to-run bit equality."""
    o = ops()
    D = 256
    g = torch.Generator().manual_seed(91)
    tok = torch.randint(0, vocab, (rows,), generator=g)
    kind = torch.rand(rows, generator=g)
    src = tok.clone().to(torch.int32)
    n_img = int((kind < 0.3).sum())
    src[kind < 0.3] = -1 - torch.arange(n_img, dtype=torch.int32)          # image rows
    src[kind > 0.95] = -(2 ** 31)                                           # padding rows
    src = src.to(DEV)
    dout = _rand((rows, D), torch.bfloat16, 92)
    outs = []
    for _ in range(2):
        d_table = torch.zeros(vocab, D, device=DEV, dtype=torch.bfloat16)
        d_feats = torch.zeros(max(n_img, 1), D, device=DEV, dtype=torch.bfloat16)
        o.splice_scatter(src.view(1, -1), dout.view(1, rows, D), d_table, d_feats)
        outs.append((d_table, d_feats))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    is_tok = src >= 0
    ref_t = torch.zeros(vocab, D, device=DEV, dtype=torch.float64)
    ref_t.index_add_(0, src[is_tok].long(), dout[is_tok].double())
    cnt = torch.bincount(src[is_tok].long(), minlength=vocab).max().item()
    _check(outs[0][0], ref_t, cnt, torch.bfloat16, "splice_scatter d_table with duplicates")
    is_img = (src < 0) & (src != -(2 ** 31))
    assert torch.equal(outs[0][1][: n_img][(-1 - src[is_img]).long()], dout[is_img])


def test_gather_rows_and_last_valid():
    o = ops()
    B, S, D = 5, 33, 128
    x = _rand((B * S, D), torch.bfloat16, 90)
    mask = torch.zeros(B, S, dtype=torch.uint8)
    lens = [33, 1, 20, 7, 12]
    for b, n in enumerate(lens):
        mask[b, :n] = 1
    mask[3, :] = 0
    mask[3, 26:] = 1  # left padded
    idx = o.last_valid_index(mask.to(DEV))
    exp = [b * S + (S - 1 if b == 3 else lens[b] - 1) for b in range(B)]
    assert idx.cpu().tolist() == exp
    rep = idx.repeat(4)
    out = o.gather_rows(x, rep)
    assert torch.equal(out, x[rep.long()])
    dx = torch.zeros_like(x)
    dout = _rand((rep.numel(), D), torch.bfloat16, 91)
    o.scatter_rows_add_(dout, rep, dx)
    ref = torch.zeros(B * S, D, device=DEV)
    ref.index_add_(0, rep.long(), dout.float())
    _check(dx, ref, 4, torch.bfloat16, "scatter_rows_add")


def test_q_sample_and_timestep_embedding():
    o = ops()
    B = 128
    x, nz = _rand((B, 16, 7), torch.float32, 100), _rand((B, 16, 7), torch.float32, 101)
    t = torch.randint(0, 100, (B,), generator=torch.Generator().manual_seed(102)).to(torch.int32).to(DEV)
    sa = torch.linspace(0.99, 0.01, 100, device=DEV)
    sb = (1 - sa * sa).sqrt()
    xt = o.q_sample(x, nz, t, sa, sb)
    ref = sa[t.long()][:, None, None] * x + sb[t.long()][:, None, None] * nz
    _check(xt, ref, 1, torch.float32, "q_sample")
    emb = o.timestep_embedding(t.float(), 256, torch.float32)
    half = 128
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32, device=DEV) / half)
    args = t.float()[:, None] * freqs[None]
    ref = torch.cat([args.cos(), args.sin()], -1)
    assert (emb - ref).abs().max().item() < 2e-4


def test_discrete_tokenizer_bit_exact():
    o = ops()
    g = torch.Generator().manual_seed(110)
    a = torch.cat([torch.rand(20000, generator=g) * 2.4 - 1.2,
                   torch.tensor([-1.0, 1.0, 0.0, 1 / 255, -1 / 255, 0.5 / 255, 1.5 / 255, 2.5 / 255, 0.003921569])])
    # exact half-way points under the x255 map
    k = torch.arange(0, 255, dtype=torch.float32)
    a = torch.cat([a, (k + 0.5) / 255 * 2 - 1]).to(DEV)
    bins = o.discretize_actions(a, 256)
    ref = ((torch.clamp(a, -1, 1) + 1) / 2 * 255).round().long()
    assert torch.equal(bins, ref)
    cont = o.bins_to_continuous(bins, 256)
    ref_c = (bins.float().cpu() / 255) * 2 - 1   # CPU torch = true division, as restated in the oracle
    assert torch.equal(cont.cpu(), ref_c)
    logits = _rand((56 * 4, 1000), torch.bfloat16, 111)
    logits[3, -255:] = 0.0          # all ties -> first index
    logits[5, -1] = 100.0
    logits[6, -255] = 100.0
    idx = o.argmax_last(logits, 255)
    assert torch.equal(idx, torch.argmax(logits[:, -255:].float(), dim=-1))
    assert idx[3].item() == 0 and idx[5].item() == 254 and idx[6].item() == 0


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_cross_entropy(dtype):
    o = ops()
    rows, V = 112, 4096 + 8
    logits = _rand((rows, V), dtype, 120, 3.0)
    labels = torch.randint(0, V, (rows,), generator=torch.Generator().manual_seed(121)).to(DEV)
    labels[::7] = -100
    loss_sum, n_valid, lse = o.cross_entropy_fwd(logits, labels)
    lr = logits.float().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(lr, labels, ignore_index=-100)
    got = loss_sum / n_valid
    assert abs(got.item() - ref.item()) < 1e-4 * max(1.0, abs(ref.item()))
    ref.backward()
    d = o.cross_entropy_bwd(logits, labels, lse, n_valid)
    _check(d, lr.grad, 1, dtype, "ce bwd")
