"""Parity at BASELINE.json's FULL sizes (cfg-2: CogACT, Qwen2.5-7B widths, batch 32, S = 309 -> 9888 token rows),
where the CPU oracle cannot run in seconds: size-independent properties of the same CUDA path the bench times —
linearity of the GEMMs, stochasticity / mask support of the attention probabilities, conservation in the splice
scatter, idempotence of the integer action tokenizer, zero-sum cross-entropy gradients, fixed points of AdamW, and
run-to-run reproducibility + exact power-of-two homogeneity of the whole 28-layer forward/backward.  Sampled entries are
additionally compared with fp32/fp64 torch on the same inputs (tolerances as in test_gpu_kernels.py)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
B, S, D, INTER, H, KVH, HD, V = 32, 309, 3584, 18944, 28, 4, 128, 152064
M = B * S


def _rand(shape, dtype, seed, scale=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return (torch.randn(shape, device=DEV, generator=g) * scale).to(dtype)


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("name,m,n,k,a_mn,b_mn", [("gate fwd", M, INTER, D, False, False),
                                                   ("down dgrad", M, INTER, D, False, True),
                                                   ("gate wgrad", INTER, D, M, True, True)])
def test_gemm_linearity_and_samples_full_size(name, m, n, k, a_mn, b_mn):
    from dexbotic_b200 import ops
    a = _rand((k, m) if a_mn else (m, k), torch.bfloat16, 1)
    b1 = _rand((k, n) if b_mn else (n, k), torch.bfloat16, 2, 0.05)
    b2 = _rand((k, n) if b_mn else (n, k), torch.bfloat16, 3, 0.05)
    y1 = ops.gemm(a, b1, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32)
    y2 = ops.gemm(a, b2, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32)
    bs = (b1.float() + b2.float()).to(torch.bfloat16)
    ys = ops.gemm(a, bs, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32)
    # linearity in B: the only difference is the bf16 rounding of (b1 + b2), bounded by 2^-9 * |a|.|b| per term
    bound = (2.0 ** -8) * 0.07 * math.sqrt(k) * 4
    assert (ys - (y1 + y2)).abs().max().item() < bound, name
    # homogeneity: exact for powers of two
    y4 = ops.gemm(a, (b1.float() * 4).to(torch.bfloat16), a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32)
    assert torch.equal(y4, y1 * 4), name
    # sampled entries against fp64
    gi = torch.Generator(device=DEV).manual_seed(4)
    rows = torch.randint(0, m, (48,), device=DEV, generator=gi)
    cols = torch.randint(0, n, (48,), device=DEV, generator=gi)
    A = (a[:, rows].t() if a_mn else a[rows]).double()
    Bm = (b1[:, cols].t() if b_mn else b1[cols]).double()
    ref = A @ Bm.t()
    got = y1[rows][:, cols].double()
    assert (got - ref).abs().max().item() < 1e-3 * math.sqrt(k) * 0.05 * 8, name      # fp32 accumulation of bf16 products


# ------------------------------------------------------------------------------------------- attention
def test_attention_full_size():
    """Flash attention at the benchmark geometry: the saved log-sum-exp reproduces softmax rows that sum to 1 over the
    allowed keys only, O = P V on sampled (batch, head) pairs, and query rows with no visible key give exactly 0."""
    from dexbotic_b200 import ops
    W = (H + 2 * KVH) * HD
    qkv = _rand((B, S, W), torch.bfloat16, 5, 0.5)
    keymask = torch.ones(B, S, dtype=torch.uint8, device=DEV)
    keymask[3, 300:] = 0
    keymask[17, 280:] = 0
    keymask[9, :5] = 0                 # left padding: rows 0..4 of sample 9 see no key at all
    sh = ops.AttnShape(B, S, H, KVH, HD, torch.bfloat16)
    out, lse = ops.attention_fwd(qkv, sh, keymask=keymask, causal=True)
    assert lse.shape == (B, H, S) and lse.dtype == torch.float32
    q_idx = torch.arange(S, device=DEV)
    allowed = (q_idx[None, :] <= q_idx[:, None])[None, None] & keymask.bool()[:, None, None, :]
    assert out[9, :5].abs().max().item() == 0.0
    assert torch.isinf(lse[9, :, :5]).all()
    for b, h in ((0, 0), (3, 27), (17, 13), (31, 5), (9, 2)):
        q = qkv[b, :, h * HD:(h + 1) * HD].float()
        kv = h // (H // KVH)
        k = qkv[b, :, (H + kv) * HD:(H + kv + 1) * HD].float()
        v = qkv[b, :, (H + KVH + kv) * HD:(H + KVH + kv + 1) * HD].float()
        s = (q @ k.t()) * HD ** -0.5
        s = s.masked_fill(~allowed[b, 0], float("-inf"))
        rows = allowed[b, 0].any(-1)
        P = torch.exp2(s * 1.4426950408889634 - lse[b, h][:, None])        # the backward's recomputation
        assert (P[rows].sum(-1) - 1).abs().max().item() < 2e-3
        assert (P * (~allowed[b, 0])).abs().max().item() == 0.0
        ref = torch.softmax(s, -1) @ v
        got = out[b, :, h * HD:(h + 1) * HD].float()
        rel = ((got[rows] - ref[rows]).norm() / ref[rows].norm()).item()
        assert rel < 2e-2, (b, h, rel)


# -------------------------------------------------------------------------------------- splice / gather
def test_splice_scatter_conserves_gradient_mass_full_size():
    """sum(d_table) + sum(d_feats) == sum(dout over non-padding rows): every gradient row lands exactly once."""
    from dexbotic_b200 import ops
    P, L = 256, 54
    g = torch.Generator(device=DEV).manual_seed(6)
    ids = torch.randint(1000, 30000, (B, L), device=DEV, generator=g)
    ids[:, 1] = -200
    mask = torch.ones(B, L, dtype=torch.uint8, device=DEV)
    mask[5, 48:] = 0
    mask[9, 50:] = 0
    lengths = ops.splice_lengths(ids, mask, P, 0)
    Smax = int(lengths.max().item())
    assert Smax == L - 1 + P == S
    src, labels, new_mask, pos = ops.splice_plan(ids, mask, None, P, 0, Smax, False)
    assert int(new_mask.sum().item()) == int(lengths.sum().item())
    table = _rand((30000, D), torch.bfloat16, 7)
    feats = _rand((B * P, D), torch.bfloat16, 8)
    emb = ops.splice_gather(src, table, feats)
    # gather reproduces its sources bit-exactly
    flat = src.reshape(-1)
    tok = flat >= 0
    assert torch.equal(emb.view(-1, D)[tok], table[flat[tok].long()])
    img = (flat < 0) & (flat != -(2 ** 31))
    assert torch.equal(emb.view(-1, D)[img], feats[(-1 - flat[img]).long()])
    assert emb.view(-1, D)[flat == -(2 ** 31)].abs().max().item() == 0.0
    dout = _rand((B * Smax, D), torch.bfloat16, 9, 0.1)
    d_table = torch.zeros_like(table)
    d_feats = torch.zeros_like(feats)
    ops.splice_scatter(src, dout, d_table, d_feats)
    live = (flat != -(2 ** 31))
    want = dout.float()[live].sum(0)
    got = d_table.float().sum(0) + d_feats.float().sum(0)
    assert (got - want).abs().max().item() < 0.05 * math.sqrt(B * Smax) * 0.1 + 1e-2


# ----------------------------------------------------------------------------------------- integer path
def test_action_tokenizer_idempotent_full_size():
    """bins -> continuous -> bins is the identity on all 256 bins for 10^7 values (bit-exact integer path)."""
    from dexbotic_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(10)
    a = torch.rand(10_000_000, device=DEV, generator=g) * 2.4 - 1.2          # includes values clamped at +-1
    bins = ops.discretize_actions(a.contiguous(), 256)
    assert bins.dtype == torch.int64 and int(bins.min()) == 0 and int(bins.max()) == 255
    cont = ops.bins_to_continuous(bins, 256)
    assert torch.equal(ops.discretize_actions(cont.contiguous(), 256), bins)
    assert (cont.abs() <= 1).all() and sorted(torch.unique(bins).tolist()) == list(range(256))


def test_cross_entropy_gradient_rows_sum_to_zero_full_vocab():
    from dexbotic_b200 import ops
    rows = 32 * 56
    logits = _rand((rows, V), torch.bfloat16, 11, 2.0)
    g = torch.Generator(device=DEV).manual_seed(12)
    labels = torch.randint(V - 255, V, (rows,), device=DEV, generator=g)
    labels[::17] = -100
    loss_sum, n_valid, lse = ops.cross_entropy_fwd(logits, labels)
    assert int(n_valid.item()) == int((labels != -100).sum().item())
    idx = torch.arange(0, rows, 97, device=DEV)
    idx = idx[labels[idx] != -100]
    ref = torch.nn.functional.cross_entropy(logits[idx].float(), labels[idx], reduction="none")
    got = lse[idx] - logits[idx, labels[idx]].float()
    assert (got - ref).abs().max().item() < 2e-3
    grad = ops.cross_entropy_bwd(logits, labels, lse, n_valid)
    gs = grad.float().sum(-1)
    assert gs.abs().max().item() < 3e-3                      # softmax - onehot sums to 0 (bf16 gradient entries)
    assert grad[labels == -100].abs().max().item() == 0.0


# --------------------------------------------------------------------------------------------- AdamW
def test_adamw_fixed_points_and_reference_large():
    from dexbotic_b200 import ops
    n = 100_000_000
    p = _rand((n,), torch.float32, 13, 0.02)
    gq = _rand((n,), torch.bfloat16, 14, 1e-3)
    m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    sh = torch.empty(n, device=DEV, dtype=torch.bfloat16)
    p0 = p.clone()
    ops.adamw_(p, gq, m, v, sh, 0.0, 0.9, 0.999, 1e-8, 0.0, 1)           # lr = 0: weights are a fixed point
    assert torch.equal(p, p0) and torch.equal(sh, p0.to(torch.bfloat16))
    gf = gq.float()
    assert torch.allclose(m, 0.1 * gf, rtol=1e-6, atol=0) and torch.allclose(v, 0.001 * gf * gf, rtol=1e-5, atol=0)
    ops.adamw_(p, gq, m, v, sh, 1e-3, 0.9, 0.999, 1e-8, 0.1, 2)
    idx = torch.arange(0, n, 9973, device=DEV)
    m2 = 0.9 * (0.1 * gf[idx]) + 0.1 * gf[idx]
    v2 = 0.999 * (0.001 * gf[idx] ** 2) + 0.001 * gf[idx] ** 2
    step = (m2 / (1 - 0.9 ** 2)) / ((v2 / (1 - 0.999 ** 2)).sqrt() + 1e-8)
    ref = p0[idx] * (1 - 1e-3 * 0.1) - 1e-3 * step
    assert torch.allclose(p[idx], ref, rtol=2e-5, atol=1e-7)
    assert torch.equal(sh, p.to(torch.bfloat16))


# ---------------------------------------------------------------------------------------- whole model
def test_full_size_model_is_deterministic_and_homogeneous():
    """The 28-layer CogACT-7B step at the bench's shapes: repeated forwards give the same loss (to the last ulp of the
    atomically reduced scalar) and bit-identical gradients; backward of
    4 * loss gives exactly 4 x the decoder gradients of backward of loss (every backward kernel is linear in the
    incoming gradient and power-of-two scaling is exact in bf16 / fp32)."""
    import bench
    w = bench.WORKLOADS["cogact_7b"]
    model = bench.build_model(w, torch.device("cuda", 0))
    model.init_weights_(seed=1234)
    model.train()
    batch = {k: v.cuda() for k, v in bench.make_batch(w, 0, pinned=False).items()}
    g = torch.Generator(device=DEV).manual_seed(15)
    R = 4
    fixed = dict(noise=torch.randn(R * B, 16, 7, device=DEV, generator=g),
                 timesteps=torch.randint(0, 100, (R * B,), device=DEV, generator=g),
                 drop_mask=torch.rand(R * B, device=DEV, generator=g) < 0.1)
    names = ["model.llm.layers.27.mlp.down_proj.weight", "model.llm.layers.13.self_attn.q_proj.weight",
             "model.llm.layers.0.mlp.gate_proj.weight", "model.mm_projector.2.weight"]
    grads, losses = [], []
    for scale in (1.0, 4.0, 1.0):
        model.zero_grad()
        out = model(**batch, **fixed)
        assert out.logits.shape == (B, S, D)
        losses.append(out.loss.item())
        (out.loss * scale).backward()
        grads.append({n: model.store.g(n).clone() for n in names})
    # the scalar MSE reduction accumulates block partials with fp32 atomics: equal to ~1 ulp of the sum, not bitwise
    assert max(losses) - min(losses) < 1e-6 * abs(losses[0]), losses
    assert math.isfinite(losses[0]) and 0.1 < losses[0] < 10
    for n in names:
        assert torch.equal(grads[0][n], grads[2][n]), f"{n}: backward is not deterministic"
        assert torch.equal(grads[1][n].float(), grads[0][n].float() * 4), f"{n}: backward is not homogeneous"
        assert grads[0][n].float().abs().max().item() > 0
