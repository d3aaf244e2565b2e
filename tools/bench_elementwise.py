"""GB/s of the HBM-bound kernels on the CogACT-7B shapes (CUDA events, L2 flushed between reps)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dexbotic_b200 import ops  # noqa: E402
from tools.bench_gemm import timeit  # noqa: E402

PEAK = 6485.5


ONCE = "--once" in sys.argv      # one launch per kernel, no timing: the run that ncu replays (tools: see profiles/README.md)


def report(name, nbytes, fn):
    if ONCE:
        fn()
        torch.cuda.synchronize()
        print(f"{name:28s} algorithmic {nbytes/1e6:9.1f} MB", flush=True)
        return
    t = timeit(fn, reps=7)
    print(f"{name:28s} {t*1e3:8.1f} us  {nbytes/t/1e6:7.0f} GB/s  {100*nbytes/t/1e6/PEAK:5.1f}% of measured copy peak", flush=True)


def main():
    dev, bf = "cuda", torch.bfloat16
    M, d, I, W = 9888, 3584, 18944, 4608
    x = torch.randn(M, d, device=dev, dtype=bf)
    dy = torch.randn(M, d, device=dev, dtype=bf)
    w = torch.randn(d, device=dev, dtype=bf)
    y, rstd = ops.rmsnorm_fwd(x, w, 1e-6)
    dx = torch.empty_like(x)
    dw = torch.zeros(d, device=dev)
    from dexbotic_b200 import _lib
    for staged in (3, 0):          # 3 = rows staged in smem by the bulk-copy engine (fwd + bwd), 0 = register prefetch
        _lib.load().b200_set_norm_staged(staged)
        tag = " [staged]" if staged else " [register-prefetch kernel]"
        report("rmsnorm_fwd" + tag, 2 * M * d * 2, lambda: ops.rmsnorm_fwd(x, w, 1e-6, out=y))
        report("rmsnorm_bwd" + tag, 3 * M * d * 2, lambda: ops.rmsnorm_bwd(dy, x, w, rstd, dx=dx, dw=dw))
        report("rmsnorm_bwd(accum)" + tag, 4 * M * d * 2,
               lambda: ops.rmsnorm_bwd(dy, x, w, rstd, dx=dx, dw=dw, accumulate_dx=True))
    _lib.load().b200_set_norm_staged(1)
    g = torch.randn(M, I, device=dev, dtype=bf)
    u = torch.randn(M, I, device=dev, dtype=bf)
    h = torch.empty_like(g)
    dh = torch.randn(M, I, device=dev, dtype=bf)
    report("glu_fwd", 3 * M * I * 2, lambda: ops.glu_fwd(g, u, "silu", out=h))
    dg, du = torch.empty_like(g), torch.empty_like(g)
    report("glu_bwd(+h)", 6 * M * I * 2, lambda: ops.glu_bwd(dh, g, u, "silu", dg=dg, du=du, h_out=h))
    B, H, S = 32, 28, 309
    sh = ops.AttnShape(B, S, H, 4, 128, bf)
    sc = torch.randn(B, H, S, sh.ld_s, device=dev)
    pr = torch.empty(B, H, S, sh.ld_p, device=dev, dtype=bf)
    km = torch.ones(B, S, device=dev, dtype=torch.uint8)
    bid = torch.arange(S, device=dev, dtype=torch.int32)[None].expand(B, S).contiguous()
    nb = B * H * S * S * (4 + 2)
    report("softmax_fwd causal", nb, lambda: ops.softmax_fwd(sc, pr, S, S, heads=H, keymask=km, causal=True))
    report("softmax_fwd bid arrays", nb, lambda: ops.softmax_fwd(sc, pr, S, S, heads=H, keymask=km, bid_q=bid, bid_k=bid))
    report("softmax_fwd nomask", nb, lambda: ops.softmax_fwd(sc, pr, S, S, heads=H))
    ds = torch.empty_like(pr)
    report("softmax_bwd", B * H * S * S * (2 + 4 + 2), lambda: ops.softmax_bwd(pr, sc, ds, B * H * S, S, sh.ld_p, sh.ld_s, sh.ld_p, 0.1))
    qkv = torch.randn(M, W, device=dev, dtype=bf)
    out = torch.zeros(W, device=dev)
    report("colsum [M,4608]", M * W * 2, lambda: ops.colsum_(qkv, out))
    pos = torch.arange(M, device=dev, dtype=torch.int32) % S
    inv = 1.0 / (1e6 ** (torch.arange(0, 128, 2, dtype=torch.float32, device=dev) / 128))
    f = torch.arange(1024, dtype=torch.float32, device=dev)[:, None] * inv[None]
    cos, sin = f.cos().contiguous(), f.sin().contiguous()
    report("rope (q,k of qkv)", 2 * M * 32 * 128 * 2, lambda: ops.rope_(qkv, pos, cos, sin, 32, 128))
    xv = torch.randn(8224, 1024, device=dev, dtype=bf)
    wv, bv = torch.randn(1024, device=dev, dtype=bf), torch.randn(1024, device=dev, dtype=bf)
    yv, mean, rs = ops.layernorm_fwd(xv, wv, bv, 1e-5)
    dxv = torch.empty_like(xv)
    dwv, dbv = torch.zeros(1024, device=dev), torch.zeros(1024, device=dev)
    for mode in (1, 5):            # 1 = default (block per row), 5 = + one warp per row and a dw / db column kernel (opt-in)
        _lib.load().b200_set_norm_staged(mode)
        tag = " [warp per row]" if mode & 4 else ""
        report("layernorm_fwd [8224,1024]" + tag, 2 * xv.numel() * 2, lambda: ops.layernorm_fwd(xv, wv, bv, 1e-5, out=yv))
        report("layernorm_bwd [8224,1024]" + tag, 3 * xv.numel() * 2,
               lambda: ops.layernorm_bwd(xv, xv, wv, mean, rs, dx=dxv, dw=dwv, db=dbv))
    _lib.load().b200_set_norm_staged(1)
    n = 1 << 28
    p32, m32, v32 = (torch.zeros(n, device=dev) for _ in range(3))
    g16 = torch.zeros(n, device=dev, dtype=bf)
    s16 = torch.empty(n, device=dev, dtype=bf)
    report("adamw (268M params)", 28 * n, lambda: ops.adamw_(p32, g16, m32, v32, s16, 1e-4, 0.9, 0.999, 1e-8, 0.0, 1))
    ss = torch.zeros((), device=dev)
    report("sumsq bf16", 2 * n, lambda: ops.sumsq_(g16, ss))
    del p32, m32, v32, g16, s16
    # flash-attention row term and the image-token splice (CogACT-7B: 32 x 309 rows of 3584, 256 image rows each)
    o = torch.randn(B, S, H * 128, device=dev, dtype=bf)
    do = torch.randn(B, S, H * 128, device=dev, dtype=bf)
    qkv3 = torch.randn(B, S, (H + 8) * 128, device=dev, dtype=bf)
    out_f, lse = ops.flash_attention_fwd(qkv3, sh, keymask=km, causal=True)
    report("flash bwd (delta+dq+dkv)", 2 * o.numel() * 2, lambda: ops.flash_attention_bwd(do, qkv3, out_f, lse, sh, keymask=km, causal=True))
    table = torch.randn(152064, d, device=dev, dtype=bf)
    feats = torch.randn(B * 257, d, device=dev, dtype=bf)
    ids = torch.randint(1000, 30000, (B, 54), device=dev)
    ids[:, 1] = -200
    mask8 = torch.ones(B, 54, device=dev, dtype=torch.uint8)
    lengths = ops.splice_lengths(ids, mask8, 256, 0)
    Ssp = int(lengths.max().item())
    src, _, _, _ = ops.splice_plan(ids, mask8, None, 256, 0, Ssp, False)
    emb = torch.empty(B, Ssp, d, device=dev, dtype=bf)
    report("splice_gather", 2 * B * Ssp * d * 2, lambda: ops.splice_gather(src, table, feats, out=emb))


if __name__ == "__main__":
    main()
