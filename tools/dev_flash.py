"""Development check of the flash-attention kernels against an fp32 torch reference, one subprocess per case
(a trapped kernel poisons its CUDA context; the other cases still run).  Usage on the GPU box:
    python tools/dev_flash.py            # the case matrix
    python tools/dev_flash.py --time     # plus timings of the production geometries
"""
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

CASES = [
    # B, S, H, KVH, hd, mask  (mask: none | causal | causal+pad | bid | bid+pad)
    (1, 64, 1, 1, 64, "none"),
    (1, 128, 1, 1, 128, "none"),
    (1, 128, 2, 1, 128, "causal"),
    (2, 77, 4, 2, 128, "causal+pad"),
    (2, 309, 28, 4, 128, "causal+pad"),
    (3, 257, 16, 16, 64, "none"),
    (2, 256, 16, 16, 72, "none"),
    (2, 512, 4, 4, 72, "bid+pad"),
    (2, 50, 8, 1, 256, "causal"),
    (2, 867, 8, 1, 256, "bid+pad"),
    (1, 1100, 4, 2, 64, "causal+pad"),
    (2, 365, 28, 4, 128, "causal"),
    (2, 130, 2, 1, 96, "bid"),
    (2, 200, 6, 2, 32, "causal+pad"),
]


def ref_attention(qkv, B, S, H, KVH, hd, keymask, bid, causal):
    import torch
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
    if causal:
        i = torch.arange(S, device=qkv.device)
        allow = allow & (i[None, :] <= i[:, None])[None, None]
    s = s.masked_fill(~allow, float("-inf"))
    lse = torch.logsumexp(s, dim=-1)
    p = torch.nan_to_num(torch.softmax(s, dim=-1), nan=0.0)
    o = torch.einsum("bhqk,bkhd->bqhd", p, v).reshape(B, S, H * hd)
    return o, lse


def one(B, S, H, KVH, hd, mask):
    import torch
    from dexbotic_b200 import ops
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(20)
    W = (H + 2 * KVH) * hd
    qkv = ((torch.rand((B, S, W), device=dev, generator=g) * 2 - 1) * 0.5 * 3 ** 0.5 * 2).to(torch.bfloat16)
    causal = mask.startswith("causal")
    keymask = bid = None
    if "pad" in mask:
        lens = torch.tensor([S - (3 * i + 5) % max(S // 2, 1) for i in range(B)], device=dev)
        keymask = (torch.arange(S, device=dev)[None, :] < lens[:, None]).to(torch.uint8)
        if mask.startswith("bid"):          # left padding as well: fully masked leading blocks
            keymask[0, : min(70, S // 3)] = 0
    if mask.startswith("bid"):
        # pi0-like: a long bidirectional prefix (0), then 1, then 2s; plus a strictly causal tail on batch 1
        bid = torch.zeros(B, S, dtype=torch.int32, device=dev)
        bid[:, S - S // 8:] = 1
        bid[:, S - S // 16:] = 2
        if B > 1:
            bid[1] = torch.arange(S, device=dev, dtype=torch.int32) // 3
    sh = ops.AttnShape(B, S, H, KVH, hd, torch.bfloat16)
    out, lse = ops.flash_attention_fwd(qkv, sh, keymask=keymask, bid_q=bid, bid_k=bid, causal=causal)
    torch.cuda.synchronize()
    qr = qkv.float().requires_grad_(True)
    ref, ref_lse = ref_attention(qr, B, S, H, KVH, hd, keymask, bid, causal)
    rowmask = keymask.bool()[:, :, None] if keymask is not None else torch.ones(B, S, 1, dtype=torch.bool, device=dev)

    def err(a, b):
        a, b = a.float(), b.float()
        return (a - b).abs().max().item(), ((a - b).norm() / b.norm().clamp_min(1e-20)).item()

    e_out = err(out * rowmask, ref * rowmask)
    lse_nat = lse * 0.6931471805599453
    fin = torch.isfinite(ref_lse) & rowmask[:, None, :, 0]
    e_lse = err(lse_nat[fin], ref_lse[fin])
    dout = ((torch.rand((B, S, H * hd), device=dev, generator=g) * 2 - 1) * 3 ** 0.5).to(torch.bfloat16) * rowmask
    dqkv = torch.full_like(qkv, float("nan"))
    ops.flash_attention_bwd(dout, qkv, out, lse, sh, keymask=keymask, bid_q=bid, bid_k=bid, causal=causal, dqkv=dqkv)
    torch.cuda.synchronize()
    (ref * rowmask).backward(dout.float())
    gr = qr.grad.view(B, S, H + 2 * KVH, hd)
    gg = dqkv.view(B, S, H + 2 * KVH, hd)
    e_dq = err(gg[:, :, :H], gr[:, :, :H])
    e_dk = err(gg[:, :, H:H + KVH], gr[:, :, H:H + KVH])
    e_dv = err(gg[:, :, H + KVH:], gr[:, :, H + KVH:])
    nan = int(torch.isnan(dqkv.float()).sum().item()) + int(torch.isnan(out.float()).sum().item())
    ok = e_out[1] < 2e-2 and e_lse[0] < 2e-2 and e_dq[1] < 3e-2 and e_dk[1] < 3e-2 and e_dv[1] < 3e-2 and nan == 0
    print(f"{'OK  ' if ok else 'FAIL'} B{B} S{S} H{H} KVH{KVH} hd{hd} {mask}: out {e_out[0]:.3e}/{e_out[1]:.3e} "
          f"lse {e_lse[0]:.3e} dq {e_dq[0]:.3e}/{e_dq[1]:.3e} dk {e_dk[0]:.3e}/{e_dk[1]:.3e} "
          f"dv {e_dv[0]:.3e}/{e_dv[1]:.3e} nan {nan}", flush=True)
    if not ok:   # locate: per (batch, head) relative error of the forward, first few rows of the worst one
        o4, r4 = (out * rowmask).float().view(B, S, H, hd), (ref * rowmask).view(B, S, H, hd)
        rel = (o4 - r4).norm(dim=(1, 3)) / r4.norm(dim=(1, 3)).clamp_min(1e-20)
        print("   fwd rel err per (b,h):", [[round(x, 3) for x in row] for row in rel.tolist()][:4])
        per_row = (o4 - r4).norm(dim=3).amax(dim=2)   # [B, S]
        bad = (per_row > 0.05 * r4.norm(dim=3).amax(dim=2).clamp_min(1e-3)).nonzero()[:12].tolist()
        print("   first bad (b, row):", bad)
        for name, a, b_ in (("dq", gg[:, :, :H], gr[:, :, :H]), ("dk", gg[:, :, H:H + KVH], gr[:, :, H:H + KVH]),
                            ("dv", gg[:, :, H + KVH:], gr[:, :, H + KVH:])):
            d = (a.float() - b_).norm(dim=3).amax(dim=2)
            bad = (d > 0.05 * b_.norm(dim=3).amax(dim=2).clamp_min(1e-3)).nonzero()
            print(f"   {name}: {bad.shape[0]} bad (b,row); first {bad[:10].tolist()}")
    return ok


def timing():
    import torch
    from dexbotic_b200 import ops
    dev = "cuda"
    for (B, S, H, KVH, hd, causal, name) in [(32, 309, 28, 4, 128, True, "cogact_7b decoder"),
                                             (32, 257, 16, 16, 64, False, "CLIP-L"),
                                             (24, 256, 16, 16, 72, False, "SigLIP (8x3 views)"),
                                             (8, 867, 8, 1, 256, False, "pi0 joint")]:
        W = (H + 2 * KVH) * hd
        qkv = (torch.randn((B, S, W), device=dev) * 0.5).to(torch.bfloat16)
        keymask = torch.ones(B, S, dtype=torch.uint8, device=dev)
        bid = None
        if name.startswith("pi0"):
            bid = torch.zeros(B, S, dtype=torch.int32, device=dev)
            bid[:, 816:] = 1
            bid[:, 817:] = 2
        sh = ops.AttnShape(B, S, H, KVH, hd, torch.bfloat16)
        dout = torch.randn((B, S, H * hd), device=dev).to(torch.bfloat16)
        dqkv = torch.empty_like(qkv)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        def run_f():
            return ops.flash_attention_fwd(qkv, sh, keymask=keymask, bid_q=bid, bid_k=bid, causal=causal)

        out, lse = run_f()

        def run_b():
            ops.flash_attention_bwd(dout, qkv, out, lse, sh, keymask=keymask, bid_q=bid, bid_k=bid, causal=causal,
                                    dqkv=dqkv)

        for fn, label in ((run_f, "fwd"), (run_b, "bwd")):
            for _ in range(3):
                fn()
            ts = []
            for _ in range(10):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            ts.sort()
            kv = S * (S + 1) / 2 if causal else S * S
            flops = 4 * B * H * kv * hd * (1 if label == "fwd" else 2.5)
            print(f"time {name} {label}: median {ts[len(ts) // 2]:.1f} us (min {ts[0]:.1f}), "
                  f"{flops / ts[len(ts) // 2] / 1e6:.1f} TFLOP/s algorithmic", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        a = sys.argv[2:]
        sys.exit(0 if one(int(a[0]), int(a[1]), int(a[2]), int(a[3]), int(a[4]), a[5]) else 1)
    if len(sys.argv) > 1 and sys.argv[1] == "--timeonly":
        timing()
        sys.exit(0)
    fails = 0
    for c in CASES:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, "--one"] + [str(x) for x in c], timeout=180,
                               capture_output=True, text=True)
            sys.stdout.write(r.stdout)
            if r.returncode != 0:
                fails += 1
                tail = r.stderr.strip().splitlines()[-6:]
                if tail and "FAIL" not in r.stdout:
                    print(f"CRASH {c}: " + " | ".join(tail))
        except subprocess.TimeoutExpired:
            fails += 1
            print(f"TIMEOUT {c} after {time.time() - t0:.0f}s")
        sys.stdout.flush()
    print(f"flash cases failed: {fails} / {len(CASES)}")
    if "--time" in sys.argv and fails == 0:
        timing()
