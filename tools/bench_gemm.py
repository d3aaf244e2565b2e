"""Micro-benchmark of the tcgen05 GEMM on the decoder's shapes (CUDA events, L2 flushed between reps)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dexbotic_b200 import ops  # noqa: E402


def timeit(fn, reps=10, warm=3):
    flush = torch.empty(256 * 1024 * 1024, device="cuda", dtype=torch.uint8)
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 9856
    d, inter, qkv = 3584, 18944, 4608
    dev = "cuda"
    x = torch.randn(M, d, device=dev, dtype=torch.bfloat16)
    for name, N, K in [("qkv", qkv, d), ("o", d, d), ("gate", inter, d), ("down", d, inter)]:
        a = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        w = torch.randn(N, K, device=dev, dtype=torch.bfloat16) * 0.02
        dy = torch.randn(M, N, device=dev, dtype=torch.bfloat16)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        dx = torch.empty(M, K, device=dev, dtype=torch.bfloat16)
        dw = torch.empty(N, K, device=dev, dtype=torch.bfloat16)
        fl = 2.0 * M * N * K
        for tag, fn, ref in [
            ("fwd  ", lambda: ops.gemm(a, w, out=out), lambda: torch.matmul(a, w.t(), out=out)),
            ("dgrad", lambda: ops.gemm(dy, w, b_mn=True, out=dx), lambda: torch.matmul(dy, w, out=dx)),
            ("wgrad", lambda: ops.gemm(dy, a, a_mn=True, b_mn=True, out=dw), lambda: torch.matmul(dy.t(), a, out=dw)),
        ]:
            t = timeit(fn)
            tr = timeit(ref)
            print(f"{name:5s} {tag} M={M} N={N} K={K}: b200 {t:7.3f} ms {fl / t / 1e9:7.1f} TF/s | cuBLAS {tr:7.3f} ms "
                  f"{fl / tr / 1e9:7.1f} TF/s", flush=True)
    # fused SwiGLU dual GEMM vs two GEMMs + glu
    a = torch.randn(M, d, device=dev, dtype=torch.bfloat16)
    wg = torch.randn(inter, d, device=dev, dtype=torch.bfloat16) * 0.02
    wu = torch.randn(inter, d, device=dev, dtype=torch.bfloat16) * 0.02
    h = torch.empty(M, inter, device=dev, dtype=torch.bfloat16)
    t = timeit(lambda: ops.gemm_dual(a, wg, wu, "silu", out=h))
    fl = 4.0 * M * inter * d
    print(f"dual swiglu: {t:7.3f} ms {fl / t / 1e9:7.1f} TF/s")
    g = torch.empty(M, inter, device=dev, dtype=torch.bfloat16)
    u = torch.empty(M, inter, device=dev, dtype=torch.bfloat16)
    t = timeit(lambda: ops.gemm_dual(a, wg, wu, "silu", out=h, aux_gate=g, aux_up=u))
    print(f"dual swiglu + aux: {t:7.3f} ms {fl / t / 1e9:7.1f} TF/s")
    # GLU backward: dgrad of the down projection + glu_bwd kernel vs the dgrad with the GLU backward in its epilogue
    dy = torch.randn(M, d, device=dev, dtype=torch.bfloat16)
    wd = torch.randn(d, inter, device=dev, dtype=torch.bfloat16) * 0.02
    dh = torch.empty(M, inter, device=dev, dtype=torch.bfloat16)
    dg, du = torch.empty_like(g), torch.empty_like(u)

    def unfused():
        ops.gemm(dy, wd, b_mn=True, out=dh)
        ops.glu_bwd(dh, g, u, "silu", dg=dg, du=du, h_out=dh)

    t1 = timeit(unfused)
    t2 = timeit(lambda: ops.gemm_glu_bwd(dy, wd, g, u, "silu", dg=dg, du=du))
    fl2 = 2.0 * M * inter * d
    print(f"glu bwd: dgrad + glu_bwd kernel {t1:7.3f} ms | fused epilogue {t2:7.3f} ms {fl2 / t2 / 1e9:7.1f} TF/s")


if __name__ == "__main__":
    main()
