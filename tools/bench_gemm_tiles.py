"""256x256 vs 256x128 CTA-pair tiles on the shapes whose 256-wide tiling leaves a partial last wave (N = 3584)."""
import os
import subprocess
import sys

code = r'''
import sys, torch
sys.path.insert(0, ".")
from dexbotic_b200 import ops
from tools.bench_gemm import timeit
M = 9856
for name, N, K in (("o fwd", 3584, 3584), ("down fwd", 3584, 18944), ("qkv fwd", 4608, 3584), ("gate fwd", 18944, 3584)):
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * 0.02
    dy = torch.randn(M, N, device="cuda", dtype=torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    dx = torch.empty(M, K, device="cuda", dtype=torch.bfloat16)
    fl = 2.0 * M * N * K
    for bn in (256, 128):
        t = timeit(lambda: ops.gemm(a, w, out=out, block_n=bn), reps=9)
        t2 = timeit(lambda: ops.gemm(dy, w, b_mn=True, out=dx, block_n=bn), reps=9)
        print(f"{name:9s} bn={bn}: fwd {t*1e3:7.1f} us {fl/t/1e9:7.1f} TF/s | its dgrad (N={K}, K={N}) {t2*1e3:7.1f} us {fl/t2/1e9:7.1f} TF/s", flush=True)
'''
r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, B200_GEMM_PAIR128="1"), capture_output=True, text=True,
                   timeout=400)
print(r.stdout, r.stderr[-2000:])
