"""Launch one fwd / dgrad / wgrad GEMM of the o_proj shape for an ncu capture."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dexbotic_b200 import ops  # noqa: E402

M, N, K = 9856, 3584, 3584
a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * 0.02
dy = torch.randn(M, N, device="cuda", dtype=torch.bfloat16)
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
dx = torch.empty(M, K, device="cuda", dtype=torch.bfloat16)
dw = torch.empty(N, K, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    ops.gemm(a, w, out=out)
    ops.gemm(dy, w, b_mn=True, out=dx)
    ops.gemm(dy, a, a_mn=True, b_mn=True, out=dw)
torch.cuda.synchronize()
