"""One launch of every GEMM mode of a CogACT-7B decoder block (for `ncu --set full -k regex:gemm_tcgen05`):
o_proj fwd / dgrad / wgrad, the fused SwiGLU forward (gate | up side by side, act(g)*u epilogue + pre-activations) and the
down-projection dgrad with the GLU backward in its epilogue.  Launch order is printed so the ncu rows can be labelled."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dexbotic_b200 import ops  # noqa: E402

M, d, I = 9856, 3584, 18944
bf = torch.bfloat16
a = torch.randn(M, d, device="cuda", dtype=bf)
w = torch.randn(d, d, device="cuda", dtype=bf) * 0.02
dy = torch.randn(M, d, device="cuda", dtype=bf)
out = torch.empty(M, d, device="cuda", dtype=bf)
dx = torch.empty(M, d, device="cuda", dtype=bf)
dw = torch.empty(d, d, device="cuda", dtype=bf)
wg = torch.randn(I, d, device="cuda", dtype=bf) * 0.02
wu = torch.randn(I, d, device="cuda", dtype=bf) * 0.02
wd = torch.randn(d, I, device="cuda", dtype=bf) * 0.02
h = torch.empty(M, I, device="cuda", dtype=bf)
g = torch.empty(M, I, device="cuda", dtype=bf)
u = torch.empty(M, I, device="cuda", dtype=bf)
dg, du = torch.empty_like(g), torch.empty_like(u)
dwd = torch.empty(d, I, device="cuda", dtype=bf)
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
calls = [
    ("o_proj fwd   M=9856 N=3584 K=3584", lambda: ops.gemm(a, w, out=out)),
    ("o_proj dgrad M=9856 N=3584 K=3584", lambda: ops.gemm(dy, w, b_mn=True, out=dx)),
    ("o_proj wgrad M=3584 N=3584 K=9856", lambda: ops.gemm(dy, a, a_mn=True, b_mn=True, out=dw)),
    ("gate|up fused SwiGLU fwd M=9856 N=2x18944 K=3584 (+aux)", lambda: ops.gemm_dual(a, wg, wu, "silu", out=h, aux_gate=g, aux_up=u)),
    ("down dgrad + GLU bwd epilogue M=9856 N=18944 K=3584", lambda: ops.gemm_glu_bwd(dy, wd, g, u, "silu", dg=dg, du=du)),
    ("down wgrad M=3584 N=18944 K=9856", lambda: ops.gemm(dy, h, a_mn=True, b_mn=True, out=dwd)),
]
for _ in range(2):          # pass 0 warms up (ncu: -s 6 skips it), pass 1 is captured
    for name, fn in calls:
        flush.zero_()
        fn()
torch.cuda.synchronize()
for i, (name, _) in enumerate(calls):
    print(i, name)
