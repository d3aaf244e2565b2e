"""GEMM pipeline experiments: B200_GEMM_2CTA x B200_GEMM_DEBUG (1 = no TMA loads, 2 = no MMAs), K sweep."""
import os
import subprocess
import sys

code = r'''
import sys, torch, subprocess
sys.path.insert(0, ".")
from dexbotic_b200 import ops
from tools.bench_gemm import timeit
M, N = 9856, 3584
res = []
for K in (1792, 3584, 7168, 14336):
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    t = timeit(lambda: ops.gemm(a, w, out=out), reps=7)
    res.append((K, t * 1e3))
clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
slope = (res[-1][1] - res[0][1]) / ((res[-1][0] - res[0][0]) / 64)      # us per k-block-row of tiles (8 tiles/SM)
icpt = res[0][1] - slope * res[0][0] / 64
print(" ".join(f"K={k}:{t:.0f}us" for k, t in res), f"| per-kblock {slope/8*1000:.0f} ns, per-tile overhead {icpt/8:.2f} us, idle clk {clk}")
'''
for two in ("0", "1"):
    for dbg in ("0", "1"):
        env = dict(os.environ, B200_GEMM_2CTA=two, B200_GEMM_DEBUG=dbg)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=200)
        print(f"2cta={two} debug={dbg} (1=noload +4=mma x2 +8=no fence): {r.stdout.strip()} {r.stderr.strip()[-200:]}", flush=True)
