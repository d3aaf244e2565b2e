"""One forward + one backward flash-attention call per production geometry (for `ncu -k regex:flash_`)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dexbotic_b200 import ops  # noqa: E402

which = sys.argv[1:] or ["cogact", "pi0"]
GEO = {"cogact": (32, 309, 28, 4, 128, True), "pi0": (8, 867, 8, 1, 256, False), "clip": (32, 257, 16, 16, 64, False)}
for name in which:
    B, S, H, KVH, hd, causal = GEO[name]
    W = (H + 2 * KVH) * hd
    qkv = (torch.randn((B, S, W), device="cuda") * 0.5).to(torch.bfloat16)
    keymask = torch.ones(B, S, dtype=torch.uint8, device="cuda")
    bid = None
    if name == "pi0":
        bid = torch.zeros(B, S, dtype=torch.int32, device="cuda")
        bid[:, 816:] = 1
        bid[:, 817:] = 2
    sh = ops.AttnShape(B, S, H, KVH, hd, torch.bfloat16)
    dout = torch.randn((B, S, H * hd), device="cuda").to(torch.bfloat16)
    out, lse = ops.flash_attention_fwd(qkv, sh, keymask=keymask, bid_q=bid, bid_k=bid, causal=causal)
    ops.flash_attention_bwd(dout, qkv, out, lse, sh, keymask=keymask, bid_q=bid, bid_k=bid, causal=causal)
    torch.cuda.synchronize()
print("done")
