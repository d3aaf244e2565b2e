"""Pipeline timeline of the flash-attention kernels: CTA 0 stamps clock64() at its hand-off points
(b200_flash_attn_set_trace) and this prints, per step, how long each phase took in SM cycles.
Usage on the GPU box:  python tools/trace_flash.py [B S H KVH hd]      (default: the CogACT-7B decoder layer)
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from dexbotic_b200 import _lib, ops  # noqa: E402

FWD = {16: "I:wait_p", 17: "I:p_full", 18: "I:pv_issued", 19: "S:top", 20: "S:s_full", 21: "S:ld_done", 22: "S:max_done",
       23: "S:exp_done", 24: "S:pv_done", 25: "S:arrived", 26: "E:start", 27: "E:pv_done", 28: "E:end"}
DQ = {0: "A:top", 1: "A:sdp_free", 2: "A:kv_full", 3: "A:issued", 4: "B:ds_full", 5: "B:issued", 6: "S:top", 7: "S:sdp_full",
      8: "S:ld_done", 9: "S:computed", 10: "S:ds_empty", 11: "S:arrived", 12: "E:start", 13: "E:dq_full", 14: "E:end"}
DKV = {32: "A:top", 33: "A:sdp_free", 34: "A:ring_full", 35: "A:issued", 36: "B:pds_full", 37: "B:issued", 38: "S:top",
       39: "S:sdp_full", 40: "S:ld_done", 41: "S:computed", 42: "S:pds_empty", 43: "S:arrived"}


def show(name, tr, table, steps):
    slots = sorted(table)
    t0 = min(int(tr[s, 0]) for s in slots if tr[s, 0] > 0)
    print(f"\n== {name}: cycles since the first stamp of CTA 0 (0 = not reached)")
    print("step " + " ".join(f"{table[s]:>12s}" for s in slots))
    for i in range(steps):
        row = [int(tr[s, i]) - t0 if tr[s, i] > 0 else 0 for s in slots]
        if not any(row):
            break
        print(f"{i:4d} " + " ".join(f"{v:12d}" for v in row))


def main():
    a = [int(x) for x in sys.argv[1:6]]
    B, S, H, KVH, hd = a if len(a) == 5 else (32, 309, 28, 4, 128)
    dev = "cuda"
    W = (H + 2 * KVH) * hd
    g = torch.Generator(device=dev).manual_seed(1)
    qkv = (torch.randn((B, S, W), device=dev, generator=g) * 0.5).to(torch.bfloat16)
    keymask = torch.ones(B, S, dtype=torch.uint8, device=dev)
    sh = ops.AttnShape(B, S, H, KVH, hd, torch.bfloat16)
    lib = _lib.load()
    out, lse = ops.flash_attention_fwd(qkv, sh, keymask=keymask, causal=True)       # warm-up (module load, attributes)
    dout = torch.randn_like(out)
    dqkv = torch.empty_like(qkv)
    ops.flash_attention_bwd(dout, qkv, out, lse, sh, keymask=keymask, causal=True, dqkv=dqkv)
    torch.cuda.synchronize()
    tr = torch.zeros(64, 64, dtype=torch.int64, device=dev)
    lib.b200_flash_attn_set_trace(tr.data_ptr())
    ops.flash_attention_fwd(qkv, sh, keymask=keymask, causal=True)
    torch.cuda.synchronize()
    show("forward (tile 0 of CTA 0; steps = key blocks)", tr.cpu(), FWD, 24)
    tr.zero_()
    ops.flash_attention_bwd(dout, qkv, out, lse, sh, keymask=keymask, causal=True, dqkv=dqkv)
    torch.cuda.synchronize()
    lib.b200_flash_attn_set_trace(None)
    t = tr.cpu()
    show("dQ kernel", t, DQ, 24)
    show("dK/dV kernel", t, DKV, 24)


if __name__ == "__main__":
    main()
