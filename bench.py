#!/usr/bin/env python
"""bench.py — VLA training samples/sec (BASELINE.json metric) for the CogACT hot path on B200.

  python bench.py --gpus N --steps K --warmup W            our arm (one process per GPU under torchrun for N>1)
  python bench.py --impl reference ...                     the reference's own CPU implementation (oracle port)

One "step" = one full training step over one synthetic batch: ViT-L/14 -> mlp2x_gelu projector -> image-token
splice -> Qwen2.5-7B-shaped decoder -> DiT action head -> loss, backward, global-norm clip, AdamW.
Prints ONE JSON line (see the task contract).  Nothing here reads /root/reference.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "VLA training samples/sec (img+text+action-chunk)"

# public model-card dimensions: Qwen2.5-7B, CLIP ViT-L/14 @224, DiT-S
WORKLOADS = {
    "cogact_7b": dict(
        llm=dict(model_type="qwen2", vocab_size=152064, hidden_size=3584, intermediate_size=18944,
                 num_hidden_layers=28, num_attention_heads=28, num_key_value_heads=4, rms_norm_eps=1e-6,
                 rope_theta=1e6, hidden_act="silu"),
        vision=dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16,
                    image_size=224, patch_size=14, hidden_act="quick_gelu", layer_norm_eps=1e-5),
        action_model_type="DiT-S", action_dim=7, chunk_size=16, batch=32, instr_tokens=32, template_tokens=20),
    # BASELINE.json configs[2] (pi0: SigLIP-So400m + Gemma-2B + flow-matching action expert), per-GPU share of the
    # global batch 64 on 8 GPUs; ONE training forward per sample ("50" is the action horizon)
    "pi0_2b": dict(
        kind="pi0",
        llm=dict(model_type="gemma", vocab_size=257152, hidden_size=2048, intermediate_size=16384, num_hidden_layers=18,
                 num_attention_heads=8, num_key_value_heads=1, head_dim=256, rms_norm_eps=1e-6, rope_theta=10000.0,
                 hidden_act="gelu_pytorch_tanh"),
        expert=dict(model_type="gemma", vocab_size=257152, hidden_size=1024, intermediate_size=4096,
                    num_hidden_layers=18, num_attention_heads=8, num_key_value_heads=1, head_dim=256, rms_norm_eps=1e-6,
                    rope_theta=10000.0, hidden_act="gelu_pytorch_tanh"),
        vision=dict(model_type="siglip_vision_model", hidden_size=1152, intermediate_size=4304, num_hidden_layers=27,
                    num_attention_heads=16, image_size=224, patch_size=14, hidden_act="gelu_pytorch_tanh",
                    layer_norm_eps=1e-6),
        action_dim=32, chunk_size=50, batch=8, n_cam=3, text_tokens=48),
    # SURVEY §8d cfg-5 (single-view layout the reference supports): CogACT trunk + memory bank + DiT-L with per_attn,
    # two 16-frame episode groups per rank batch
    "memvla_7b": dict(
        kind="memvla",
        llm=dict(model_type="qwen2", vocab_size=152064, hidden_size=3584, intermediate_size=18944,
                 num_hidden_layers=28, num_attention_heads=28, num_key_value_heads=4, rms_norm_eps=1e-6,
                 rope_theta=1e6, hidden_act="silu"),
        vision=dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16,
                    image_size=224, patch_size=14, hidden_act="quick_gelu", layer_norm_eps=1e-5),
        action_model_type="DiT-L", action_dim=7, chunk_size=16, batch=32, instr_tokens=32, template_tokens=20,
        mem=dict(dataloader_type="group", group_size=16, per_token_size=256, mem_length=16, retrieval_layers=2,
                 use_timestep_pe=True, fusion_type="gate", consolidate_type="tome", update_fused=True)),
    # SURVEY §8d cfg-4 with the L1-regression head (OpenVLA-OFT style): chunk 8 x dim 7 = 56 action-query rows
    "oft_l1_7b": dict(
        kind="oft_l1",
        llm=dict(model_type="qwen2", vocab_size=152064, hidden_size=3584, intermediate_size=18944,
                 num_hidden_layers=28, num_attention_heads=28, num_key_value_heads=4, rms_norm_eps=1e-6,
                 rope_theta=1e6, hidden_act="silu"),
        vision=dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16,
                    image_size=224, patch_size=14, hidden_act="quick_gelu", layer_norm_eps=1e-5),
        action_dim=7, chunk_size=8, batch=32, instr_tokens=32, template_tokens=20, extra_tokens=56),
    # BASELINE.json configs[3]: OFT with the discrete action tokenizer head, chunk 8 x dim 7 = 56 action tokens
    # (Qwen2.5-7B-shaped decoder as the reference instantiates it; per-GPU share of the global batch 128 on 8 GPUs)
    "oft_discrete_7b": dict(
        kind="oft_discrete",
        llm=dict(model_type="qwen2", vocab_size=152064, hidden_size=3584, intermediate_size=18944,
                 num_hidden_layers=28, num_attention_heads=28, num_key_value_heads=4, rms_norm_eps=1e-6,
                 rope_theta=1e6, hidden_act="silu"),
        vision=dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16,
                    image_size=224, patch_size=14, hidden_act="quick_gelu", layer_norm_eps=1e-5),
        action_dim=7, chunk_size=8, batch=16, instr_tokens=32, template_tokens=20, extra_tokens=56, num_bins=256),
    # (BASELINE.json configs[4] words MemVLA as "multi-view (3 cams)": the reference's MemVLA forward is single-view only —
    # with 5-D images its memory bank indexes episode_ids[i] for i >= B, memvla_arch.py:351-353 — so the measurable
    # configuration is `memvla_7b` above, the single-view layout the reference supports)
    # small stand-in with the same structure for smoke tests / CPU-only debugging of the harness
    "cogact_tiny": dict(
        llm=dict(model_type="qwen2", vocab_size=1024, hidden_size=256, intermediate_size=704, num_hidden_layers=2,
                 num_attention_heads=4, num_key_value_heads=2, rms_norm_eps=1e-6, rope_theta=1e6, hidden_act="silu"),
        vision=dict(hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=2, image_size=56,
                    patch_size=14, hidden_act="quick_gelu", layer_norm_eps=1e-5),
        action_model_type="DiT-S", action_dim=7, chunk_size=16, batch=4, instr_tokens=8, template_tokens=6),
}


def _gemma_flops(c: dict, n_tok: int, S_attn: int) -> float:
    d, I, nl, H, KV, hd = (c["hidden_size"], c["intermediate_size"], c["num_hidden_layers"], c["num_attention_heads"],
                           c["num_key_value_heads"], c["head_dim"])
    return nl * (2 * n_tok * (d * H * hd + 2 * d * KV * hd + H * hd * d + 3 * d * I) + 4 * n_tok * S_attn * H * hd)


def pi0_flops_per_sample(w: dict) -> float:
    V = w["vision"]
    dv, mv, lv = V["hidden_size"], V["intermediate_size"], V["num_hidden_layers"]
    P = (V["image_size"] // V["patch_size"]) ** 2
    vit = w["n_cam"] * lv * (2 * P * (4 * dv * dv + 2 * dv * mv) + 4 * P * P * dv)
    Sp, Ss = w["n_cam"] * P + w["text_tokens"], w["chunk_size"] + 1
    proj = 2 * w["n_cam"] * P * dv * w["llm"]["hidden_size"]
    return 3.0 * (vit + proj + _gemma_flops(w["llm"], Sp, Sp + Ss) + _gemma_flops(w["expert"], Ss, Sp + Ss))


def train_flops_per_sample(w: dict, S: int) -> float:
    """Algorithmic FLOPs (2*MACs; training = 3x forward; recompute NOT counted) — SURVEY.md §8(d) formulae."""
    if w.get("kind") == "pi0":
        return pi0_flops_per_sample(w)
    L = w["llm"]
    d, I, nl = L["hidden_size"], L["intermediate_size"], L["num_hidden_layers"]
    H, KV = L["num_attention_heads"], L["num_key_value_heads"]
    hd = d // H
    dec = nl * (2 * S * (d * H * hd + 2 * d * KV * hd + H * hd * d + 3 * d * I) + 4 * S * S * H * hd)
    V = w["vision"]
    dv, mv, lv = V["hidden_size"], V["intermediate_size"], V["num_hidden_layers"] - 1
    T = (V["image_size"] // V["patch_size"]) ** 2 + 1
    nv = w.get("n_views", 1)
    vit = nv * lv * (2 * T * (4 * dv * dv + 2 * dv * mv) + 4 * T * T * dv)
    proj = nv * 2 * (T - 1) * (dv * d + d * d)
    if w.get("kind") == "oft_discrete":    # lm_head on the 56 action rows only (oft_discrete_arch.py:166-168)
        return 3.0 * (dec + vit + proj + 2 * w["extra_tokens"] * d * L["vocab_size"])
    if w.get("kind") == "oft_l1":      # MLPResNet head: fc1 [A*d -> d] + 2 residual blocks + fc2, per chunk row
        A, Tc = w["action_dim"], w["chunk_size"]
        head = Tc * 2 * (A * d * d + 2 * d * d + d * A)
        return 3.0 * (dec + vit + proj + head)
    depth, wd, _ = {"DiT-S": (6, 384, 4), "DiT-B": (12, 768, 12), "DiT-L": (24, 1024, 16)}[w["action_model_type"]]
    dit = 4 * depth * (2 * 17 * 12 * wd * wd + 4 * 17 * 17 * wd)
    return 3.0 * (dec + vit + proj + dit)


def make_batch(w: dict, rank: int, pinned: bool):
    """Seeded synthetic batch on the HOST: [bos, <image>, instruction, template] ids,
    10 % of rows right-padded by 1-8 tokens, images ~N(0,1), actions ~U(-1,1)."""
    import torch
    g = torch.Generator().manual_seed(1234 + rank)
    B = w["batch"]
    if w.get("kind") == "pi0":
        L, V, img = w["text_tokens"], w["llm"]["vocab_size"], w["vision"]["image_size"]
        ids = torch.randint(1, min(30000, V), (B, L), generator=g)
        n_real = torch.randint(8, 41, (B,), generator=g)
        mask = torch.arange(L)[None, :] < n_real[:, None]
        ids = ids * mask
        batch = dict(input_ids=ids, attention_mask=mask, images=torch.randn(B, w["n_cam"], 3, img, img, generator=g),
                     image_masks=torch.ones(B, w["n_cam"], dtype=torch.bool),
                     actions=torch.randn(B, w["chunk_size"], w["action_dim"], generator=g),
                     states=torch.randn(B, w["action_dim"], generator=g))
        return {k: v.pin_memory() for k, v in batch.items()} if pinned else batch
    L = 2 + w["instr_tokens"] + w["template_tokens"] + (w["extra_tokens"] if w.get("kind") == "oft_discrete" else 0)
    V = w["llm"]["vocab_size"]
    ids = torch.randint(1000 if V > 40000 else 1, min(30000, V), (B, L), generator=g)
    ids[:, 0] = 1
    ids[:, 1] = -200
    mask = torch.ones(B, L, dtype=torch.long)
    for b in range(B):
        if torch.rand((), generator=g).item() < 0.10:
            n = int(torch.randint(1, 9, (), generator=g).item())
            mask[b, L - n:] = 0
    img = w["vision"]["image_size"]
    nv = w.get("n_views", 1)
    images = torch.randn(B, 3, img, img, generator=g) if nv == 1 else torch.randn(B, nv, 3, img, img, generator=g)
    actions = torch.rand(B, w["chunk_size"] * w["action_dim"], generator=g) * 2 - 1
    batch = dict(input_ids=ids, attention_mask=mask, images=images, actions=actions)
    if w.get("kind") == "oft_discrete":
        # the A action-label tokens sit right before the last valid token of every row (oft_discrete_arch.py:66-106);
        # labels there are action token ids U{V-255 .. V-1}, everything else is ignored
        A = w["extra_tokens"]
        labels = torch.full((B, L), -100, dtype=torch.long)
        npl = mask.sum(1)
        for b in range(B):
            lo = int(npl[b]) - A - 1
            tok = torch.randint(V - w["num_bins"] + 1, V, (A,), generator=g)
            ids[b, lo:lo + A] = tok
            labels[b, lo:lo + A] = tok
        batch["labels"] = labels
    if pinned:
        batch = {k: v.pin_memory() for k, v in batch.items()}
    if w.get("kind") == "memvla":      # (dataset, episode, frame): consecutive frames of B/group episodes
        G = w["mem"]["group_size"]
        batch["indexes"] = [(0, 100 * rank + b // G, 10 + b % G) for b in range(B)]
    return batch


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


class GemmTimer:
    """CUDA events around every tcgen05 GEMM launch on the launching stream (roofline.achieved)."""

    def __init__(self):
        self.records = []
        self.enabled = False

    def install(self):
        import torch
        from dexbotic_b200 import _lib, ops
        lib = _lib.load()
        orig = lib.b200_gemm
        timer = self

        def wrapped(args_ref, stream):
            if not timer.enabled:
                return orig(args_ref, stream)
            a = args_ref._obj
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            rc = orig(args_ref, stream)
            e.record()
            z = max(a.z_lo, 1) * max(a.z_hi, 1)
            fl = 2.0 * a.m * a.n * a.k * max(a.k_segs, 1) * z * (2 if a.dual_b else 1)
            timer.records.append((s, e, fl))
            return rc

        class _Proxy:
            def __getattr__(self, name):
                return wrapped if name == "b200_gemm" else getattr(lib, name)

        _lib._lib = _Proxy()

    def summary(self, large_flops: float = 5e9):
        """(seconds, flops, launches) over all timed GEMM launches, and the same over the launches with at least
        `large_flops` each (decoder / tower shapes: event-pair overhead of a few us is negligible there)."""
        ts = [(s.elapsed_time(e) * 1e-3, fl) for s, e, fl in self.records]
        big = [(t, fl) for t, fl in ts if fl >= large_flops]
        return (sum(t for t, _ in ts), sum(fl for _, fl in ts), len(ts),
                sum(t for t, _ in big), sum(fl for _, fl in big), len(big))


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return json.loads(p.read_text()), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def run_ours(args) -> dict:
    import torch
    import torch.distributed as dist
    from dexbotic_b200 import _lib
    from dexbotic_b200.model import CogActConfig, CogACTForCausalLM
    from dexbotic_b200.parallel import GradientOverlap

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    w = WORKLOADS[args.workload]
    want_ce = (world > 1 and os.environ.get("B200_DP", "zero1") == "zero1"
               and os.environ.get("B200_DP_TRANSPORT", "ce") == "ce")
    from dexbotic_b200.params import ParamStore
    ParamStore.SYMMETRIC = want_ce           # gradient / weight buffers in symmetric memory: peer-memory exchange
    try:
        model = build_model(w, dev)
    except Exception as e:                   # symmetric allocation unavailable on this box: NCCL transport instead
        if not want_ce:
            raise
        print(f"[bench] symmetric memory unavailable ({type(e).__name__}: {e}); using the NCCL transport", file=sys.stderr)
        ParamStore.SYMMETRIC = False
        model = build_model(w, dev)
    model.init_weights_(seed=1234)          # same seeded random init on every rank (no checkpoints offline)
    model.train()
    host = make_batch(w, rank, pinned=True)
    B = w["batch"]
    return _run_timed(args, w, model, host, B, rank, world, local, dev)


def build_model(w: dict, dev):
    from dexbotic_b200.model import CogActConfig, CogACTForCausalLM
    if w.get("kind") == "pi0":
        from dexbotic_b200.model import Pi0Config, Pi0ForCausalLM
        cfg = Pi0Config(llm_config=w["llm"], action_config=w["expert"], vision_config=w["vision"],
                        action_dim=w["action_dim"], chunk_size=w["chunk_size"])
        model = Pi0ForCausalLM(cfg, device=dev)
    elif w.get("kind") == "memvla":
        from dexbotic_b200.model import MemVLAConfig, MemVLAForCausalLM
        cfg = MemVLAConfig(llm_config=w["llm"], mm_vision_tower=w["vision"], mm_projector_type="mlp2x_gelu",
                           action_model_type=w["action_model_type"], action_dim=w["action_dim"],
                           chunk_size=w["chunk_size"], **w["mem"])
        model = MemVLAForCausalLM(cfg, device=dev)
    elif w.get("kind") == "oft_discrete":
        from dexbotic_b200.model import OFTDiscreteConfig, OFTDiscreteForCausalLM
        cfg = OFTDiscreteConfig(llm_config=w["llm"], mm_vision_tower=w["vision"], mm_projector_type="mlp2x_gelu",
                                action_model_type="Discrete", action_dim=w["action_dim"], chunk_size=w["chunk_size"],
                                num_bins=w["num_bins"])
        model = OFTDiscreteForCausalLM(cfg, device=dev)
    elif w.get("kind") == "oft_l1":
        from dexbotic_b200.model import OFTConfig, OFTForCausalLM
        cfg = OFTConfig(llm_config=w["llm"], mm_vision_tower=w["vision"], mm_projector_type="mlp2x_gelu",
                        action_model_type="Linear", action_dim=w["action_dim"], chunk_size=w["chunk_size"])
        model = OFTForCausalLM(cfg, device=dev)
    else:
        cfg = CogActConfig(llm_config=w["llm"], mm_vision_tower=w["vision"], mm_projector_type="mlp2x_gelu",
                           action_model_type=w["action_model_type"], action_dim=w["action_dim"],
                           chunk_size=w["chunk_size"])
        model = CogACTForCausalLM(cfg, device=dev)
    return model


def _run_timed(args, w, model, host, B, rank, world, local, dev):
    import torch
    import torch.distributed as dist
    from dexbotic_b200 import _lib
    from dexbotic_b200.parallel import GradientOverlap

    def to_dev(hb):
        return {k: (v.to(dev, non_blocking=True) if hasattr(v, "to") else v) for k, v in hb.items()}

    # data-parallel exchange, in place on the flat buffers and overlapped with backward.  Default: ZeRO-1
    # (reduce-scatter of gradients, AdamW on 1/N of every chunk, all-gather of the bf16 weights under the next forward);
    # B200_DP=allreduce selects the plain gradient all-reduce + replicated AdamW
    dp_mode = os.environ.get("B200_DP", "zero1") if world > 1 else "none"
    if dp_mode == "zero1":
        from dexbotic_b200.parallel import ShardedDataParallel
        try:
            overlap = ShardedDataParallel(model.store)
        except Exception as e:               # peer mapping failed: same sharded optimizer over NCCL
            print(f"[bench] peer-memory transport unavailable ({type(e).__name__}: {e}); NCCL transport", file=sys.stderr)
            overlap = ShardedDataParallel(model.store, transport="nccl")
    else:
        overlap = GradientOverlap(model.store, reserve_sms=int(os.environ.get("B200_DP_RESERVE_SMS", "0")))
    # the HBM-bound per-block AdamW of step t runs on a side stream under the tensor-bound forward of step t+1
    # (ParamStore.adamw_step); the timed region below waits for the LAST step's updates before it closes
    model.store.async_optimizer = True

    def step(batch):
        model.zero_grad()
        out = model(**batch)
        out.loss.backward()
        overlap.finish()
        model.optimizer_step(base_lr=2e-5)
        return out

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n_steps, from_host: bool, gemm_timer=None):
        sync_all()
        launches0 = _lib.launch_count()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if gemm_timer is not None:
            gemm_timer.enabled = True
        # `ncu --profile-from-start off ... B200_PROFILER_RANGE=1 python bench.py`: the launch list covers exactly the
        # device-timed steps (numbers printed by such a run are not bench values)
        ranged = bool(os.environ.get("B200_PROFILER_RANGE")) and not from_host
        if ranged:
            torch.cuda.profiler.start()
        s.record()
        last = None
        for _ in range(n_steps):
            batch = to_dev(host) if from_host else resident
            out = step(batch)
            if from_host:
                last = out.loss.item()             # device->host read of the step's result, every step
        model.store.wait_all_params()              # the last step's overlapped optimizer updates belong to this region
        e.record()
        if ranged:
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
        if gemm_timer is not None:
            gemm_timer.enabled = False
        sync_all()
        ms = s.elapsed_time(e)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, _lib.launch_count() - launches0, (last if from_host else out.loss.item())

    resident = to_dev(host)
    S = None
    for _ in range(max(args.warmup, 3)):
        out = step(resident)
        S = out.logits.shape[1]
    if w.get("kind") == "pi0":
        S = w["n_cam"] * (w["vision"]["image_size"] // w["vision"]["patch_size"]) ** 2 + w["text_tokens"] + w["chunk_size"] + 1
    if w.get("kind") == "oft_discrete":  # logits are [B, 56, V] there; S = text + image + 56 action placeholders
        S = (1 + w["instr_tokens"] + w["template_tokens"] + (w["vision"]["image_size"] // w["vision"]["patch_size"]) ** 2
             + w["extra_tokens"])
    if w.get("kind") == "oft_l1":       # logits are the predicted actions there; S = text + image + action-query rows
        S = (1 + w["instr_tokens"] + w["template_tokens"] + (w["vision"]["image_size"] // w["vision"]["patch_size"]) ** 2
             + w["extra_tokens"])
    gt = GemmTimer()
    gt.install()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, launches, loss_dev = timed(args.steps, from_host=False, gemm_timer=gt)
    ms_e2e, _, loss_e2e = timed(args.steps, from_host=True)
    clocks = sampler.stop() if rank == 0 else {}
    gemm_s, gemm_flops, gemm_n, big_s, big_flops, big_n = gt.summary()

    peaks, peak_src = measured_peaks()
    flops_sample = train_flops_per_sample(w, S)
    value = B * world * args.steps / (ms_dev * 1e-3)
    e2e = B * world * args.steps / (ms_e2e * 1e-3)
    h2d = sum(v.numel() * v.element_size() for v in host.values() if hasattr(v, "numel"))
    peak_tf = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))
    traffic = None
    tp = ROOT / "profiles" / "r2_gemm_traffic.json"
    if tp.exists():       # dram__bytes_read+write of one `ncu --set full` capture of the dominant kernel (committed)
        traffic = json.loads(tp.read_text())
    achieved_tf = gemm_flops / gemm_s / 1e12 if gemm_s > 0 else 0.0
    res = {
        "metric": METRIC, "value": round(value, 3), "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms_dev / args.steps, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": (f"{args.workload}: pi0 SigLIP-So400m x{w['n_cam']} cams + Gemma-2B + 300M action expert, "
                                f"chunk {w['chunk_size']}, batch={B}/GPU, joint S={S}, random-init weights, AdamW + clip; "
                                "no recompute;" if w.get("kind") == "pi0" else
                                f"{args.workload}: CogACT ViT-L/14@224 + Qwen2.5-7B-shaped decoder + "
                                f"{w.get('action_model_type', w.get('kind'))}, batch={B}/GPU, S={S}, random-init weights, AdamW + clip, "
                                f"{model.model_engine.llm.keep_layers}/{len(model.model_engine.llm.blocks)} decoder "
                                "blocks keep activations (rest recompute);") +
                               " inputs << L2 but weights+grads+moments stream through HBM every step "
                               "(>120 GB for the 7B model), so L2 is cold for the timed kernels",
                   "global_batch": B * world, "seq_len": S,
                   "parallelism": f"dp{world}" + ("" if world == 1 else
                                                  (" (ZeRO-1 over NVLink peer memory: one kernel per gradient chunk loads the "
                                                   "rank's piece from every peer and averages in fp32 [reduce-scatter], "
                                                   "sharded AdamW, copy-engine push-all-gather of bf16 weights; no NCCL "
                                                   "kernel on the data path)"
                                                   if getattr(overlap, "ce", False) else
                                                   " (ZeRO-1: NCCL reduce-scatter grads, sharded AdamW, all-gather bf16 weights)")
                                                  if dp_mode == "zero1" else " (gradient all-reduce)"),
                   "train_tflop_per_sample": round(flops_sample / 1e12, 3)},
        "e2e": {"value": round(e2e, 3), "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": round(ms_e2e / args.steps, 3)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": round(achieved_tf, 1), "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": round(achieved_tf / peak_tf, 4) if peak_tf else None, "traffic": traffic,
                     "kernel": "gemm_tcgen05_kernel", "launches_timed": gemm_n,
                     # the same measurement restricted to launches of >= 5 GFLOP (decoder / tower GEMMs): the few-us
                     # CUDA-event overhead per pair no longer dilutes it (DiT / memory-bank GEMMs take ~5 us each)
                     "achieved_large_launches": round(big_flops / big_s / 1e12, 1) if big_s > 0 else None,
                     "large_launches": big_n,
                     "large_launch_flop_share": round(big_flops / gemm_flops, 4) if gemm_flops else None,
                     "peak_source": f"{peak_src} (bf16_tflops_sustained)",
                     "gemm_share_of_step": round(gemm_s / (ms_dev * 1e-3), 4),
                     "step_mfu_algorithmic": round(value / world * flops_sample / 1e12 / peak_tf, 4) if peak_tf else None},
        "loss": round(float(loss_dev), 5),
        "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2),
    }
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        res["cpu_baseline"] = cpu_baseline(w, S, seconds_budget=20.0, workload=args.workload, limit_s=90.0,
                                            prefer_reference=w.get("kind") is None)
    if world > 1:
        dist.destroy_process_group()
    return res if rank == 0 else {}


def cpu_baseline(w: dict, S: int, seconds_budget: float = 20.0, steps: int = 1, workload: str = "cogact_7b",
                 limit_s: float = 120.0, prefer_reference: bool = False) -> dict:
    """The reference's CPU implementation of the step, timed on this box's host cores on a bounded sample, in a child
    process under a wall-clock limit (a slow host can never stall the bench).  `prefer_reference`: first the UNMODIFIED
    reference classes from the vendored baseline/_ref (`kind: "reference"`, oracle/cpu_baseline.py:time_reference_sample);
    else / on failure the oracle port (`kind: "port"`) at batch 4, then batch 1; if nothing finishes the line says so."""
    last = "not run"
    plan = ([("reference", 4, 2.25 * limit_s)] if prefer_reference else []) + [("port", 4, limit_s), ("port", 1, limit_s / 2)]
    if prefer_reference:
        plan = plan[:2]
    for kind, batch, lim in plan:
        try:
            env = {k: v for k, v in os.environ.items() if k not in ("OMP_NUM_THREADS", "MKL_NUM_THREADS")}
            env["CUDA_VISIBLE_DEVICES"] = ""
            r = subprocess.run([sys.executable, "-m", "oracle.cpu_baseline", workload, str(S),
                                str(1 if kind == "reference" else steps), str(batch), kind],
                               cwd=str(ROOT), env=env, capture_output=True, text=True, timeout=lim)
            lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            if r.returncode == 0 and lines:
                return json.loads(lines[-1])
            last = f"{kind} batch {batch}: exit {r.returncode}: {r.stderr.strip()[-200:]}"
        except subprocess.TimeoutExpired:
            last = f"{kind} batch {batch} did not finish in {lim:.0f} s"
    return {"value": None, "unit": "samples/s", "cores": None, "kind": "port", "sample": f"unavailable ({last})"}


def run_reference(args) -> dict:
    """--impl reference: the reference's own CPU implementation of the path on the host cores, bounded sample per step:
    the UNMODIFIED reference classes from the vendored baseline/_ref when they load (270 s limit), else the oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return {}
    w = WORKLOADS[args.workload]
    S = 2 + w["instr_tokens"] + w["template_tokens"] - 1 + (w["vision"]["image_size"] // w["vision"]["patch_size"]) ** 2
    cb = cpu_baseline(w, S, seconds_budget=30.0, steps=max(1, min(args.steps, 2)), workload=args.workload, limit_s=120.0,
                      prefer_reference=True)
    return {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 / cb["value"], 1) if cb["value"] else None,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload} (reference CPU arm, kind={cb.get('kind')}, bounded sample: {cb['sample']})"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def run_reference_gpu(args) -> dict:
    """--impl reference_gpu: the number the north_star is written against — the UNMODIFIED reference
    model classes (vendored under git-ignored baseline/_ref/ by tools/install_reference.sh, imported through the compat
    loader oracle/ref_loader.py because this image has transformers 5.5) training on the same B200 with the reference's
    own settings: fp32 parameters + bf16 autocast (HF Trainer bf16=True, base_exp.py:253), tf32 allowed (:254), HF's
    default SDPA attention + cuBLAS, gradient_checkpointing=True non-reentrant (base_exp.py:245, trainer.py:120),
    torch.optim.AdamW, max_grad_norm=1.0 (trainer.py:122), plain DDP for N>1 (deepspeed=None; DeepSpeed is not
    installable offline).  Same synthetic batches and shapes as our arm; nothing of dexbotic_b200 is on this path."""
    import torch
    import torch.distributed as dist
    from oracle import ref_loader

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    w = WORKLOADS[args.workload]
    torch.manual_seed(1234)
    with torch.device(dev):
        if w.get("kind") == "pi0":
            model = ref_loader.build_reference_pi0(dict(w["llm"]), dict(w["expert"]), dict(w["vision"]),
                                                   action_dim=w["action_dim"], chunk_size=w["chunk_size"])
        elif w.get("kind") is None:
            from transformers import CLIPVisionConfig, Qwen2Config
            llm = {k: v for k, v in w["llm"].items() if k != "model_type"}
            model = ref_loader.build_reference_cogact(Qwen2Config(max_position_embeddings=4096, **llm),
                                                      CLIPVisionConfig(**w["vision"]), w["action_model_type"],
                                                      action_dim=w["action_dim"], chunk_size=w["chunk_size"])
        else:
            raise SystemExit(f"reference_gpu: workload {args.workload} not wired")
    model.to(dev)
    for p_ in model.model.parameters():          # base_exp.py:318-321: everything under model.model trains
        p_.requires_grad = True
    model.train()
    model.gradient_checkpointing_enable(gradient_checkpointing_kwargs={"use_reentrant": False})
    params = [p_ for p_ in model.parameters() if p_.requires_grad]
    opt = torch.optim.AdamW(params, lr=2e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, fused=True)
    net = model
    if world > 1:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=True)
    B = args.ref_batch or w["batch"]
    accum = max(1, w["batch"] // B)
    wb = dict(w, batch=B)
    host = [make_batch(wb, rank * accum + i, pinned=True) for i in range(accum)]

    def to_dev(hb):
        out = {}
        for k, v in hb.items():
            if not hasattr(v, "to"):
                continue
            v = v.to(dev, non_blocking=True)
            out[k] = v.to(torch.bfloat16) if (v.is_floating_point() and k == "images") else v
        return out

    def step(batches):
        opt.zero_grad(set_to_none=True)
        loss = None
        for bt in batches:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = net(**bt)
            (out.loss / len(batches)).backward()
            loss = out.loss
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        return loss

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    resident = [to_dev(h) for h in host]
    for _ in range(max(args.warmup, 3)):
        step(resident)

    def timed(n, from_host):
        sync_all()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        last = None
        for _ in range(n):
            loss = step([to_dev(h) for h in host] if from_host else resident)
            if from_host:
                last = loss.item()
        e.record()
        sync_all()
        ms = s.elapsed_time(e)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, (last if from_host else loss.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, loss_dev = timed(args.steps, False)
    ms_e2e, _ = timed(args.steps, True)
    clocks = sampler.stop() if rank == 0 else {}
    per_step = w["batch"] * world
    h2d = sum(v.numel() * v.element_size() for h in host for v in h.values() if hasattr(v, "numel"))
    import transformers
    res = {"impl": "reference_gpu", "metric": METRIC, "value": round(per_step * args.steps / (ms_dev * 1e-3), 3),
           "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
           "ms_per_step": round(ms_dev / args.steps, 3), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "bf16 autocast over fp32 parameters (HF Trainer bf16=True)", "data": "synthetic",
           "config": {"workload": f"{args.workload}: unmodified reference classes from baseline/_ref via the compat "
                                  f"loader, transformers {transformers.__version__}, torch {torch.__version__}; SDPA + "
                                  "cuBLAS, gradient_checkpointing (non-reentrant), fused torch AdamW, clip 1.0, "
                                  f"micro-batch {B} x {accum} accumulation = {w['batch']}/GPU, "
                                  + ("DDP (find_unused_parameters)" if world > 1 else "single GPU"),
                      "global_batch": per_step, "parallelism": f"dp{world}"},
           "e2e": {"value": round(per_step * args.steps / (ms_e2e * 1e-3), 3), "unit": "samples/s",
                   "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / args.steps, 3)},
           "clocks": clocks, "loss": round(float(loss_dev), 5),
           "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}
    if world > 1:
        dist.destroy_process_group()
    return res if rank == 0 else {}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference_gpu"])
    ap.add_argument("--ref-batch", type=int, default=0,
                    help="reference_gpu: micro-batch per GPU (gradient accumulation up to the workload's batch)")
    ap.add_argument("--workload", default="cogact_7b", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    res = (run_reference(args) if args.impl == "reference" else
           run_reference_gpu(args) if args.impl == "reference_gpu" else run_ours(args))
    if res:
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
