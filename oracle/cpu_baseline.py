"""TEST / MEASUREMENT INFRASTRUCTURE — times the oracle port (oracle/vla_oracle.py) of the reference's CogACT
training step on the host cores.  Used only by bench.py's cpu_baseline leg and `--impl reference`.

Bounded sample (a full 7B fp32 step is ~38 TFLOP per sample and > 120 GB of host memory for weights + gradients +
Adam moments): batch 4 (so that the weight matrices are reused across rows as in the GPU arm's batch, instead of the
weight-streaming-bound batch 1), full-size ViT-L/14 + projector + DiT, and the decoder timed at 1 and 2 full-size layers;
the 28-layer time is the linear extrapolation T(1) + (L-1) * (T(2) - T(1)) — decoder layers are identical, so the cost is
linear in L.  The JSON says `extrapolated: true`.  bench.py runs this module in a child process without torchrun's OMP_NUM_THREADS=1, under a
wall-clock limit, so the thread count is torch's default (physical cores) at every N.
"""
from __future__ import annotations

import os
import time

import torch

from . import vla_oracle
from .weights import seeded_tensor


def _shapes(w: dict, n_dec: int) -> dict:
    L, V = w["llm"], w["vision"]
    d, I, H, KV = L["hidden_size"], L["intermediate_size"], L["num_attention_heads"], L["num_key_value_heads"]
    hd = d // H
    sh = {"model.llm.embed_tokens.weight": (L["vocab_size"], d), "model.llm.norm.weight": (d,)}
    for i in range(n_dec):
        q = f"model.llm.layers.{i}."
        sh.update({q + "input_layernorm.weight": (d,), q + "post_attention_layernorm.weight": (d,),
                   q + "self_attn.q_proj.weight": (H * hd, d), q + "self_attn.q_proj.bias": (H * hd,),
                   q + "self_attn.k_proj.weight": (KV * hd, d), q + "self_attn.k_proj.bias": (KV * hd,),
                   q + "self_attn.v_proj.weight": (KV * hd, d), q + "self_attn.v_proj.bias": (KV * hd,),
                   q + "self_attn.o_proj.weight": (d, H * hd), q + "mlp.gate_proj.weight": (I, d),
                   q + "mlp.up_proj.weight": (I, d), q + "mlp.down_proj.weight": (d, I)})
    dv, mv, lv, ps = V["hidden_size"], V["intermediate_size"], V["num_hidden_layers"], V["patch_size"]
    P = (V["image_size"] // ps) ** 2
    p = "model.mm_vision_tower.vision_tower.vision_model."
    sh.update({p + "embeddings.class_embedding": (dv,), p + "embeddings.patch_embedding.weight": (dv, 3, ps, ps),
               p + "embeddings.position_embedding.weight": (P + 1, dv), p + "pre_layrnorm.weight": (dv,),
               p + "pre_layrnorm.bias": (dv,)})
    for i in range(lv - 1):
        q = f"{p}encoder.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sh[q + f"self_attn.{n}.weight"] = (dv, dv)
            sh[q + f"self_attn.{n}.bias"] = (dv,)
        for n in ("layer_norm1", "layer_norm2"):
            sh[q + n + ".weight"] = (dv,)
            sh[q + n + ".bias"] = (dv,)
        sh.update({q + "mlp.fc1.weight": (mv, dv), q + "mlp.fc1.bias": (mv,), q + "mlp.fc2.weight": (dv, mv),
                   q + "mlp.fc2.bias": (dv,)})
    sh.update({"model.mm_projector.0.weight": (d, dv), "model.mm_projector.0.bias": (d,),
               "model.mm_projector.2.weight": (d, d), "model.mm_projector.2.bias": (d,)})
    depth, wd, _ = {"DiT-S": (6, 384, 4), "DiT-B": (12, 768, 12), "DiT-L": (24, 1024, 16)}[w["action_model_type"]]
    a = "model.action_head.net."
    A = w["action_dim"]
    sh.update({a + "positional_embedding": (w["chunk_size"] + 1, wd), a + "x_embedder.linear.weight": (wd, A),
               a + "x_embedder.linear.bias": (wd,), a + "t_embedder.mlp.0.weight": (wd, 256),
               a + "t_embedder.mlp.0.bias": (wd,), a + "t_embedder.mlp.2.weight": (wd, wd),
               a + "t_embedder.mlp.2.bias": (wd,), a + "z_embedder.uncondition": (1, d),
               a + "z_embedder.linear.weight": (wd, d), a + "z_embedder.linear.bias": (wd,),
               a + "final_layer.linear.weight": (A, wd), a + "final_layer.linear.bias": (A,)})
    for i in range(depth):
        q = f"{a}blocks.{i}."
        sh.update({q + "attn.qkv.weight": (3 * wd, wd), q + "attn.qkv.bias": (3 * wd,), q + "attn.proj.weight": (wd, wd),
                   q + "attn.proj.bias": (wd,), q + "mlp.fc1.weight": (4 * wd, wd), q + "mlp.fc1.bias": (4 * wd,),
                   q + "mlp.fc2.weight": (wd, 4 * wd), q + "mlp.fc2.bias": (wd,)})
    return sh


def _one_step(sd, cfg, batch, opt):
    R = 4
    t0 = time.perf_counter()
    opt.zero_grad(set_to_none=True)
    out = vla_oracle.cogact_forward(sd, cfg, batch["input_ids"], batch["attention_mask"], batch["images"],
                                    batch["actions"], batch["noise"], batch["timesteps"], batch["drop"], R)
    out["loss"].backward()
    torch.nn.utils.clip_grad_norm_([p for p in sd.values() if p.grad is not None], 1.0)
    opt.step()
    return time.perf_counter() - t0


def _time_with_layers(w: dict, n_dec: int, steps: int, B: int = 4) -> float:
    g = torch.Generator().manual_seed(1)
    sd = {}
    for name, shape in _shapes(w, n_dec).items():
        if name.endswith("embed_tokens.weight"):      # 545 M values: cheap uniform init instead of a seeded normal
            sd[name] = (torch.rand(shape) - 0.5) * 0.04
        else:
            sd[name] = seeded_tensor(name, shape, 1)
    sd = {k: v.requires_grad_(True) for k, v in sd.items()}
    L = dict(w["llm"])
    L["num_hidden_layers"] = n_dec
    cfg = dict(llm=L, vision=w["vision"], action_dim=w["action_dim"], chunk_size=w["chunk_size"], projector_depth=2,
               diffusion_steps=100, tokenizer_model_max_length=None, tokenizer_padding_side="right")
    Lq = 2 + w["instr_tokens"] + w["template_tokens"]
    ids = torch.randint(1000, 30000, (B, Lq), generator=g) % w["llm"]["vocab_size"]
    ids[:, 1] = vla_oracle.IMAGE_TOKEN_INDEX
    img = w["vision"]["image_size"]
    batch = dict(input_ids=ids, attention_mask=torch.ones(B, Lq, dtype=torch.long),
                 images=torch.randn(B, 3, img, img, generator=g),
                 actions=torch.rand(B, w["chunk_size"] * w["action_dim"], generator=g) * 2 - 1,
                 noise=torch.randn(4 * B, w["chunk_size"], w["action_dim"], generator=g),
                 timesteps=torch.randint(0, 100, (4 * B,), generator=g), drop=torch.zeros(4 * B, dtype=torch.bool))
    opt = torch.optim.AdamW(list(sd.values()), lr=2e-5)
    _one_step(sd, cfg, batch, opt)                       # warm-up (allocations, thread pool)
    ts = sorted(_one_step(sd, cfg, batch, opt) for _ in range(max(1, steps)))
    return ts[len(ts) // 2]


def _per_layer(t1: float, t2: float) -> tuple[float, str]:
    """Seconds per decoder layer from the 1- and 2-layer step times.  A decoder layer is >= 30 % of the 2-layer step's
    arithmetic (6 * tokens * 233 M parameters against ViT-L + embedding AdamW + 2 layers), so when host noise makes
    t2 - t1 smaller than a quarter of t2 the difference is not a measurement: the floor is used and the JSON says so."""
    d = t2 - t1
    if d >= 0.25 * t2:
        return d, ""
    return 0.25 * t2, " [t2 - t1 below the arithmetic floor of one layer: 0.25 * t2 used per layer]"


def time_cogact_sample(w: dict, S: int, seconds_budget: float = 20.0, steps: int = 1, batch: int = 4) -> dict:
    # Threads: torch's own default (one per physical core).  bench.py starts this in a child process WITHOUT the
    # OMP_NUM_THREADS=1 that torchrun exports, so the same count applies at every N; forcing one thread per LOGICAL core
    # was measured to make the port several times slower on the GPU boxes' 2-socket hosts.
    if os.environ.get("B200_CPU_THREADS"):
        torch.set_num_threads(int(os.environ["B200_CPU_THREADS"]))
    B = batch
    t1 = _time_with_layers(w, 1, steps, B)
    t2 = _time_with_layers(w, 2, steps, B)
    n = w["llm"]["num_hidden_layers"]
    per_layer, note = _per_layer(t1, t2)
    total = t1 + (n - 1) * per_layer
    return {"value": round(B / total, 5), "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
            "extrapolated": True,
            "sample": (f"batch={B} fp32 oracle port, full-size ViT/projector/DiT + AdamW; decoder timed at 1 and 2 "
                       f"full-size layers ({t1:.2f}s, {t2:.2f}s per step) and extrapolated linearly to {n} layers "
                       f"-> {total:.1f} s/step of {B} samples{note}"),
            "seconds_per_sample": round(total / B, 3)}


def _time_reference_with_layers(w: dict, n_dec: int, steps: int, B: int) -> float:
    """One training step of the UNMODIFIED reference CogACTForCausalLM (vendored tree, oracle/ref_loader.py) on the host:
    fp32 parameters, the reference's own forward (HF Qwen2 + CLIP + its DiT / diffusion loss), backward, clip 1.0
    (trainer.py:122) and torch AdamW (trainer.py:25-36), `n_dec` full-size decoder layers."""
    from transformers import CLIPVisionConfig, Qwen2Config

    import bench
    from . import ref_loader
    llm = {k: v for k, v in w["llm"].items() if k != "model_type"}
    llm["num_hidden_layers"] = n_dec
    torch.manual_seed(1)
    model = ref_loader.build_reference_cogact(Qwen2Config(max_position_embeddings=4096, **llm),
                                              CLIPVisionConfig(**w["vision"]), w["action_model_type"],
                                              action_dim=w["action_dim"], chunk_size=w["chunk_size"])
    for p_ in model.model.parameters():          # base_exp.py:318-321: everything under model.model trains
        p_.requires_grad = True
    model.train()
    params = [p_ for p_ in model.parameters() if p_.requires_grad]
    opt = torch.optim.AdamW(params, lr=2e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)
    batch = {k: v for k, v in bench.make_batch(dict(w, batch=B), 0, pinned=False).items() if hasattr(v, "to")}

    def one_step() -> float:
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        out = model(**batch)
        out.loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        return time.perf_counter() - t0

    one_step()                                   # warm-up (allocations, thread pool, Adam state)
    ts = sorted(one_step() for _ in range(max(1, steps)))
    return ts[len(ts) // 2]


def time_reference_sample(w: dict, S: int, steps: int = 1, batch: int = 4) -> dict:
    """`kind: "reference"`: the same bounded sample as time_cogact_sample, run through the reference's own classes."""
    if os.environ.get("B200_CPU_THREADS"):
        torch.set_num_threads(int(os.environ["B200_CPU_THREADS"]))
    B = batch
    t1 = _time_reference_with_layers(w, 1, steps, B)
    t2 = _time_reference_with_layers(w, 2, steps, B)
    n = w["llm"]["num_hidden_layers"]
    per_layer, note = _per_layer(t1, t2)
    total = t1 + (n - 1) * per_layer
    return {"value": round(B / total, 5), "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "reference",
            "extrapolated": True,
            "sample": (f"batch={B} fp32, UNMODIFIED reference CogACTForCausalLM (baseline/_ref via the compat loader) + "
                       f"clip + torch AdamW on the host cores; full-size ViT / projector / DiT / embedding table, decoder "
                       f"timed at 1 and 2 full-size layers ({t1:.2f}s, {t2:.2f}s per step) and extrapolated linearly to "
                       f"{n} layers -> {total:.1f} s/step of {B} samples{note}"),
            "seconds_per_sample": round(total / B, 3)}


def main() -> None:
    """`python -m oracle.cpu_baseline <workload> <S> <steps> <batch> [port|reference]`: one JSON line.  bench.py runs the
    CPU arm in this child process under a wall-clock limit, so a slow host can never stall the GPU bench."""
    import json
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    import bench
    wl, S, steps, batch = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    kind = sys.argv[5] if len(sys.argv) > 5 else "port"
    fn = time_reference_sample if kind == "reference" else time_cogact_sample
    print(json.dumps(fn(bench.WORKLOADS[wl], S, steps=steps, batch=batch)), flush=True)


if __name__ == "__main__":
    main()
