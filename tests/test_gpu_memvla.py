"""GPU parity of MemVLA: the memory-path kernels against plain fp32 torch, and the full
training forward/backward (BottleneckSE, per/cog memory bank with token-merge consolidation, DiT with perceptual
cross-attention) against the golden vectors of the UNMODIFIED reference (dropout 0 on both sides; the reference
hard-codes 0.1, see oracle/ref_loader.py:build_reference_memvla).

Tolerances: bf16 trunk + bf16 memory modules, fp32/TF32 action head -> loss within 3e-2 relative, fused memory
tokens within 5e-2 relative Frobenius error, gradients cosine > 0.98 (the tiny widths 16/64 carry the most bf16
rounding noise; the per-kernel tests below are the tight ones)."""
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"


def _rel(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    return ((a - b).norm() / (b.norm() + 1e-12)).item(), torch.nn.functional.cosine_similarity(a, b, dim=0).item()


# ------------------------------------------------------------------------------------------ kernels
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_dropout_statistics_and_replay(dtype):
    from dexbotic_b200 import ops
    x = torch.ones(512, 1000, device="cuda", dtype=dtype)
    p = 0.1
    y = ops.dropout(x, p, seed=1234)
    keep = (y != 0).float().mean().item()
    assert abs(keep - (1 - p)) < 4e-3, keep                       # 512k Bernoulli draws: sigma = 4e-4
    assert torch.allclose(y[y != 0].float(), torch.full_like(y[y != 0].float(), 1 / (1 - p)), rtol=1e-2)
    assert torch.equal(y, ops.dropout(x, p, seed=1234))           # same seed -> same mask (backward replays it)
    assert not torch.equal(y, ops.dropout(x, p, seed=1235))
    # a row-padded view must see the same logical mask as the dense tensor
    xp = torch.ones(512, 1008, device="cuda", dtype=dtype)
    yp = ops.dropout(xp[:, :1000], p, seed=1234, out=torch.zeros_like(xp)[:, :1000])
    assert torch.equal(yp, y)
    # rows / columns are not correlated: every column keeps ~90 %
    assert ((y != 0).float().mean(0) - (1 - p)).abs().max().item() < 0.08


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_se_gate_matches_torch(dtype):
    from dexbotic_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    B, P, C = 3, 16, 64
    x = torch.randn(B, P, C, device="cuda", generator=g).to(dtype)
    y = torch.randn(B, P, C, device="cuda", generator=g).to(dtype)
    w = torch.rand(B, C, device="cuda", generator=g).to(dtype)
    a = torch.randn(B, C, device="cuda", generator=g).to(dtype)
    tol = 2e-2 if dtype == torch.bfloat16 else 1e-5
    assert torch.allclose(ops.se_reduce(x, None, 1.0 / P), x.float().mean(1), atol=tol)
    assert torch.allclose(ops.se_reduce(x, y, 1.0), (x.float() * y.float()).sum(1), atol=tol * 8, rtol=tol)
    assert torch.allclose(ops.se_scale(x, w).float(), x.float() * w.float()[:, None], atol=tol, rtol=tol)
    assert torch.allclose(ops.se_scale(x, w, a, 0.25).float(), x.float() * w.float()[:, None] + 0.25 * a.float()[:, None],
                          atol=tol * 2, rtol=tol)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_gate_fuse_matches_torch(dtype):
    from dexbotic_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    z, x1, x2, d = (torch.randn(40, 64, device="cuda", generator=g).to(dtype) for _ in range(4))
    zf, af, bf = (t.float().requires_grad_(True) for t in (z, x1, x2))
    s = torch.sigmoid(zf)
    ref = s * af + (1 - s) * bf
    ref.backward(d.float())
    tol = 2e-2 if dtype == torch.bfloat16 else 1e-5
    assert torch.allclose(ops.gate_fuse_fwd(z, x1, x2).float(), ref, atol=tol, rtol=tol)
    dz, d1, d2 = ops.gate_fuse_bwd(d, z, x1, x2)
    for got, want in ((dz, zf.grad), (d1, af.grad), (d2, bf.grad)):
        assert torch.allclose(got.float(), want, atol=tol, rtol=tol)


@pytest.mark.parametrize("dtype,B,Sq,Sk,H,hd", [(torch.bfloat16, 1, 1, 5, 4, 16), (torch.bfloat16, 2, 7, 40, 4, 64),
                                                (torch.bfloat16, 1, 4, 12, 4, 16), (torch.float32, 3, 68, 4, 4, 96),
                                                (torch.float32, 2, 17, 17, 4, 96), (torch.bfloat16, 1, 1, 3, 4, 896)])
def test_cross_attention_fwd_bwd(dtype, B, Sq, Sk, H, hd):
    """CrossAttnFn (separate query / key lengths, no mask) vs fp32 torch; shapes: cog role (1 query over T keys, head
    dim 16 tiny / 896 production), per role, DiT per_attn with the repeats stacked as query rows, DiT self-attention."""
    from dexbotic_b200.functional import CrossAttnFn
    g = torch.Generator(device="cuda").manual_seed(2)
    D = H * hd
    q, k, v = (torch.randn(B * n, D, device="cuda", generator=g).to(dtype).requires_grad_(True) for n in (Sq, Sk, Sk))
    do = torch.randn(B * Sq, D, device="cuda", generator=g).to(dtype)
    out = CrossAttnFn.apply(q, k, v, B, Sq, Sk, H)
    out.backward(do)
    qf, kf, vf = (t.detach().float().requires_grad_(True) for t in (q, k, v))
    sp = lambda t, n: t.view(B, n, H, hd).transpose(1, 2)  # noqa: E731
    ref = torch.nn.functional.scaled_dot_product_attention(sp(qf, Sq), sp(kf, Sk), sp(vf, Sk))
    ref = ref.transpose(1, 2).reshape(B * Sq, D)
    ref.backward(do.float())
    lim = 2e-2 if dtype == torch.bfloat16 else 3e-3            # bf16 probabilities / TF32 products
    for name, got, want in (("out", out, ref), ("dq", q.grad, qf.grad), ("dk", k.grad, kf.grad), ("dv", v.grad, vf.grad)):
        rel, cos = _rel(got, want)
        assert rel < lim * 2 and cos > 0.999, (name, rel, cos)


def test_cross_attention_dropout_is_consistent():
    """With dropout the backward must use the SAME mask as the forward: check dV = P_drop^T dO against the dropped
    probabilities reconstructed from a second forward with the same seed."""
    from dexbotic_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    B, Sq, Sk, H, hd = 1, 64, 96, 4, 64
    q, k, v = (torch.randn(B * n, H * hd, device="cuda", generator=g).bfloat16() for n in (Sq, Sk, Sk))
    o1, p1, pd1 = ops.cross_attention_fwd(q, k, v, B, Sq, Sk, H, dropout_p=0.1, seed=77)
    o2, p2, pd2 = ops.cross_attention_fwd(q, k, v, B, Sq, Sk, H, dropout_p=0.1, seed=77)
    assert torch.equal(o1, o2) and torch.equal(pd1, pd2)
    kept = (pd1[..., :Sk] != 0).float().mean().item()
    assert abs(kept - 0.9) < 0.02, kept
    ref = torch.einsum("bhqk,bkhd->bqhd", pd1[..., :Sk].float(), v.float().view(B, Sk, H, hd)).reshape(B * Sq, H * hd)
    rel, cos = _rel(o1, ref)
    assert rel < 2e-2, rel
    do = torch.randn(B * Sq, H * hd, device="cuda", generator=g).bfloat16()
    dq, dk, dv = ops.cross_attention_bwd(do, q, k, v, p1, pd1, B, Sq, Sk, H, dropout_p=0.1, seed=77)
    dv_ref = torch.einsum("bhqk,bqhd->bkhd", pd1[..., :Sk].float(), do.float().view(B, Sq, H, hd)).reshape(B * Sk, H * hd)
    assert _rel(dv, dv_ref)[0] < 2e-2
    # dq through the masked softmax gradient, fp32 reference with the same mask
    mask = (pd1[..., :Sk] != 0).float() / 0.9
    qf, kf = q.float().requires_grad_(True), k.float().requires_grad_(True)
    s = torch.einsum("bqhd,bkhd->bhqk", qf.view(B, Sq, H, hd), kf.view(B, Sk, H, hd)) * hd ** -0.5
    o = torch.einsum("bhqk,bkhd->bqhd", torch.softmax(s, -1) * mask, v.float().view(B, Sk, H, hd)).reshape(B * Sq, -1)
    o.backward(do.float())
    assert _rel(dq, qf.grad)[0] < 4e-2 and _rel(dk, kf.grad)[0] < 4e-2


# ------------------------------------------------------------------------------------------ model
def _build(fx, **over):
    from dexbotic_b200.model import MemVLAConfig, MemVLAForCausalLM
    from oracle.weights import seeded_state_dict
    cfg = fx["cfg"]
    mem = dict(cfg["mem"])
    mem.update(over)
    c = MemVLAConfig(llm_config=cfg["llm"], mm_vision_tower=cfg["vision"], mm_projector_type="mlp2x_gelu",
                     action_model_type=cfg["action_model_type"], action_dim=cfg["action_dim"],
                     chunk_size=cfg["chunk_size"], **mem)
    model = MemVLAForCausalLM(c, device="cuda")
    sd = {k: v for k, v in seeded_state_dict(fx["shapes"], fx["seed"]).items() if "position_ids" not in k}
    model.load_state_dict(sd, strict=True)
    return model


def test_memvla_state_dict_keys_match_reference():
    fx = torch.load(GOLDEN / "memvla_tiny.pt", weights_only=False)
    model = _build(fx, mem_dropout=0.0)
    ours = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert ours == {k: tuple(v) for k, v in fx["shapes"].items() if "position_ids" not in k}


def test_memvla_tiny_matches_reference_golden():
    fx = torch.load(GOLDEN / "memvla_tiny.pt", weights_only=False)
    model = _build(fx, mem_dropout=0.0)
    model.train()
    i = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in fx["inputs"].items()}
    model.zero_grad()
    out = model(input_ids=i["input_ids"], attention_mask=i["attention_mask"], images=i["images"], actions=i["actions"],
                indexes=i["indexes"], noise=i["noise"], timesteps=i["timesteps"], drop_mask=i["drop_mask"])
    ref = fx["outputs"]
    assert abs(out.loss.item() - ref["loss"].item()) < 3e-2 * abs(ref["loss"].item()), (out.loss.item(), ref["loss"].item())
    eng = model.model_engine
    with torch.no_grad():
        per = eng.per_compr(out.vision_proj_feats)
    rel, cos = _rel(per, ref["per_tokens"].cuda())
    assert rel < 5e-2 and cos > 0.998, ("per_compr", rel, cos)
    # episode (0,4) saw 3 frames with mem_length 2 -> one token merge; episode (0,9) holds its 2 frames
    bank = eng.per_cog_mem_bank
    assert sorted(len(v) for v in bank.banks["cog"].values()) == [2, 2]
    assert sorted(len(v) for v in bank.banks["per"].values()) == [2, 2]
    merged_t = bank.banks["cog"][(0, 4)][0][0].item()
    assert merged_t in (0.5, 0.0), merged_t            # frames 0,1 fused (t=0.5) unless 1,2 were the closest pair
    # last written entries are the fused tokens of the last frame of each episode
    rel, cos = _rel(bank.banks["cog"][(0, 9)][-1][1], ref["cog_fused"][4].cuda())
    assert rel < 5e-2, ("cog_fused", rel, cos)
    rel, cos = _rel(bank.banks["per"][(0, 9)][-1][1], ref["per_fused"][4].cuda())
    assert rel < 6e-2, ("per_fused", rel, cos)
    out.loss.backward()
    bad = []
    for name, gref in ref["grads"].items():
        g = model.store.g(name)
        assert g is not None, name
        rel, cos = _rel(g, gref.cuda())
        if not (rel < 0.25 and cos > 0.97):
            bad.append((name, round(rel, 4), round(cos, 5)))
    assert not bad, bad


def test_memvla_training_steps_reduce_loss_with_dropout():
    fx = torch.load(GOLDEN / "memvla_tiny.pt", weights_only=False)
    model = _build(fx)                                    # mem_dropout = 0.1 as the reference hard-codes
    model.train()
    i = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in fx["inputs"].items()}
    losses = []
    for _ in range(8):
        model.zero_grad()
        out = model(input_ids=i["input_ids"], attention_mask=i["attention_mask"], images=i["images"],
                    actions=i["actions"], indexes=i["indexes"], noise=i["noise"], timesteps=i["timesteps"],
                    drop_mask=i["drop_mask"])
        out.loss.backward()
        model.optimizer_step(base_lr=3e-3, max_grad_norm=1.0)
        losses.append(out.loss.item())
    assert all(torch.isfinite(torch.tensor(losses))), losses
    assert losses[-1] < 0.7 * losses[0], losses


def test_memvla_inference_is_stateful():
    fx = torch.load(GOLDEN / "memvla_tiny.pt", weights_only=False)
    model = _build(fx, mem_dropout=0.0)
    model.eval()
    i = fx["inputs"]
    ids, img = i["input_ids"][:1].cuda(), i["images"][:1].cuda()
    ids = ids[:, : int(i["attention_mask"][0].sum())]
    norms = {"action_norms": {"min": [-1.0] * 7, "max": [1.0] * 7}, "cfg_scale": 1.5, "num_ddim_steps": 4}
    noise = torch.randn(1, 16, 7, generator=torch.Generator().manual_seed(5)).cuda()
    a0 = model.inference_action(ids, img, "True", norms, noise=noise)
    bank = model.model_engine.per_cog_mem_bank
    assert len(bank.banks["cog"][(0, 0)]) == 1 and model.cur_timestep == 1
    a1 = model.inference_action(ids, img, "False", norms, noise=noise)
    assert len(bank.banks["cog"][(0, 0)]) == 2 and model.cur_timestep == 2
    a2 = model.inference_action(ids, img, "False", norms, noise=noise)
    assert len(bank.banks["cog"][(0, 0)]) == 2, "mem_length 2: consolidation keeps the bank bounded"
    t = torch.tensor([a0, a1, a2])
    assert t.shape == (3, 16, 7) and torch.isfinite(t).all() and t.abs().max() <= 1.0
    assert not torch.equal(t[0], t[1]), "the second frame must see the memory of the first"
    b0 = model.inference_action(ids, img, "True", norms, noise=noise)
    assert torch.allclose(torch.tensor(b0), t[0], atol=1e-6), "episode_first_frame='True' resets bank and timestep"


def test_memvla_production_dims_one_layer_matches_oracle():
    """Production widths (Qwen2.5-7B d=3584 -> cog role head_dim 896, FFN 14336; per_token_size 256, 256 vision
    tokens; CLIP-L width; DiT-B with per_attn) with one decoder layer: the tile / head geometry the real model hits."""
    from oracle import vla_oracle
    from oracle.weights import seeded_state_dict
    from dexbotic_b200.model import MemVLAConfig, MemVLAForCausalLM
    mem = dict(dataloader_type="group", group_size=3, per_token_size=256, mem_length=2, retrieval_layers=2,
               use_timestep_pe=True, fusion_type="gate", consolidate_type="tome", update_fused=True)
    cfg = dict(
        llm=dict(vocab_size=2048, hidden_size=3584, intermediate_size=18944, num_hidden_layers=1, num_attention_heads=28,
                 num_key_value_heads=4, rope_theta=1e6, rms_norm_eps=1e-6, hidden_act="silu", model_type="qwen2"),
        vision=dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=2, num_attention_heads=16, image_size=224,
                    patch_size=14, hidden_act="quick_gelu", layer_norm_eps=1e-5),
        action_model_type="DiT-B", action_dim=7, chunk_size=16, projector_depth=2, diffusion_steps=100,
        tokenizer_model_max_length=None, tokenizer_padding_side="right", mem=mem)
    c = MemVLAConfig(llm_config=cfg["llm"], mm_vision_tower=cfg["vision"], action_model_type="DiT-B", action_dim=7,
                     chunk_size=16, mem_dropout=0.0, **mem)
    model = MemVLAForCausalLM(c)
    sd = seeded_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, 11)
    model.load_state_dict(sd)
    model.train()
    g = torch.Generator().manual_seed(13)
    B, L, R = 4, 40, 4
    ids = torch.randint(1, 2048, (B, L), generator=g)
    ids[:, 1] = -200
    mask = torch.ones(B, L, dtype=torch.long)
    mask[2, 33:] = 0
    images = torch.randn(B, 3, 224, 224, generator=g)
    actions = torch.rand(B, 112, generator=g) * 2 - 1
    indexes = [(0, 1, 0), (0, 1, 1), (0, 1, 2), (0, 2, 7)]          # 3 frames of one episode (one token merge) + 1
    noise = torch.randn(R * B, 16, 7, generator=g)
    t = torch.randint(0, 100, (R * B,), generator=g)
    drop = torch.tensor([False, False, False, True] * R)
    names = ["model.per_compr.excite.1.weight", "model.per_compr.reduce.0.weight", "model.per_compr.reduce.2.weight",
             "model.per_cog_mem_bank.retrieval_blocks.cog.0.q_proj.weight",
             "model.per_cog_mem_bank.retrieval_blocks.cog.1.k_proj.weight",
             "model.per_cog_mem_bank.retrieval_blocks.cog.1.ffn.0.weight",
             "model.per_cog_mem_bank.retrieval_blocks.per.0.v_proj.weight",
             "model.per_cog_mem_bank.retrieval_blocks.per.1.ffn.3.weight",
             "model.per_cog_mem_bank.gate_fusion_blocks.cog.proj.weight",
             "model.per_cog_mem_bank.timestep_embedders.per.mlp.2.weight",
             "model.action_head.net.per_token_embedder.linear.weight",
             "model.action_head.net.blocks.0.per_attn.in_proj_weight",
             "model.action_head.net.blocks.11.per_attn.out_proj.weight", "model.action_head.net.blocks.5.norm3.weight",
             "model.mm_projector.2.weight", "model.llm.layers.0.mlp.down_proj.weight"]
    sd_g = {k: (v.clone().requires_grad_(True) if k in names else v) for k, v in sd.items()}
    ora = vla_oracle.memvla_forward(sd_g, cfg, ids, mask, images, actions, indexes, noise, t, drop, R)
    ora["loss"].backward()
    model.zero_grad()
    out = model(input_ids=ids.cuda(), attention_mask=mask.cuda(), images=images.cuda(), actions=actions.cuda(),
                indexes=indexes, repeated_diffusion_steps=R, noise=noise.cuda(), timesteps=t.cuda(), drop_mask=drop.cuda())
    assert abs(out.loss.item() - ora["loss"].item()) < 3e-2 * abs(ora["loss"].item()), (out.loss.item(), ora["loss"].item())
    bank = model.model_engine.per_cog_mem_bank
    rel, cos = _rel(bank.banks["cog"][(0, 2)][-1][1].cpu(), ora["cog_fused"][3].detach())
    assert rel < 4e-2 and cos > 0.999, ("cog_fused", rel, cos)
    rel, cos = _rel(bank.banks["per"][(0, 2)][-1][1].cpu(), ora["per_fused"][3].detach())
    assert rel < 4e-2 and cos > 0.999, ("per_fused", rel, cos)
    out.loss.backward()
    bad = []
    for name in names:
        rel, cos = _rel(model.store.g(name).cpu(), sd_g[name].grad)
        if not (rel < 0.2 and cos > 0.98):
            bad.append((name, round(rel, 4), round(cos, 5)))
    assert not bad, bad


def test_memvla_inference_matches_reference_golden():
    """Four consecutive inference_action calls of one episode (memory bank grows to mem_length 2, then token-merges)
    against the reference's own outputs on the same noise draws (memvla_arch.py:666-745), dropout 0."""
    fx = torch.load(GOLDEN / "memvla_inference_tiny.pt", weights_only=False)
    model = _build(fx, mem_dropout=0.0)
    model.eval()
    args = {"cfg_scale": 1.5, "num_ddim_steps": 10, "action_norms": {"min": [-1.0] * 7, "max": [1.0] * 7}}
    errs = []
    for f, fr in enumerate(fx["frames"]):
        acts = model.inference_action(fr["input_ids"].cuda(), fr["images"].cuda(), "True" if f == 0 else "False", args,
                                      noise=fr["noise"].cuda())
        got, ref = torch.tensor(acts), fr["actions"]
        assert got.shape == ref.shape == (16, 7)
        errs.append(((got - ref).abs().max().item(), (got - ref).abs().mean().item()))
    print("per-frame (max, mean) abs error:", [(round(a, 4), round(b, 4)) for a, b in errs])
    # ten guided DiT evaluations amplify the bf16 rounding of the trunk and of the memory tokens (tiny widths 32 / 64),
    # and later frames read bf16 memories written by earlier ones: max 0.1 / mean 0.02 absolute on actions in [-1, 1]
    # (CogACT's sampler, without the memory path, is held to 0.05)
    assert max(a for a, _ in errs) < 0.1 and max(b for _, b in errs) < 0.02, errs
    assert len(model.model_engine.per_cog_mem_bank.banks["cog"][(0, 0)]) == 2 and model.cur_timestep == 4
