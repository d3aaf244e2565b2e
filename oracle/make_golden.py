"""TEST INFRASTRUCTURE — generates tests/golden/*.pt by running the UNMODIFIED reference
(/root/reference, via oracle/ref_loader.py) in THIS container, and checks the oracle restatement
(oracle/vla_oracle.py) against it.  Run:  python -m oracle.make_golden

Fixtures hold inputs + the reference's outputs only; weights are regenerated from a per-tensor seed
(oracle.weights.seeded_state_dict) so the files stay small.  The GPU box has no /root/reference: there
the committed fixtures + the restatement are the checker.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import ref_loader, vla_oracle  # noqa: E402
from oracle.weights import seeded_state_dict  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"


def tiny_cogact_configs():
    from transformers import CLIPVisionConfig, Qwen2Config
    llm = Qwen2Config(vocab_size=128, hidden_size=64, intermediate_size=160, num_hidden_layers=2,
                      num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=256, rope_theta=1e6,
                      rms_norm_eps=1e-6)
    clip = CLIPVisionConfig(hidden_size=32, intermediate_size=64, num_hidden_layers=3, num_attention_heads=2,
                            image_size=28, patch_size=14, hidden_act="quick_gelu")
    cfg = dict(
        llm=dict(vocab_size=128, hidden_size=64, intermediate_size=160, num_hidden_layers=2, num_attention_heads=4,
                 num_key_value_heads=2, rope_theta=1e6, rms_norm_eps=1e-6, hidden_act="silu", attention_bias=True,
                 model_type="qwen2"),
        vision=dict(hidden_size=32, intermediate_size=64, num_hidden_layers=3, num_attention_heads=2, image_size=28,
                    patch_size=14, hidden_act="quick_gelu", layer_norm_eps=1e-5),
        action_model_type="DiT-S", action_dim=7, chunk_size=16, projector_depth=2, diffusion_steps=100,
        tokenizer_model_max_length=None, tokenizer_padding_side="right")
    return llm, clip, cfg


def make_cogact_tiny(seed: int = 1234):
    llm, clip, cfg = tiny_cogact_configs()
    model = ref_loader.build_reference_cogact(llm, clip, "DiT-S")
    sd = seeded_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed)
    missing = model.load_state_dict(sd, strict=True)
    model.train()

    g = torch.Generator().manual_seed(seed)
    B, L = 3, 14
    ids = torch.randint(1, 128, (B, L), generator=g)
    ids[:, 1] = vla_oracle.IMAGE_TOKEN_INDEX
    mask = torch.ones(B, L, dtype=torch.long)
    mask[1, 10:] = 0
    mask[2, 12:] = 0
    images = torch.randn(B, 3, 28, 28, generator=g)
    actions = torch.rand(B, 16 * 7, generator=g) * 2 - 1

    # the reference draws noise / timesteps / drop ids from the global RNG, in this order:
    #   torch.randn_like(x) (action_models.py:106), torch.randint (:107-109), torch.rand (dit.py:86-88)
    R = 4
    torch.manual_seed(seed + 1)
    out = model(input_ids=ids, attention_mask=mask, images=images, actions=actions, repeated_diffusion_steps=R)
    out.loss.backward()
    torch.manual_seed(seed + 1)
    noise = torch.randn(R * B, 16, 7)
    timesteps = torch.randint(0, 100, (R * B,))
    drop = torch.rand(R * B) < 0.1

    ora = vla_oracle.cogact_forward(sd, cfg, ids, mask, images, actions, noise, timesteps, drop, R)
    err_loss = abs(ora["loss"].item() - out.loss.item())
    valid = ora["attention_mask"][:, :, None]      # rows at padded positions are unspecified (never consumed)
    err_hid = ((ora["last_hidden"] - out.logits) * valid).abs().max().item()
    print(f"[cogact_tiny] reference loss {out.loss.item():.8f} oracle {ora['loss'].item():.8f} |d|={err_loss:.2e}; "
          f"last_hidden max|d|={err_hid:.2e}")
    assert err_loss < 1e-5 and err_hid < 1e-4, "oracle restatement disagrees with the reference"

    # reference gradients for a few parameters (checks the product's backward)
    grads = {}
    for name in ["model.llm.layers.0.self_attn.q_proj.weight", "model.llm.layers.1.mlp.down_proj.weight",
                 "model.llm.embed_tokens.weight", "model.mm_projector.0.weight",
                 "model.action_head.net.blocks.0.attn.qkv.weight", "model.action_head.net.z_embedder.linear.weight",
                 "model.llm.norm.weight", "model.llm.layers.0.self_attn.k_proj.bias"]:
        p = dict(model.named_parameters())[name]
        grads[name] = p.grad.clone()
    # oracle gradient check (autograd through the restatement)
    sd_g = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point()}
    ora_g = vla_oracle.cogact_forward(sd_g, cfg, ids, mask, images, actions, noise, timesteps, drop, R)
    ora_g["loss"].backward()
    for name, gref in grads.items():
        d = (sd_g[name].grad - gref).abs().max().item()
        assert d < 1e-5 + 1e-3 * gref.abs().max().item(), f"oracle grad mismatch for {name}: {d}"
    fixture = dict(
        seed=seed, cfg=cfg, shapes={k: tuple(v.shape) for k, v in sd.items()},
        inputs=dict(input_ids=ids, attention_mask=mask, images=images, actions=actions, noise=noise,
                    timesteps=timesteps, drop_mask=drop, repeated_diffusion_steps=R),
        outputs=dict(loss=out.loss.detach(), last_hidden=(out.logits * valid).detach(),
                     valid=ora["attention_mask"], cognition=ora["cognition"].detach(),
                     grads=grads))
    GOLDEN.mkdir(parents=True, exist_ok=True)
    torch.save(fixture, GOLDEN / "cogact_tiny.pt")
    print(f"[cogact_tiny] wrote {GOLDEN / 'cogact_tiny.pt'}")


def make_splice_cases(seed: int = 7):
    """Splice layouts from the reference's own _prepare_inputs_labels_for_multimodal (dexbotic_arch.py:182-373)."""
    llm, clip, cfg = tiny_cogact_configs()
    model = ref_loader.build_reference_cogact(llm, clip, "DiT-S")
    sd = seeded_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed)
    model.load_state_dict(sd)
    model.eval()
    g = torch.Generator().manual_seed(seed)
    cases = []
    for name, n_img, side, max_len in [("one_image_right", 1, "right", None), ("no_image_row", 1, "right", None),
                                       ("two_images_left", 2, "left", None), ("truncate", 1, "right", 9)]:
        B, L = 4, 12
        ids = torch.randint(1, 128, (B, L), generator=g)
        labels = torch.randint(0, 128, (B, L), generator=g)
        mask = torch.ones(B, L, dtype=torch.long)
        mask[0, 8:] = 0
        mask[3, 11:] = 0
        n_entries = 0
        for b in range(B):
            k = 0 if (name == "no_image_row" and b == 2) else n_img
            pos = torch.randperm(7, generator=g)[:k] + 1
            ids[b, pos] = vla_oracle.IMAGE_TOKEN_INDEX
            n_entries += max(k, 1)
        images = torch.randn(n_entries, 3, 28, 28, generator=g)
        model.model.config.tokenizer_padding_side = side
        model.model.config.tokenizer_model_max_length = max_len
        with torch.no_grad():
            _, pid, m, _, emb, lab, _ = model.model._prepare_inputs_labels_for_multimodal(
                ids, None, mask, None, labels, None, images)
            feats = model.model._extract_vision_features(images)
        o_emb, o_lab, o_msk, o_pid = vla_oracle.splice(sd["model.llm.embed_tokens.weight"], feats, ids, mask, labels,
                                                       max_len, side)
        assert torch.equal(o_emb, emb) and torch.equal(o_lab, lab) and torch.equal(o_msk, m.bool())
        cases.append(dict(name=name, input_ids=ids, attention_mask=mask, labels=labels, padding_side=side,
                          max_len=max_len, n_img_tokens=feats.shape[1], image_features=feats,
                          embed_weight=sd["model.llm.embed_tokens.weight"], inputs_embeds=emb, new_labels=lab,
                          new_mask=m.bool()))
        print(f"[splice] {name}: S={emb.shape[1]} ok")
    torch.save(cases, GOLDEN / "splice_cases.pt")


def make_integer_kats():
    """Known-answer vectors for the OFT discrete tokenizer, from the reference's own code."""
    ref_loader.load_reference()
    from dexbotic.model.oft.action_model.model import DiscreteActionHead
    from dexbotic.data.dataset.transform.action import ActionNormAnd2String
    head = DiscreteActionHead(input_dim=64, vocab_size=512, action_dim=7, action_chunk=8, num_bins=256)
    g = torch.Generator().manual_seed(99)
    k = torch.arange(0, 255, dtype=torch.float32)
    a = torch.cat([torch.rand(4096, generator=g) * 2.4 - 1.2,
                   torch.tensor([-1.0, 1.0, 0.0, 1 / 255, -1 / 255, 0.5 / 255, 1.5 / 255, 2.5 / 255]),
                   (k + 0.5) / 255 * 2 - 1])           # exact ties under the x255 map (half-to-even)
    a = a[: (a.numel() // 56) * 56].reshape(-1, 8, 7)
    bins = head.discretize_actions(a)
    cont = head.discrete_tokens_to_continuous(bins.reshape(bins.shape[0], -1))
    assert np.array_equal(vla_oracle.oft_discretize(a.numpy()), bins.numpy())
    assert np.array_equal(vla_oracle.oft_bins_to_continuous(bins.numpy()), cont.numpy())
    logits = torch.randn(2, 56, 512, generator=g)
    logits[0, 3, -255:] = 0.25
    idx = torch.argmax(logits[:, :, -255:], dim=-1)     # oft_discrete_arch.py:222-224
    assert np.array_equal(vla_oracle.oft_argmax_decode(logits.numpy()), idx.numpy())
    t = ActionNormAnd2String.__new__(ActionNormAnd2String)
    a64 = (torch.rand(512, 7, generator=g, dtype=torch.float64) * 2.2 - 1.1).numpy()
    data_bins = t._action2bin(a64, 255)
    assert np.array_equal(vla_oracle.data_action_to_bin(a64, 255), data_bins)
    torch.save(dict(actions=a, bins=bins, continuous=cont, logits=logits, argmax=idx,
                    data_actions=torch.from_numpy(a64), data_bins=torch.from_numpy(data_bins)),
               GOLDEN / "oft_integer_kats.pt")
    print("[oft] integer KATs ok")
    # cosine schedule constants observed in the reference (SURVEY.md §8c golden (i))
    from dexbotic.model.cogact.action_model.diffusion import create_diffusion
    d = create_diffusion(timestep_respacing="", noise_schedule="squaredcos_cap_v2", diffusion_steps=100,
                         sigma_small=True, learn_sigma=False)
    sa, sb = vla_oracle.cosine_schedule(100)
    assert np.array_equal(sa, d.sqrt_alphas_cumprod) and np.array_equal(sb, d.sqrt_one_minus_alphas_cumprod)
    torch.save(dict(sqrt_alphas_cumprod=torch.from_numpy(d.sqrt_alphas_cumprod),
                    sqrt_one_minus_alphas_cumprod=torch.from_numpy(d.sqrt_one_minus_alphas_cumprod)),
               GOLDEN / "cosine_schedule_T100.pt")
    print("[diffusion] schedule ok", d.sqrt_alphas_cumprod[[0, 50, 99]])


def make_oft_discrete_tiny(seed: int = 4321):
    llm, clip, cfg = tiny_cogact_configs()
    llm.vocab_size = 512
    cfg = dict(cfg, chunk_size=8, action_dim=7, num_bins=256)
    cfg["llm"] = dict(cfg["llm"], vocab_size=512)
    model = ref_loader.build_reference_oft_discrete(llm, clip, 7, 8, 256)
    sd = seeded_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed)
    model.load_state_dict(sd, strict=True)
    model.train()
    g = torch.Generator().manual_seed(seed)
    B, A = 3, 56
    L = 10 + A + 1 + 3                       # prefix(10) + action labels(56) + suffix token + up to 3 pads
    ids = torch.randint(1, 250, (B, L), generator=g)
    ids[:, 1] = vla_oracle.IMAGE_TOKEN_INDEX
    mask = torch.ones(B, L, dtype=torch.long)
    mask[1, L - 3:] = 0
    mask[2, L - 1:] = 0
    labels = torch.full((B, L), -100, dtype=torch.long)
    for b in range(B):
        npl = int(mask[b].sum())
        tok = torch.randint(512 - 255, 512, (A,), generator=g)      # action tokens live in the last 255 vocab entries
        ids[b, npl - 1 - A:npl - 1] = tok
        labels[b, npl - 1 - A:npl - 1] = tok
    images = torch.randn(B, 3, 28, 28, generator=g)
    actions = torch.rand(B, 8 * 7, generator=g) * 2 - 1      # only `actions is not None` matters (oft_discrete_arch.py:171)
    out = model(input_ids=ids, attention_mask=mask, images=images, labels=labels, actions=actions)
    out.loss.backward()
    ora = vla_oracle.oft_discrete_forward(sd, cfg, ids, mask, images, labels)
    d_loss = abs(ora["loss"].item() - out.loss.item())
    d_log = (ora["logits"] - out.logits).abs().max().item()
    print(f"[oft_discrete_tiny] reference loss {out.loss.item():.8f} oracle {ora['loss'].item():.8f}; logits max|d|={d_log:.2e}")
    assert d_loss < 1e-5 and d_log < 1e-4
    model.eval()
    with torch.no_grad():
        inf = model(input_ids=ids[:, :11], attention_mask=torch.ones(B, 11, dtype=torch.long), images=images)
        ref_idx = torch.argmax(inf.logits[:, :, -255:], dim=-1)          # oft_discrete_arch.py:222-224
        ref_cont = model.model.action_head.discrete_tokens_to_continuous(ref_idx)
    grads = {n: dict(model.named_parameters())[n].grad.clone() for n in
             ["lm_head.weight", "model.llm.layers.1.mlp.up_proj.weight", "model.llm.embed_tokens.weight",
              "model.mm_projector.2.weight", "model.llm.layers.0.self_attn.v_proj.weight"]}
    torch.save(dict(seed=seed, cfg=cfg, shapes={k: tuple(v.shape) for k, v in sd.items()},
                    inputs=dict(input_ids=ids, attention_mask=mask, images=images, labels=labels, actions=actions),
                    outputs=dict(loss=out.loss.detach(), logits=out.logits.detach(), grads=grads,
                                 infer_logits=inf.logits, infer_idx=ref_idx, infer_cont=ref_cont)),
               GOLDEN / "oft_discrete_tiny.pt")
    print("[oft_discrete_tiny] wrote fixture")


def make_oft_discrete_proprio_tiny(seed: int = 4322):
    """OFT-discrete with the proprio state token (oft_discrete_arch.py:132-137,161-162)."""
    llm, clip, cfg = tiny_cogact_configs()
    llm.vocab_size = 512
    cfg = dict(cfg, chunk_size=8, action_dim=7, num_bins=256, use_proprio=True, proprio_dim=9)
    cfg["llm"] = dict(cfg["llm"], vocab_size=512)
    model = ref_loader.build_reference_oft_discrete(llm, clip, 7, 8, 256, use_proprio=True, proprio_dim=9)
    sd = seeded_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed)
    model.load_state_dict(sd, strict=True)
    model.train()
    g = torch.Generator().manual_seed(seed)
    B, A = 3, 56
    L = 10 + A + 1 + 3
    ids = torch.randint(1, 250, (B, L), generator=g)
    ids[:, 1] = vla_oracle.IMAGE_TOKEN_INDEX
    mask = torch.ones(B, L, dtype=torch.long)
    mask[1, L - 3:] = 0
    mask[2, L - 1:] = 0
    labels = torch.full((B, L), -100, dtype=torch.long)
    for b in range(B):
        npl = int(mask[b].sum())
        tok = torch.randint(512 - 255, 512, (A,), generator=g)
        ids[b, npl - 1 - A:npl - 1] = tok
        labels[b, npl - 1 - A:npl - 1] = tok
    images = torch.randn(B, 3, 28, 28, generator=g)
    actions = torch.rand(B, 8 * 7, generator=g) * 2 - 1
    states = torch.randn(B, 9, generator=g)
    out = model(input_ids=ids, attention_mask=mask, images=images, labels=labels, actions=actions, states=states)
    out.loss.backward()
    ora = vla_oracle.oft_discrete_forward(sd, cfg, ids, mask, images, labels, states)
    d_loss = abs(ora["loss"].item() - out.loss.item())
    d_log = (ora["logits"] - out.logits).abs().max().item()
    print(f"[oft_discrete_proprio] reference loss {out.loss.item():.8f} oracle {ora['loss'].item():.8f}; logits max|d|={d_log:.2e}")
    assert d_loss < 1e-5 and d_log < 1e-4
    params = dict(model.named_parameters())
    grads = {n: params[n].grad.clone() for n in
             ["lm_head.weight", "model.action_head.proprio_projector.fc1.weight",
              "model.action_head.proprio_projector.fc2.bias", "model.llm.layers.1.mlp.up_proj.weight",
              "model.llm.embed_tokens.weight"]}
    torch.save(dict(seed=seed, cfg=cfg, shapes={k: tuple(v.shape) for k, v in sd.items()},
                    inputs=dict(input_ids=ids, attention_mask=mask, images=images, labels=labels, actions=actions,
                                states=states),
                    outputs=dict(loss=out.loss.detach(), logits=out.logits.detach(), grads=grads)),
               GOLDEN / "oft_discrete_proprio_tiny.pt")
    print("[oft_discrete_proprio] wrote fixture")


def make_oft_linear_tiny(seed: int = 8642):
    """OFTForCausalLM with the L1-regression head, with and without the proprio token (oft_arch.py:58-166)."""
    out_fx = {}
    for use_proprio in (False, True):
        llm, clip, cfg = tiny_cogact_configs()
        cfg = dict(cfg, chunk_size=8, action_dim=7, use_proprio=use_proprio, proprio_dim=9 if use_proprio else None)
        model = ref_loader.build_reference_oft_linear(llm, clip, 7, 8, use_proprio, cfg["proprio_dim"])
        sd = seeded_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed)
        model.load_state_dict(sd, strict=True)
        model.train()
        g = torch.Generator().manual_seed(seed)
        B, L = 3, 12
        ids = torch.randint(1, 128, (B, L), generator=g)
        ids[:, 1] = vla_oracle.IMAGE_TOKEN_INDEX
        mask = torch.ones(B, L, dtype=torch.long)
        mask[1, L - 3:] = 0
        mask[2, L - 1:] = 0
        images = torch.randn(B, 3, 28, 28, generator=g)
        actions = torch.rand(B, 8 * 7, generator=g) * 2 - 1
        states = torch.randn(B, 9, generator=g) if use_proprio else None
        out = model(input_ids=ids, attention_mask=mask, images=images, actions=actions, states=states)
        out.loss.backward()
        ora = vla_oracle.oft_l1_forward(sd, cfg, ids, mask, images, actions, states)
        d_loss = abs(ora["loss"].item() - out.loss.item())
        d_pred = (ora["predicted_actions"] - out.logits).abs().max().item()
        print(f"[oft_linear_tiny proprio={use_proprio}] reference loss {out.loss.item():.8f} oracle "
              f"{ora['loss'].item():.8f}; actions max|d|={d_pred:.2e}")
        assert d_loss < 1e-5 and d_pred < 1e-4
        names = ["model.action_head.action_query", "model.action_head.model.fc1.weight",
                 "model.action_head.model.layer_norm1.weight", "model.action_head.model.mlp_resnet_blocks.1.ffn.1.weight",
                 "model.action_head.model.fc2.bias", "model.llm.layers.1.mlp.up_proj.weight",
                 "model.llm.layers.0.self_attn.v_proj.weight", "model.mm_projector.2.weight",
                 "model.llm.embed_tokens.weight"]
        if use_proprio:
            names += ["model.action_head.proprio_projector.fc1.weight", "model.action_head.proprio_projector.fc2.bias"]
        params = dict(model.named_parameters())
        grads = {n: params[n].grad.clone() for n in names}
        out_fx[use_proprio] = dict(cfg=cfg, shapes={k: tuple(v.shape) for k, v in sd.items()},
                                   inputs=dict(input_ids=ids, attention_mask=mask, images=images, actions=actions,
                                               states=states),
                                   outputs=dict(loss=out.loss.detach(), predicted_actions=out.logits.detach(),
                                                grads=grads))
    torch.save(dict(seed=seed, cases=out_fx), GOLDEN / "oft_linear_tiny.pt")
    print("[oft_linear_tiny] wrote fixture")


def make_oft_diffusion_tiny(seed: int = 8643):
    """OFTForCausalLM with the `DiT` DiffusionActionHead (oft_arch.py:103-154, model.py:197-271), with and without the
    proprio token: training forward/backward at a fixed noisy_dict, and the DDIM inference loop from a fixed start.  The
    reference code runs around the restated scheduler (oracle/ddim_oracle.py): the scheduler's parity is unpinned."""
    out_fx = {}
    for use_proprio in (False, True):
        llm, clip, cfg = tiny_cogact_configs()
        cfg = dict(cfg, chunk_size=8, action_dim=7, use_proprio=use_proprio, proprio_dim=9 if use_proprio else None)
        model = ref_loader.build_reference_oft_diffusion(llm, clip, 7, 8, use_proprio, cfg["proprio_dim"])
        sd = seeded_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed)
        model.load_state_dict(sd, strict=True)
        model.train()
        g = torch.Generator().manual_seed(seed)
        B, L = 3, 12
        ids = torch.randint(1, 128, (B, L), generator=g)
        ids[:, 1] = vla_oracle.IMAGE_TOKEN_INDEX
        mask = torch.ones(B, L, dtype=torch.long)
        mask[1, L - 3:] = 0
        mask[2, L - 1:] = 0
        images = torch.randn(B, 3, 28, 28, generator=g)
        actions = torch.rand(B, 8 * 7, generator=g) * 2 - 1
        states = torch.randn(B, 9, generator=g) if use_proprio else None
        torch.manual_seed(seed + 1)
        noisy = model.model.action_head.sample_noisy_actions(actions.reshape(B, 8, 7))
        out = model(input_ids=ids, attention_mask=mask, images=images, actions=actions, states=states, noisy_dict=noisy)
        out.loss.backward()
        ora = vla_oracle.oft_diffusion_forward(sd, cfg, ids, mask, images, noisy, actions, states)
        d_loss = abs(ora["loss"].item() - out.loss.item())
        d_pred = (ora["predicted_noise"] - out.logits).abs().max().item()
        print(f"[oft_diffusion_tiny proprio={use_proprio}] reference loss {out.loss.item():.8f} oracle "
              f"{ora['loss'].item():.8f}; noise max|d|={d_pred:.2e}")
        assert d_loss < 1e-5 and d_pred < 1e-4
        names = ["model.action_head.noisy_action_projector.fc1.weight", "model.action_head.noisy_action_projector.fc2.weight",
                 "model.action_head.noise_predictor.mlp_resnet.fc1.weight",
                 "model.action_head.noise_predictor.mlp_resnet.layer_norm1.weight",
                 "model.action_head.noise_predictor.mlp_resnet.mlp_resnet_blocks.1.ffn.1.weight",
                 "model.action_head.noise_predictor.mlp_resnet.fc2.bias", "model.llm.layers.1.mlp.up_proj.weight",
                 "model.llm.layers.0.self_attn.v_proj.weight", "model.mm_projector.2.weight",
                 "model.llm.embed_tokens.weight"]
        if use_proprio:
            names += ["model.action_head.proprio_projector.fc1.weight", "model.action_head.proprio_projector.fc2.bias"]
        params = dict(model.named_parameters())
        grads = {n: params[n].grad.clone() for n in names}
        # inference: the reference draws its own start noise from the global RNG (oft_arch.py:226-229)
        model.eval()
        norms = dict(min=[-1.0] * 7, max=[1.0] * 7)
        traj = []                                   # (timestep token, x_t fed to the model, predicted noise) per DDIM step
        hook = model.register_forward_hook(
            lambda m, a, kw, out: traj.append((kw["noisy_dict"]["diffusion_timestep_embeddings"].clone(),
                                               kw["noisy_dict"]["noisy_actions"].clone(), out.logits.clone())),
            with_kwargs=True)
        torch.manual_seed(seed + 2)
        ref_actions = model.inference_action(ids[:1], images[:1], dict(action_norms=norms, num_ddim_steps=5,
                                                                       states=states[:1] if use_proprio else None))
        hook.remove()
        assert len(traj) == 5
        torch.manual_seed(seed + 2)
        start = torch.randn(1, 8, 7, dtype=images.dtype)
        ora_actions = vla_oracle.oft_diffusion_inference(sd, cfg, ids[:1], images[:1], start, 5,
                                                         states[:1] if use_proprio else None)
        d_inf = (torch.tensor(ref_actions) - ora_actions[0].clamp(-1, 1)).abs().max().item()
        print(f"[oft_diffusion_tiny proprio={use_proprio}] inference max|d|={d_inf:.2e}")
        assert d_inf < 1e-4
        out_fx[use_proprio] = dict(cfg=cfg, shapes={k: tuple(v.shape) for k, v in sd.items()},
                                   inputs=dict(input_ids=ids, attention_mask=mask, images=images, actions=actions,
                                               states=states, noisy_dict={k: v.detach() for k, v in noisy.items()},
                                               start_noise=start, num_ddim_steps=5),
                                   outputs=dict(loss=out.loss.detach(), predicted_noise=out.logits.detach(),
                                                grads=grads, inference_actions=torch.tensor(ref_actions),
                                                trajectory=[tuple(x.detach() for x in st) for st in traj]))
    torch.save(dict(seed=seed, cases=out_fx), GOLDEN / "oft_diffusion_tiny.pt")
    print("[oft_diffusion_tiny] wrote fixture")


def tiny_pi0_configs():
    llm = dict(model_type="gemma", vocab_size=128, hidden_size=64, intermediate_size=128, num_hidden_layers=2,
               num_attention_heads=4, num_key_value_heads=1, head_dim=16, rms_norm_eps=1e-6, rope_theta=10000.0,
               hidden_act="gelu_pytorch_tanh")
    exp = dict(llm, hidden_size=32, intermediate_size=64)
    vis = dict(model_type="siglip_vision_model", hidden_size=32, intermediate_size=64, num_hidden_layers=2,
               num_attention_heads=2, image_size=28, patch_size=14, hidden_act="gelu_pytorch_tanh", layer_norm_eps=1e-6)
    return llm, exp, vis


def make_pi0_tiny(seed: int = 2468):
    llm, exp, vis = tiny_pi0_configs()
    T, A = 10, 32
    drop = ("rms_norm_eps", "rope_theta", "hidden_act", "layer_norm_eps")
    model = ref_loader.build_reference_pi0({k: v for k, v in llm.items() if k not in drop},
                                           {k: v for k, v in exp.items() if k not in drop},
                                           {k: v for k, v in vis.items() if k not in drop}, A, T)
    sd = seeded_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed)
    model.load_state_dict(sd, strict=True)
    model.train()
    g = torch.Generator().manual_seed(seed)
    B, L = 3, 12
    ids = torch.randint(1, 128, (B, L), generator=g)
    mask = torch.ones(B, L, dtype=torch.bool)
    mask[1, 8:] = False
    mask[2, 5:] = False
    images = torch.randn(B, 3, 3, 28, 28, generator=g)
    image_masks = torch.ones(B, 3, dtype=torch.bool)
    image_masks[1, 2] = False
    actions = torch.randn(B, T, A, generator=g)
    states = torch.randn(B, A, generator=g)
    # reference RNG order: torch.normal (pi0_arch.py:338-344) then Beta(1.5, 1).sample (:345-351)
    torch.manual_seed(seed + 1)
    out = model(input_ids=ids, attention_mask=mask, images=images, image_masks=image_masks, actions=actions,
                states=states)
    out.loss.backward()
    torch.manual_seed(seed + 1)
    noise = torch.normal(mean=torch.zeros_like(actions), std=torch.ones_like(actions))
    time = torch.distributions.Beta(1.5, 1).sample((B,)) * 0.999 + 0.001
    cfg = dict(llm=llm, expert=exp, vision=vis, chunk_size=T, action_dim=A)
    ora = vla_oracle.pi0_forward(sd, cfg, ids, mask, images, image_masks, actions, states, noise, time)
    d_loss = abs(ora["loss"].item() - out.loss.item())
    d_v = (ora["v_t"] - out.logits).abs().max().item()
    print(f"[pi0_tiny] reference loss {out.loss.item():.8f} oracle {ora['loss'].item():.8f}; v_t max|d|={d_v:.2e}")
    assert d_loss < 1e-5 and d_v < 1e-4
    names = ["model.llm.layers.0.self_attn.q_proj.weight", "model.llm.layers.1.mlp.gate_proj.weight",
             "model.action_expert.layers.1.mlp.down_proj.weight", "model.action_expert.layers.0.self_attn.k_proj.weight",
             "model.action_in_proj.weight", "model.action_time_mlp_in.weight", "model.action_out_proj.bias",
             "model.state_proj.weight", "model.mm_projector.weight", "model.llm.embed_tokens.weight",
             "model.mm_vision_tower.vision_tower.vision_model.encoder.layers.0.mlp.fc1.weight",
             "model.mm_vision_tower.vision_tower.vision_model.embeddings.patch_embedding.weight",
             "model.action_expert.norm.weight"]
    params = dict(model.named_parameters())
    grads = {n: params[n].grad.clone() for n in names if params[n].grad is not None}
    none_grad = sorted(n for n, p in params.items() if p.grad is None)
    torch.save(dict(seed=seed, cfg=cfg, shapes={k: tuple(v.shape) for k, v in sd.items()},
                    inputs=dict(input_ids=ids, attention_mask=mask, images=images, image_masks=image_masks,
                                actions=actions, states=states, noise=noise, time=time),
                    outputs=dict(loss=out.loss.detach(), v_t=out.logits.detach(), grads=grads, none_grad=none_grad)),
               GOLDEN / "pi0_tiny.pt")
    print("[pi0_tiny] wrote fixture; params without grad:", none_grad[:6], len(none_grad))


def make_pi0_inference_tiny(seed: int = 2468):
    """Pi0 inference_action (KV-cached prefix + Euler steps, pi0_arch.py:402-491) from the reference."""
    llm, exp, vis = tiny_pi0_configs()
    T, A = 10, 32
    drop = ("rms_norm_eps", "rope_theta", "hidden_act", "layer_norm_eps")
    model = ref_loader.build_reference_pi0({k: v for k, v in llm.items() if k not in drop},
                                           {k: v for k, v in exp.items() if k not in drop},
                                           {k: v for k, v in vis.items() if k not in drop}, A, T)
    sd = seeded_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed)
    model.load_state_dict(sd, strict=True)
    model.eval()
    g = torch.Generator().manual_seed(seed + 3)
    B, L = 3, 12
    ids = torch.randint(1, 128, (B, L), generator=g)
    mask = torch.ones(B, L, dtype=torch.bool)
    mask[0, 9:] = False
    mask[2, 4:] = False
    images = torch.randn(B, 3, 3, 28, 28, generator=g)
    image_masks = torch.ones(B, 3, dtype=torch.bool)
    image_masks[2, 1] = False
    states = torch.randn(B, A, generator=g)
    cfg = dict(llm=llm, expert=exp, vision=vis, chunk_size=T, action_dim=A)
    outs = {}
    for steps in (10, 4):
        torch.manual_seed(seed + 5)
        ref = model.inference_action(input_ids=ids, attention_mask=mask, states=states, images=images,
                                     image_masks=image_masks, diffusion_steps=steps)
        torch.manual_seed(seed + 5)
        noise = torch.normal(0, 1, size=(B, T, A))                     # pi0_arch.py:417-422
        ora = vla_oracle.pi0_inference(sd, cfg, ids, mask, images, image_masks, states, noise, steps)
        d = (ora - ref).abs().max().item()
        print(f"[pi0_inference] steps={steps}: oracle vs reference max|d|={d:.2e} (|x|max {ref.abs().max():.3f})")
        assert d < 1e-4
        outs[steps] = dict(noise=noise, actions=ref.detach())
    torch.save(dict(seed=seed, cfg=cfg, shapes={k: tuple(v.shape) for k, v in sd.items()},
                    inputs=dict(input_ids=ids, attention_mask=mask, images=images, image_masks=image_masks,
                                states=states), outputs=outs), GOLDEN / "pi0_inference_tiny.pt")


MEMVLA_MEM = dict(dataloader_type="group", group_size=3, per_token_size=32, mem_length=2, retrieval_layers=2,
                  use_timestep_pe=True, fusion_type="gate", consolidate_type="tome", update_fused=True)


def make_memvla_tiny(seed: int = 1357):
    """MemVLA training forward/backward (memvla_arch.py:546-664) with dropout 0: two episodes (3 + 2 frames) in one
    `group` batch, mem_length 2 so the third frame of episode 0 triggers the token-merge consolidation."""
    llm, clip, cfg = tiny_cogact_configs()
    model = ref_loader.build_reference_memvla(llm, clip, "DiT-S", **MEMVLA_MEM)
    sd = seeded_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed)
    model.load_state_dict(sd, strict=True)
    model.train()
    g = torch.Generator().manual_seed(seed)
    B, L, R = 5, 11, 4
    ids = torch.randint(1, 128, (B, L), generator=g)
    ids[:, 1] = vla_oracle.IMAGE_TOKEN_INDEX
    mask = torch.ones(B, L, dtype=torch.bool)
    mask[1, 9:] = False
    mask[3, 7:] = False
    ids[~mask] = 0
    images = torch.randn(B, 3, 28, 28, generator=g)
    actions = torch.rand(B, 16 * 7, generator=g) * 2 - 1
    indexes = [(0, 4, 0), (0, 4, 1), (0, 4, 2), (0, 9, 5), (0, 9, 6)]
    # reference RNG order inside ActionModel.loss: randn_like(x), randint (action_models.py:76-79), then token_drop's
    # torch.rand (dit.py:86-88)
    torch.manual_seed(seed + 1)
    out = model(input_ids=ids, attention_mask=mask, images=images, actions=actions, indexes=indexes, labels=None)
    out.loss.backward()
    torch.manual_seed(seed + 1)
    noise = torch.randn(R * B, 16, 7)
    timesteps = torch.randint(0, 100, (R * B,))
    drop = torch.rand(R * B) < 0.1
    mcfg = dict(cfg, mem=MEMVLA_MEM)
    ora = vla_oracle.memvla_forward(sd, mcfg, ids, mask, images, actions, indexes, noise, timesteps, drop, R)
    d = abs(ora["loss"].item() - out.loss.item())
    print(f"[memvla_tiny] reference loss {out.loss.item():.8f} oracle {ora['loss'].item():.8f}")
    assert d < 1e-5, d
    # intermediate pins straight from the reference modules
    with torch.no_grad():
        per_ref = model.model.per_compr(out.vision_proj_feats)
    d_per = (per_ref - ora["per_tokens"]).abs().max().item()
    print(f"[memvla_tiny] per_compr max|d| {d_per:.2e}")
    assert d_per < 1e-4
    names = ["model.per_compr.excite.1.weight", "model.per_compr.excite.3.bias", "model.per_compr.reduce.0.weight",
             "model.per_compr.reduce.2.weight",
             "model.per_cog_mem_bank.retrieval_blocks.cog.0.q_proj.weight",
             "model.per_cog_mem_bank.retrieval_blocks.cog.1.k_proj.weight",
             "model.per_cog_mem_bank.retrieval_blocks.cog.1.ffn.0.weight",
             "model.per_cog_mem_bank.retrieval_blocks.per.0.v_proj.weight",
             "model.per_cog_mem_bank.retrieval_blocks.per.1.ffn.3.weight",
             "model.per_cog_mem_bank.retrieval_blocks.per.1.ffn_norm.weight",
             "model.per_cog_mem_bank.gate_fusion_blocks.cog.proj.weight",
             "model.per_cog_mem_bank.gate_fusion_blocks.per.proj.bias",
             "model.per_cog_mem_bank.timestep_embedders.cog.mlp.0.weight",
             "model.per_cog_mem_bank.timestep_embedders.per.mlp.2.weight",
             "model.action_head.net.per_token_embedder.linear.weight",
             "model.action_head.net.blocks.0.per_attn.in_proj_bias",
             "model.action_head.net.blocks.5.per_attn.out_proj.weight",
             "model.action_head.net.blocks.3.norm3.weight",
             "model.action_head.net.blocks.2.mlp.fc1.bias",
             "model.action_head.net.z_embedder.linear.weight",
             "model.mm_projector.0.weight", "model.mm_projector.2.weight",
             "model.llm.layers.1.mlp.down_proj.weight", "model.llm.layers.0.self_attn.q_proj.weight",
             "model.llm.embed_tokens.weight"]
    params = dict(model.named_parameters())
    grads = {n: params[n].grad.clone() for n in names if params[n].grad is not None}
    missing = [n for n in names if params[n].grad is None]
    # oracle gradients must agree with the reference's
    osd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    ora2 = vla_oracle.memvla_forward(osd, mcfg, ids, mask, images, actions, indexes, noise, timesteps, drop, R)
    ora2["loss"].backward()
    worst = max(((osd[n].grad - g_).norm() / (g_.norm() + 1e-12)).item() for n, g_ in grads.items())
    print(f"[memvla_tiny] oracle-vs-reference worst relative grad error {worst:.2e}; params without grad: {missing}")
    assert worst < 2e-3
    none_grad = sorted(n for n, p in params.items() if p.grad is None)
    torch.save(dict(seed=seed, cfg=mcfg, shapes={k: tuple(v.shape) for k, v in sd.items()},
                    inputs=dict(input_ids=ids, attention_mask=mask, images=images, actions=actions, indexes=indexes,
                                noise=noise, timesteps=timesteps, drop_mask=drop),
                    outputs=dict(loss=out.loss.detach(), per_tokens=per_ref, cog_fused=ora["cog_fused"].detach(),
                                 per_fused=ora["per_fused"].detach(), noise_pred=ora["noise_pred"].detach(),
                                 grads=grads, none_grad=none_grad)),
               GOLDEN / "memvla_tiny.pt")
    print("[memvla_tiny] wrote fixture;", len(none_grad), "params without grad")


def make_pi05_tiny(seed: int = 9753):
    """pi0.5 training forward/backward and inference_action from the reference (pi05_arch.py)."""
    llm, exp, vis = tiny_pi0_configs()
    exp = dict(exp, use_adarms=True, adarms_cond_dim=exp["hidden_size"], width=exp["hidden_size"])
    T, A = 10, 32
    drop = ("rms_norm_eps", "hidden_act", "layer_norm_eps", "model_type")
    model = ref_loader.build_reference_pi05({k: v for k, v in llm.items() if k not in drop},
                                            {k: v for k, v in exp.items() if k not in drop},
                                            {k: v for k, v in vis.items() if k not in ("rms_norm_eps", "rope_theta",
                                                                                        "hidden_act", "layer_norm_eps")},
                                            A, T)
    sd = seeded_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed)
    model.load_state_dict(sd, strict=True)
    model.train()
    g = torch.Generator().manual_seed(seed)
    B, L = 3, 12
    ids = torch.randint(1, 128, (B, L), generator=g)
    mask = torch.ones(B, L, dtype=torch.bool)
    mask[1, 8:] = False
    mask[2, 5:] = False
    images = torch.randn(B, 3, 3, 28, 28, generator=g)
    image_masks = torch.ones(B, 3, dtype=torch.bool)
    image_masks[1, 2] = False
    actions = torch.randn(B, T, A, generator=g)
    torch.manual_seed(seed + 1)
    out = model(input_ids=ids, attention_mask=mask, images=images, image_masks=image_masks, actions=actions)
    out.loss.backward()
    torch.manual_seed(seed + 1)
    noise = torch.normal(mean=torch.zeros_like(actions), std=torch.ones_like(actions))       # pi05_arch.py:352-358
    time = torch.distributions.Beta(1.5, 1).sample((B,)) * 0.999 + 0.001
    cfg = dict(llm=llm, expert=exp, vision=vis, chunk_size=T, action_dim=A)
    ora = vla_oracle.pi05_forward(sd, cfg, ids, mask, images, image_masks, actions, noise, time)
    d_loss = abs(ora["loss"].item() - out.loss.item())
    d_v = (ora["v_t"] - out.logits).abs().max().item()
    print(f"[pi05_tiny] reference loss {out.loss.item():.8f} oracle {ora['loss'].item():.8f}; v_t max|d|={d_v:.2e}")
    assert d_loss < 1e-5 and d_v < 1e-4
    names = ["model.llm.layers.0.self_attn.q_proj.weight", "model.llm.layers.1.self_attn.k_proj.weight",
             "model.action_expert.layers.1.mlp.down_proj.weight", "model.action_expert.layers.0.self_attn.k_proj.weight",
             "model.action_expert.layers.0.input_layernorm.dense.weight",
             "model.action_expert.layers.1.post_attention_layernorm.dense.bias",
             "model.action_expert.norm.dense.weight", "model.time_mlp_in.weight", "model.time_mlp_out.bias",
             "model.action_in_proj.weight", "model.action_out_proj.bias", "model.mm_projector.weight",
             "model.llm.embed_tokens.weight",
             "model.mm_vision_tower.vision_tower.vision_model.encoder.layers.0.mlp.fc1.weight"]
    params = dict(model.named_parameters())
    grads = {n: params[n].grad.clone() for n in names if params[n].grad is not None}
    none_grad = sorted(n for n, p in params.items() if p.grad is None)
    # inference
    model.eval()
    states = torch.randn(B, A, generator=g)
    inf = {}
    for steps in (10, 4):
        torch.manual_seed(seed + 5)
        with torch.no_grad():
            ref = model.inference_action(input_ids=ids, attention_mask=mask, states=states, images=images,
                                         image_masks=image_masks, diffusion_steps=steps)
        torch.manual_seed(seed + 5)
        n0 = torch.normal(0, 1, size=(B, T, A))
        got = vla_oracle.pi05_inference(sd, cfg, ids, mask, images, image_masks, n0, steps)
        d = (got - ref).abs().max().item()
        print(f"[pi05_inference] steps={steps}: oracle vs reference max|d|={d:.2e}")
        assert d < 1e-4
        inf[steps] = dict(noise=n0, actions=ref.detach())
    torch.save(dict(seed=seed, cfg=cfg, shapes={k: tuple(v.shape) for k, v in sd.items()},
                    inputs=dict(input_ids=ids, attention_mask=mask, images=images, image_masks=image_masks,
                                actions=actions, noise=noise, time=time, states=states),
                    outputs=dict(loss=out.loss.detach(), v_t=out.logits.detach(), grads=grads, none_grad=none_grad,
                                 inference=inf)),
               GOLDEN / "pi05_tiny.pt")
    print("[pi05_tiny] wrote fixture; params without grad:", len(none_grad), [n for n in names if n not in grads])


def make_memvla_inference_tiny(seed: int = 1357):
    """MemVLA inference_action over four consecutive frames of one episode (memory grows, then token-merges at
    mem_length 2), dropout 0: reference vs oracle, per-frame noise and outputs stored."""
    llm, clip, cfg = tiny_cogact_configs()
    model = ref_loader.build_reference_memvla(llm, clip, "DiT-S", **MEMVLA_MEM)
    sd = seeded_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed)
    model.load_state_dict(sd, strict=True)
    model.eval()
    mcfg = dict(cfg, mem=MEMVLA_MEM)
    m = MEMVLA_MEM
    banks = {r: vla_oracle.MemBankOracle(sd, "model.per_cog_mem_bank.", r, m["mem_length"], m["retrieval_layers"],
                                         m["dataloader_type"], m["use_timestep_pe"], m["fusion_type"],
                                         m["consolidate_type"], m["update_fused"]) for r in ("per", "cog")}
    g = torch.Generator().manual_seed(seed + 11)
    L = 11
    norms = {"min": [-1.0] * 7, "max": [1.0] * 7}
    frames = []
    for f in range(4):
        ids = torch.randint(1, 128, (1, L), generator=g)
        ids[:, 1] = vla_oracle.IMAGE_TOKEN_INDEX
        images = torch.randn(1, 3, 28, 28, generator=g)
        torch.manual_seed(seed + 20 + f)
        acts = model.inference_action(ids, images, "True" if f == 0 else "False",
                                      {"cfg_scale": 1.5, "num_ddim_steps": 10, "action_norms": norms})
        torch.manual_seed(seed + 20 + f)
        noise = torch.randn(1, 16, 7)                                # memvla_arch.py:704-709
        ora = vla_oracle.memvla_inference(sd, mcfg, banks, ids, images, noise, f, 1.5, 10)
        ref = torch.tensor(acts)
        d = (ora[0].clamp(-1, 1) - ref).abs().max().item()
        print(f"[memvla_inference] frame {f}: oracle vs reference max|d|={d:.2e}; bank length "
              f"{len(banks['cog'].banks[(0, 0)])}")
        assert d < 1e-4
        frames.append(dict(input_ids=ids, images=images, noise=noise, actions=ref))
    assert len(banks["cog"].banks[(0, 0)]) == 2
    torch.save(dict(seed=seed, cfg=mcfg, shapes={k: tuple(v.shape) for k, v in sd.items()}, frames=frames),
               GOLDEN / "memvla_inference_tiny.pt")


def make_cogact_inference_tiny(seed: int = 1234):
    """CogACT inference_action (CFG 1.5, 10-step DDIM, eta=0) from the reference (cogact_arch.py:149-198)."""
    llm, clip, cfg = tiny_cogact_configs()
    model = ref_loader.build_reference_cogact(llm, clip, "DiT-S")
    sd = seeded_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed)
    model.load_state_dict(sd, strict=True)
    model.eval()
    g = torch.Generator().manual_seed(seed + 7)
    B, L = 2, 11
    ids = torch.randint(1, 128, (B, L), generator=g)
    ids[:, 1] = vla_oracle.IMAGE_TOKEN_INDEX
    images = torch.randn(B, 3, 28, 28, generator=g)
    norms = {"min": [-1.0] * 7, "max": [1.0] * 7}
    outs = {}
    for scale in (1.5, 1.0):
        model.model.action_head.ddim_diffusion = None
        torch.manual_seed(seed + 9)
        acts = model.inference_action(ids, images, {"cfg_scale": scale, "num_ddim_steps": 10, "action_norms": norms})
        torch.manual_seed(seed + 9)
        noise = torch.randn(B, 16, 7)                      # cogact_arch.py:161-166
        ora = vla_oracle.cogact_inference(sd, cfg, ids, images, noise, scale, 10)
        ref = torch.tensor(acts)
        d = (ora[0].clamp(-1, 1) - ref).abs().max().item()
        print(f"[cogact_inference] cfg_scale={scale}: oracle vs reference max|d|={d:.2e}")
        assert d < 1e-4
        outs[scale] = dict(noise=noise, actions_sample0=ref, samples=ora.detach())
    torch.save(dict(seed=seed, cfg=cfg, shapes={k: tuple(v.shape) for k, v in sd.items()},
                    inputs=dict(input_ids=ids, images=images), outputs=outs), GOLDEN / "cogact_inference_tiny.pt")


def make_hybrid_cogact_tiny(seed: int = 4321):
    """HybridCogACTForCausalLM (hybrid_cogact_arch.py:60-218): mixed batch (some rows text-only, some action-only, some
    both) and an action-only batch (text loss is NaN in the reference: see the oracle's docstring)."""
    llm, clip, cfg = tiny_cogact_configs()
    model = ref_loader.build_reference_hybrid_cogact(llm, clip, "DiT-S")
    sd = seeded_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed)
    model.load_state_dict(sd, strict=True)
    model.train()
    g = torch.Generator().manual_seed(seed)
    B, L, R = 4, 14, 4
    ids = torch.randint(1, 128, (B, L), generator=g)
    ids[:, 1] = vla_oracle.IMAGE_TOKEN_INDEX
    mask = torch.ones(B, L, dtype=torch.long)
    mask[1, 10:] = 0
    images = torch.randn(B, 3, 28, 28, generator=g)
    actions = torch.rand(B, 16 * 7, generator=g) * 2 - 1
    labels = ids.clone()
    labels[:, :7] = vla_oracle.IGNORE_INDEX
    labels[mask == 0] = vla_oracle.IGNORE_INDEX
    cases = {}
    for name, has_action, has_text in (("mixed", [1, 0, 1, 1], [1, 1, 0, 1]), ("no_text", [1, 1, 0, 1], [0, 0, 0, 0])):
        ha, ht = torch.tensor(has_action).view(B, 1), torch.tensor(has_text).view(B, 1)
        lab = labels.clone()
        lab[ht.view(-1) == 0] = vla_oracle.IGNORE_INDEX            # the collator gives text-less rows no targets
        model.zero_grad()
        torch.manual_seed(seed + 1)
        out = model(input_ids=ids, attention_mask=mask, images=images, actions=actions, labels=lab, has_action=ha,
                    has_text=ht, repeated_diffusion_steps=R)
        torch.manual_seed(seed + 1)
        noise = torch.randn(R * B, 16, 7)
        timesteps = torch.randint(0, 100, (R * B,))
        drop = torch.rand(R * B) < 0.1
        sd_g = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
        ora = vla_oracle.hybrid_cogact_forward(sd_g, cfg, ids, mask, images, actions, lab, ha, ht, noise, timesteps,
                                               drop, R)
        d_act = abs(ora["action_loss"].item() - out.action_loss.item())
        same_text = (torch.isnan(out.text_loss) and torch.isnan(ora["text_loss"])) or \
            abs(ora["text_loss"].item() - out.text_loss.item()) < 1e-5
        print(f"[hybrid_cogact/{name}] reference text {out.text_loss.item():.6f} action {out.action_loss.item():.6f}; "
              f"oracle text {ora['text_loss'].item():.6f} action {ora['action_loss'].item():.6f}")
        assert d_act < 1e-5 and same_text
        grads = {}
        if name == "mixed":
            out.loss.backward()
            ora["loss"].backward()
            params = dict(model.named_parameters())
            for n in ["lm_head.weight", "model.llm.layers.0.self_attn.q_proj.weight", "model.llm.layers.1.mlp.down_proj.weight",
                      "model.llm.embed_tokens.weight", "model.mm_projector.0.weight",
                      "model.action_head.net.final_layer.linear.weight", "model.action_head.net.blocks.0.attn.qkv.weight"]:
                gr = params[n].grad
                rel = ((sd_g[n].grad - gr).norm() / gr.norm()).item()
                assert rel < 1e-4, (n, rel)
                grads[n] = gr.detach().clone()
        cases[name] = dict(inputs=dict(input_ids=ids, attention_mask=mask, images=images, actions=actions, labels=lab,
                                       has_action=ha, has_text=ht, noise=noise, timesteps=timesteps, drop_mask=drop,
                                       repeated_diffusion_steps=R),
                           outputs=dict(text_loss=out.text_loss.detach(), action_loss=out.action_loss.detach(),
                                        loss=out.loss.detach(), grads=grads))
    torch.save(dict(seed=seed, cfg=cfg, shapes={k: tuple(v.shape) for k, v in sd.items()}, cases=cases),
               GOLDEN / "hybrid_cogact_tiny.pt")
    print(f"[hybrid_cogact_tiny] wrote {GOLDEN / 'hybrid_cogact_tiny.pt'}")


def make_image_preprocess(seed: int = 97):
    """PreprocessRGB (rgb_preprocess.py:13-44, image_aspect_ratio='pad', both pad modes) with the PIL-backed HF CLIP image
    processor (what `CLIPImageProcessor` is under the reference's pinned transformers 4.5x; 5.5 renamed it
    CLIPImageProcessorPil and made a torchvision backend the default), and ActionNorm (action.py:229-275)."""
    import numpy as np
    from PIL import Image
    from transformers import CLIPImageProcessorPil
    from oracle import image_oracle
    ref_loader.load_reference()
    from dexbotic.data.dataset.rgb_preprocess import PreprocessRGB
    from dexbotic.data.dataset.transform.action import ActionNorm
    rng = np.random.default_rng(seed)
    out = {}
    for size in (32, 224):
        proc = CLIPImageProcessorPil(size={"shortest_edge": size}, crop_size={"height": size, "width": size})
        for mode in ("mean", "zero"):
            pp = PreprocessRGB(proc, image_aspect_ratio="pad", image_pad_mode=mode)
            for (H, W) in ((60, 80), (90, 50), (33, 33), (240, 320)):
                if size == 224 and H < 200:
                    continue
                img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
                img[: H // 3] = np.linspace(0, 255, W)[None, :, None].astype(np.uint8)     # ramps + noise: ringing, clipping
                ref = pp(Image.fromarray(img)).numpy()
                ora, u8 = image_oracle.preprocess_rgb(img, size, proc.image_mean, proc.image_std, proc.rescale_factor, mode)
                sq = image_oracle.expand2square(img, (0, 0, 0) if mode == "zero" else tuple(int(x * 255) for x in proc.image_mean))
                pil = np.asarray(Image.fromarray(sq).resize((size, size), resample=Image.BICUBIC))
                assert np.array_equal(ref, ora) and np.array_equal(pil, u8), (size, mode, H, W)
                key = f"s{size}_{mode}_{H}x{W}"
                out[key + "_img"], out[key + "_ref"], out[key + "_u8"] = img, ref, pil
    out["image_mean"], out["image_std"] = np.array(proc.image_mean), np.array(proc.image_std)
    # action normalisation, both modes, through the reference transform
    a = rng.normal(size=(17, 7)) * 3
    stats = dict(min=rng.normal(size=7) - 4, max=rng.normal(size=7) + 4, mean=rng.normal(size=7),
                 std=np.abs(rng.normal(size=7)) + 0.1)          # numpy arrays: the transform subtracts them directly
    for q in (True, False):
        ref = ActionNorm({"action": stats}, use_quantiles=q)({"action": a.copy()})["action"]
        assert np.array_equal(ref, image_oracle.action_normalize(a, stats, q)) and ref.dtype == np.float32
        out[f"action_{'quantile' if q else 'meanstd'}"] = ref
    out["action_in"] = a
    for k, v in stats.items():
        out["stat_" + k] = np.array(v)
    np.savez_compressed(GOLDEN / "image_preprocess.npz", **out)
    print(f"[image_preprocess] {len(out)} arrays -> {GOLDEN / 'image_preprocess.npz'}; oracle == reference == Pillow")


def tiny_navila_configs():
    from transformers import SiglipVisionConfig
    llm = dict(model_type="llama", vocab_size=160, hidden_size=64, intermediate_size=160, num_hidden_layers=2,
               num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=256, rope_theta=10000.0,
               rms_norm_eps=1e-6, hidden_act="silu")
    # 70 / 14 = 5 x 5 patches: an ODD grid, so the 2x2 down-sampling pads it to 6 x 6 -> 9 tokens of 4 x 32 channels
    vis_kw = dict(hidden_size=32, intermediate_size=64, num_hidden_layers=3, num_attention_heads=2, image_size=70,
                  patch_size=14, hidden_act="gelu_pytorch_tanh", layer_norm_eps=1e-6)
    return llm, SiglipVisionConfig(**vis_kw), dict(vis_kw, model_type="siglip_vision_model")


def make_navila_tiny(seed: int = 1357):
    """NaVILAForCausalLM training forward (navila_arch.py:362-497): plain shifted CE and the soft cross entropy over the
    time tokens (navila/loss.py), single-image rows and a 2-frame 5-D batch."""
    ref_loader.load_reference_navila()
    from dexbotic.model.modules.mm_projector.builder import DownSampleBlock
    llm, vis_cfg, vis = tiny_navila_configs()
    x = torch.randn(2, 25, 6)
    assert torch.equal(DownSampleBlock()(x), vla_oracle.downsample_2x2(x))
    time_tokens = [150, 151, 152, 153, 154, 155]
    cfg = dict(llm=llm, vision=vis, tokenizer_model_max_length=None, tokenizer_padding_side="right")
    cases = {}
    for name, soft, frames in (("ce", None, 1), ("soft_ce", time_tokens, 1), ("ce_two_frames", None, 2)):
        model = ref_loader.build_reference_navila(llm, vis_cfg, time_token_ids=soft, soft_ce_std=1.5)
        sd = seeded_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed)
        model.load_state_dict(sd, strict=True)
        model.train()
        g = torch.Generator().manual_seed(seed + len(name))
        B, L = 3, 16
        ids = torch.randint(1, 150, (B, L), generator=g)
        ids[:, 2] = vla_oracle.IMAGE_TOKEN_INDEX
        if frames == 2:
            ids[:, 5] = vla_oracle.IMAGE_TOKEN_INDEX
        mask = torch.ones(B, L, dtype=torch.long)
        mask[1, 12:] = 0
        labels = ids.clone()
        labels[:, :8] = vla_oracle.IGNORE_INDEX                      # the prompt is not a target
        labels[mask == 0] = vla_oracle.IGNORE_INDEX
        if soft:                                                     # some targets are time tokens, incl. both ends
            labels[0, 9], labels[0, 12], labels[2, 10], labels[1, 9] = 150, 155, 152, 153
            ids[0, 9], ids[0, 12], ids[2, 10], ids[1, 9] = 150, 155, 152, 153
        images = torch.randn(B, 3, 70, 70, generator=g) if frames == 1 else torch.randn(B, 2, 3, 70, 70, generator=g)
        out = model(input_ids=ids, attention_mask=mask, images=images, labels=labels)
        out.loss.backward()
        sd_g = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
        ora = vla_oracle.navila_forward(sd_g, cfg, ids, mask, images, labels, time_token_ids=soft, soft_ce_std=1.5)
        ora["loss"].backward()
        d_loss = abs(ora["loss"].item() - out.loss.item())
        valid = ora["attention_mask"][:, :, None]
        d_log = ((ora["logits"] - out.logits) * valid).abs().max().item()
        names = ["model.llm.layers.0.self_attn.q_proj.weight", "model.llm.layers.1.mlp.down_proj.weight",
                 "model.llm.embed_tokens.weight", "model.llm.norm.weight", "lm_head.weight",
                 "model.mm_projector.1.weight", "model.mm_projector.2.weight", "model.mm_projector.4.bias",
                 "model.mm_vision_tower.vision_tower.vision_model.encoder.layers.0.self_attn.k_proj.weight",
                 "model.mm_vision_tower.vision_tower.vision_model.encoder.layers.1.mlp.fc2.weight",
                 "model.mm_vision_tower.vision_tower.vision_model.embeddings.patch_embedding.weight",
                 "model.mm_vision_tower.vision_tower.vision_model.embeddings.position_embedding.weight"]
        params = dict(model.named_parameters())
        grads, worst = {}, 0.0
        for n in names:
            gr = params[n].grad
            grads[n] = gr.detach().clone()
            worst = max(worst, ((sd_g[n].grad - gr).norm() / gr.norm()).item())
        none_grad = sorted(n for n, p_ in params.items() if p_.grad is None or p_.grad.abs().max() == 0)
        print(f"[navila_tiny/{name}] reference loss {out.loss.item():.8f} oracle {ora['loss'].item():.8f} "
              f"|d|={d_loss:.2e}; logits max|d|={d_log:.2e}; grads rel {worst:.2e}; {len(none_grad)} params without grad")
        assert d_loss < 1e-5 and d_log < 2e-4 and worst < 1e-4
        cases[name] = dict(inputs=dict(input_ids=ids, attention_mask=mask, images=images, labels=labels),
                           time_token_ids=soft, soft_ce_std=1.5,
                           outputs=dict(loss=out.loss.detach(), logits=(out.logits * valid).detach(),
                                        valid=ora["attention_mask"], grads=grads, none_grad=none_grad))
    torch.save(dict(seed=seed, cfg=cfg, shapes={k: tuple(v.shape) for k, v in sd.items()}, cases=cases),
               GOLDEN / "navila_tiny.pt")
    print(f"[navila_tiny] wrote {GOLDEN / 'navila_tiny.pt'}")


if __name__ == "__main__":
    if len(sys.argv) > 1:                      # python oracle/make_golden.py make_navila_tiny
        for fn in sys.argv[1:]:
            globals()[fn]()
        sys.exit(0)
    make_navila_tiny()
    make_hybrid_cogact_tiny()
    make_image_preprocess()
    make_cogact_tiny()
    make_cogact_inference_tiny()
    make_pi0_tiny()
    make_pi0_inference_tiny()
    make_pi05_tiny()
    make_memvla_tiny()
    make_memvla_inference_tiny()
    make_oft_discrete_tiny()
    make_oft_discrete_proprio_tiny()
    make_oft_linear_tiny()
    make_oft_diffusion_tiny()
    make_splice_cases()
    make_integer_kats()
