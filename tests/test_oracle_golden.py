"""CPU tests: the oracle restatement (oracle/vla_oracle.py) against the committed golden vectors that
oracle/make_golden.py produced from the UNMODIFIED reference (tests/golden/*.pt)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import vla_oracle
from oracle.weights import seeded_state_dict

GOLDEN = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def cogact_tiny():
    return torch.load(GOLDEN / "cogact_tiny.pt", weights_only=False)


def test_cogact_tiny_forward_matches_reference(cogact_tiny):
    fx = cogact_tiny
    sd = seeded_state_dict(fx["shapes"], fx["seed"])
    i = fx["inputs"]
    out = vla_oracle.cogact_forward(sd, fx["cfg"], i["input_ids"], i["attention_mask"], i["images"], i["actions"],
                                    i["noise"], i["timesteps"], i["drop_mask"], i["repeated_diffusion_steps"])
    assert abs(out["loss"].item() - fx["outputs"]["loss"].item()) < 1e-5
    valid = fx["outputs"]["valid"][:, :, None]
    assert ((out["last_hidden"] * valid) - fx["outputs"]["last_hidden"]).abs().max().item() < 1e-4
    assert (out["cognition"] - fx["outputs"]["cognition"]).abs().max().item() < 1e-4


def test_cogact_tiny_gradients_match_reference(cogact_tiny):
    fx = cogact_tiny
    sd = seeded_state_dict(fx["shapes"], fx["seed"])
    sd = {k: (v.requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    i = fx["inputs"]
    out = vla_oracle.cogact_forward(sd, fx["cfg"], i["input_ids"], i["attention_mask"], i["images"], i["actions"],
                                    i["noise"], i["timesteps"], i["drop_mask"], i["repeated_diffusion_steps"])
    out["loss"].backward()
    for name, gref in fx["outputs"]["grads"].items():
        d = (sd[name].grad - gref).abs().max().item()
        assert d < 1e-5 + 1e-3 * gref.abs().max().item(), name


def test_splice_cases_bit_exact():
    for c in torch.load(GOLDEN / "splice_cases.pt", weights_only=False):
        emb, lab, msk, pid = vla_oracle.splice(c["embed_weight"], c["image_features"], c["input_ids"],
                                               c["attention_mask"], c["labels"], c["max_len"], c["padding_side"])
        assert torch.equal(emb, c["inputs_embeds"]), c["name"]
        assert torch.equal(lab, c["new_labels"]), c["name"]
        assert torch.equal(msk, c["new_mask"]), c["name"]


def test_oft_integer_kats_bit_exact():
    k = torch.load(GOLDEN / "oft_integer_kats.pt", weights_only=False)
    assert np.array_equal(vla_oracle.oft_discretize(k["actions"].numpy()), k["bins"].numpy())
    assert np.array_equal(vla_oracle.oft_bins_to_continuous(k["bins"].numpy()), k["continuous"].numpy())
    assert np.array_equal(vla_oracle.oft_argmax_decode(k["logits"].numpy()), k["argmax"].numpy())
    assert np.array_equal(vla_oracle.data_action_to_bin(k["data_actions"].numpy(), 255), k["data_bins"].numpy())
    # ties: half-way points round to even, +-1 map to the end bins
    assert vla_oracle.oft_discretize(np.array([-1.0, 1.0, 0.0], np.float32)).tolist() == [0, 255, 128]


def test_cosine_schedule_matches_reference():
    k = torch.load(GOLDEN / "cosine_schedule_T100.pt", weights_only=False)
    sa, sb = vla_oracle.cosine_schedule(100)
    assert np.array_equal(sa, k["sqrt_alphas_cumprod"].numpy())
    assert np.array_equal(sb, k["sqrt_one_minus_alphas_cumprod"].numpy())
    assert abs(sa[0] - 0.999684309) < 1e-9 and abs(sa[50] - 0.691566796) < 1e-9 and abs(sa[99] - 4.92805467e-4) < 1e-12


def test_oft_discrete_tiny_matches_reference():
    fx = torch.load(GOLDEN / "oft_discrete_tiny.pt", weights_only=False)
    sd = seeded_state_dict(fx["shapes"], fx["seed"])
    i = fx["inputs"]
    out = vla_oracle.oft_discrete_forward(sd, fx["cfg"], i["input_ids"], i["attention_mask"], i["images"], i["labels"])
    assert abs(out["loss"].item() - fx["outputs"]["loss"].item()) < 1e-5
    assert (out["logits"] - fx["outputs"]["logits"]).abs().max().item() < 1e-4
    # integer decode of the reference's own inference logits (oft_discrete_arch.py:222-224; model.py:314-347)
    idx = vla_oracle.oft_argmax_decode(fx["outputs"]["infer_logits"].numpy())
    assert np.array_equal(idx, fx["outputs"]["infer_idx"].numpy())
    cont = vla_oracle.oft_bins_to_continuous(idx).reshape(idx.shape[0], 8, 7)
    assert np.array_equal(cont, fx["outputs"]["infer_cont"].numpy())


def test_pi0_tiny_matches_reference():
    fx = torch.load(GOLDEN / "pi0_tiny.pt", weights_only=False)
    sd = seeded_state_dict(fx["shapes"], fx["seed"])
    i = fx["inputs"]
    out = vla_oracle.pi0_forward(sd, fx["cfg"], i["input_ids"], i["attention_mask"], i["images"], i["image_masks"],
                                 i["actions"], i["states"], i["noise"], i["time"])
    assert abs(out["loss"].item() - fx["outputs"]["loss"].item()) < 1e-5
    assert (out["v_t"] - fx["outputs"]["v_t"]).abs().max().item() < 1e-4


def test_pi0_inference_matches_reference():
    """Oracle Euler sampler (cache-free restatement) vs the reference's KV-cached inference_action output."""
    fx = torch.load(GOLDEN / "pi0_inference_tiny.pt", weights_only=False)
    sd = seeded_state_dict(fx["shapes"], fx["seed"])
    i = fx["inputs"]
    for steps, ref in fx["outputs"].items():
        got = vla_oracle.pi0_inference(sd, fx["cfg"], i["input_ids"], i["attention_mask"], i["images"],
                                       i["image_masks"], i["states"], ref["noise"], steps)
        assert (got - ref["actions"]).abs().max().item() < 1e-4


def test_oft_linear_tiny_matches_reference():
    fx = torch.load(GOLDEN / "oft_linear_tiny.pt", weights_only=False)
    for use_proprio, case in fx["cases"].items():
        sd = seeded_state_dict(case["shapes"], fx["seed"])
        i, ref = case["inputs"], case["outputs"]
        out = vla_oracle.oft_l1_forward(sd, case["cfg"], i["input_ids"], i["attention_mask"], i["images"], i["actions"],
                                        i["states"])
        assert abs(out["loss"].item() - ref["loss"].item()) < 1e-5
        assert (out["predicted_actions"] - ref["predicted_actions"]).abs().max().item() < 1e-4


def test_oft_diffusion_tiny_matches_reference():
    """OFT `DiT` head: oracle forward (fixed noisy_dict) and DDIM inference vs the reference code run around the restated
    scheduler (oracle/ddim_oracle.py; the scheduler itself is parity-unpinned: diffusers is absent)."""
    fx = torch.load(GOLDEN / "oft_diffusion_tiny.pt", weights_only=False)
    for use_proprio, case in fx["cases"].items():
        sd = seeded_state_dict(case["shapes"], fx["seed"])
        i, ref = case["inputs"], case["outputs"]
        out = vla_oracle.oft_diffusion_forward(sd, case["cfg"], i["input_ids"], i["attention_mask"], i["images"],
                                               i["noisy_dict"], i["actions"], i["states"])
        assert abs(out["loss"].item() - ref["loss"].item()) < 1e-5
        assert (out["predicted_noise"] - ref["predicted_noise"]).abs().max().item() < 1e-4
        st = i["states"][:1] if use_proprio else None
        acts = vla_oracle.oft_diffusion_inference(sd, case["cfg"], i["input_ids"][:1], i["images"][:1], i["start_noise"],
                                                  i["num_ddim_steps"], st)
        assert (acts[0].clamp(-1, 1) - ref["inference_actions"]).abs().max().item() < 1e-4


def test_product_ddim_scheduler_matches_the_restated_one():
    """dexbotic_b200.model.ddim (host arithmetic, no CUDA) vs oracle/ddim_oracle.py: tables, add_noise, timesteps, step —
    and two closed-form properties of the DDIM update: with the true noise it returns sqrt(a_prev) x0 + sqrt(1-a_prev) eps,
    and the last step (prev_t < 0, alpha = 1) returns the clipped x0 estimate."""
    from dexbotic_b200.model.ddim import DDIMScheduler
    from oracle.ddim_oracle import DDIMSchedulerOracle
    a, b = DDIMSchedulerOracle(100, "squaredcos_cap_v2"), DDIMScheduler(100)
    assert (a.alphas_cumprod - b.alphas_cumprod).abs().max().item() < 1e-6
    assert 0.99 < float(b.alphas_cumprod[0]) < 1 and float(b.alphas_cumprod[-1]) < 1e-5
    g = torch.Generator().manual_seed(3)
    x0 = torch.rand(4, 8, 7, generator=g) * 2 - 1
    eps = torch.randn(4, 8, 7, generator=g)
    t = torch.tensor([0, 17, 50, 99])
    assert (a.add_noise(x0, eps, t) - b.add_noise(x0, eps, t)).abs().max().item() < 1e-6
    for n in (5, 10, 25, 100):
        a.set_timesteps(n)
        b.set_timesteps(n)
        assert a.timesteps.tolist() == b.timesteps.tolist() and len(b.timesteps) == n and b.timesteps[-1] == 0
        x = torch.randn(4, 8, 7, generator=g)
        for ts in b.timesteps.tolist():
            pa, pb = a.step(eps, ts, x).prev_sample, b.step(eps, ts, x).prev_sample
            assert (pa - pb).abs().max().item() < 2e-5
            x = pb
    b.set_timesteps(10)
    ts = 50
    xt = b.add_noise(x0, eps, torch.full((4,), ts))
    prev = b.step(eps, ts, xt).prev_sample
    ap = float(b.alphas_cumprod[ts - 10])
    assert (prev - (ap ** 0.5 * x0 + (1 - ap) ** 0.5 * eps)).abs().max().item() < 1e-4
    last = b.step(eps, 0, b.add_noise(x0, eps, torch.zeros(4, dtype=torch.long)))
    assert (last.prev_sample - x0).abs().max().item() < 1e-4


def test_memvla_tiny_matches_reference():
    """MemVLA oracle (BottleneckSE, memory bank with token-merge consolidation, DiT per_attn) vs the reference."""
    fx = torch.load(GOLDEN / "memvla_tiny.pt", weights_only=False)
    sd = seeded_state_dict(fx["shapes"], fx["seed"])
    i, ref = fx["inputs"], fx["outputs"]
    out = vla_oracle.memvla_forward(sd, fx["cfg"], i["input_ids"], i["attention_mask"], i["images"], i["actions"],
                                    i["indexes"], i["noise"], i["timesteps"], i["drop_mask"])
    assert abs(out["loss"].item() - ref["loss"].item()) < 1e-5
    assert (out["per_tokens"] - ref["per_tokens"]).abs().max().item() < 1e-4
    # episode 0 holds 3 frames with mem_length 2: the token merge must have fired (bank length stays 2)
    assert [len(v) for v in out["banks"]["cog"].banks.values()] == [2, 2]


def test_pi05_tiny_matches_reference():
    fx = torch.load(GOLDEN / "pi05_tiny.pt", weights_only=False)
    sd = seeded_state_dict(fx["shapes"], fx["seed"])
    i, ref = fx["inputs"], fx["outputs"]
    out = vla_oracle.pi05_forward(sd, fx["cfg"], i["input_ids"], i["attention_mask"], i["images"], i["image_masks"],
                                  i["actions"], i["noise"], i["time"])
    assert abs(out["loss"].item() - ref["loss"].item()) < 1e-5
    assert (out["v_t"] - ref["v_t"]).abs().max().item() < 1e-4
    for steps, r in ref["inference"].items():
        got = vla_oracle.pi05_inference(sd, fx["cfg"], i["input_ids"], i["attention_mask"], i["images"],
                                        i["image_masks"], r["noise"], steps)
        assert (got - r["actions"]).abs().max().item() < 1e-4


def test_memvla_inference_matches_reference():
    fx = torch.load(GOLDEN / "memvla_inference_tiny.pt", weights_only=False)
    sd = seeded_state_dict(fx["shapes"], fx["seed"])
    m = fx["cfg"]["mem"]
    banks = {r: vla_oracle.MemBankOracle(sd, "model.per_cog_mem_bank.", r, m["mem_length"], m["retrieval_layers"],
                                         m["dataloader_type"], m["use_timestep_pe"], m["fusion_type"],
                                         m["consolidate_type"], m["update_fused"]) for r in ("per", "cog")}
    for f, fr in enumerate(fx["frames"]):
        got = vla_oracle.memvla_inference(sd, fx["cfg"], banks, fr["input_ids"], fr["images"], fr["noise"], f)
        assert (got[0].clamp(-1, 1) - fr["actions"]).abs().max().item() < 1e-4


def test_pi0_attn_mask_truth_table():
    """make_attn_mask (pi0_arch.py:22-33): prefix bidirectional, state token sees prefix + itself, action tokens see
    prefix + state + each other; invalid positions neither attend nor are attended."""
    input_mask = torch.tensor([[True, True, False, True, True, True]])
    ar = torch.tensor([False, False, False, True, True, False])
    m = vla_oracle.pi0_make_attn_mask(input_mask, ar)[0].int().tolist()
    assert m == [[1, 1, 0, 0, 0, 0],
                 [1, 1, 0, 0, 0, 0],
                 [0, 0, 0, 0, 0, 0],
                 [1, 1, 0, 1, 0, 0],
                 [1, 1, 0, 1, 1, 1],
                 [1, 1, 0, 1, 1, 1]]


def test_cogact_inference_matches_reference():
    fx = torch.load(GOLDEN / "cogact_inference_tiny.pt", weights_only=False)
    sd = seeded_state_dict(fx["shapes"], fx["seed"])
    for scale, o in fx["outputs"].items():
        x = vla_oracle.cogact_inference(sd, fx["cfg"], fx["inputs"]["input_ids"], fx["inputs"]["images"], o["noise"],
                                        scale, 10)
        assert (x[0].clamp(-1, 1) - o["actions_sample0"]).abs().max().item() < 1e-4
    tmap, ac, ac_prev = vla_oracle.ddim_tables(100, 10)
    assert tmap == list(range(0, 100, 10)) and ac_prev[0] == 1.0 and abs(ac[0] - 0.999684309 ** 2) < 1e-8


@pytest.mark.parametrize("case", ["ce", "soft_ce", "ce_two_frames"])
def test_navila_tiny_matches_reference(case):
    """NaVILA (VLM + shifted CE / soft CE over the time tokens, one or two frames per row) vs the unmodified
    reference's loss, logits and gradients (oracle/make_golden.py:make_navila_tiny)."""
    fx = torch.load(GOLDEN / "navila_tiny.pt", weights_only=False)
    c = fx["cases"][case]
    sd = seeded_state_dict(fx["shapes"], fx["seed"])
    sd = {k: (v.requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    i = c["inputs"]
    out = vla_oracle.navila_forward(sd, fx["cfg"], i["input_ids"], i["attention_mask"], i["images"], i["labels"],
                                    time_token_ids=c["time_token_ids"], soft_ce_std=c["soft_ce_std"])
    assert abs(out["loss"].item() - c["outputs"]["loss"].item()) < 1e-5
    valid = c["outputs"]["valid"][:, :, None]
    assert ((out["logits"] * valid) - c["outputs"]["logits"]).abs().max().item() < 2e-4
    out["loss"].backward()
    for name, gref in c["outputs"]["grads"].items():
        rel = ((sd[name].grad - gref).norm() / gref.norm()).item()
        assert rel < 1e-4, (name, rel)
    for name in c["outputs"]["none_grad"]:
        g = sd[name].grad
        assert g is None or g.abs().max().item() == 0.0, name


def test_hybrid_cogact_tiny_matches_reference():
    """Text + action co-training (hybrid_cogact_arch.py:60-218): weighted action loss, text loss, NaN on text-less batches."""
    fx = torch.load(GOLDEN / "hybrid_cogact_tiny.pt", weights_only=False)
    sd = seeded_state_dict(fx["shapes"], fx["seed"])
    for name, case in fx["cases"].items():
        i = case["inputs"]
        out = vla_oracle.hybrid_cogact_forward(sd, fx["cfg"], i["input_ids"], i["attention_mask"], i["images"],
                                               i["actions"], i["labels"], i["has_action"], i["has_text"], i["noise"],
                                               i["timesteps"], i["drop_mask"], i["repeated_diffusion_steps"])
        ref = case["outputs"]
        assert abs(out["action_loss"].item() - ref["action_loss"].item()) < 1e-5
        if name == "no_text":
            assert torch.isnan(out["text_loss"]) and torch.isnan(ref["text_loss"])
        else:
            assert abs(out["text_loss"].item() - ref["text_loss"].item()) < 1e-5
