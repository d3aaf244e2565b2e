"""The reference-facing training surface: a loop shaped like DexboticTrainer (dexbotic/exp/trainer.py):
`model(**collator_batch)`, every `*_loss` key of the output read as the trainer's `compute_loss` does (:126-138), the
properties `_get_optimizer_grouped_parameters` reads (base_exp.py:95-203), and the torch.optim / LR-scheduler protocol
HF Trainer drives an optimizer with — for each policy class, on the tiny golden configurations."""
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"


def _cogact(hybrid=False):
    from dexbotic_b200.model import CogActConfig, CogACTForCausalLM, HybridCogACTForCausalLM
    fx = torch.load(GOLDEN / "cogact_tiny.pt", weights_only=False)
    cfg = fx["cfg"]
    c = CogActConfig(llm_config=cfg["llm"], mm_vision_tower=cfg["vision"], action_model_type="DiT-S", action_dim=7,
                     chunk_size=16)
    model = (HybridCogACTForCausalLM if hybrid else CogACTForCausalLM)(c)
    i = fx["inputs"]
    B = i["input_ids"].shape[0]
    batch = dict(input_ids=i["input_ids"], attention_mask=i["attention_mask"], images=i["images"], actions=i["actions"])
    if hybrid:
        labels = i["input_ids"].clone()
        labels[:, :6] = -100
        batch.update(labels=labels, has_action=torch.ones(B, 1, dtype=torch.long), has_text=torch.ones(B, 1, dtype=torch.long))
    return model, batch


def _pi0():
    from dexbotic_b200.model import Pi0Config, Pi0ForCausalLM
    fx = torch.load(GOLDEN / "pi0_tiny.pt", weights_only=False)
    cfg = fx["cfg"]
    model = Pi0ForCausalLM(Pi0Config(llm_config=cfg["llm"], action_config=cfg["expert"], vision_config=cfg["vision"],
                                     action_dim=cfg["action_dim"], chunk_size=cfg["chunk_size"]), device="cuda")
    i = fx["inputs"]
    return model, {k: i[k] for k in ("input_ids", "attention_mask", "images", "image_masks", "actions", "states")}


def _oft_discrete():
    from dexbotic_b200.model import OFTDiscreteConfig, OFTDiscreteForCausalLM
    fx = torch.load(GOLDEN / "oft_discrete_tiny.pt", weights_only=False)
    cfg = fx["cfg"]
    model = OFTDiscreteForCausalLM(OFTDiscreteConfig(llm_config=cfg["llm"], mm_vision_tower=cfg["vision"],
                                                     action_model_type="Discrete", action_dim=cfg["action_dim"],
                                                     chunk_size=cfg["chunk_size"], num_bins=cfg.get("num_bins", 256)))
    i = fx["inputs"]
    return model, {k: i[k] for k in ("input_ids", "attention_mask", "images", "labels", "actions")}


def _navila():
    from dexbotic_b200.model import NaVILAConfig, NaVILAForCausalLM
    fx = torch.load(GOLDEN / "navila_tiny.pt", weights_only=False)
    model = NaVILAForCausalLM(NaVILAConfig(llm_config=fx["cfg"]["llm"], mm_vision_tower=fx["cfg"]["vision"]))
    i = fx["cases"]["ce"]["inputs"]
    return model, dict(i)


@pytest.mark.parametrize("policy", ["cogact", "hybrid_cogact", "pi0", "oft_discrete", "navila"])
def test_trainer_shaped_loop(policy):
    from dexbotic_b200.optim import B200AdamW
    model, batch = {"cogact": _cogact, "hybrid_cogact": lambda: _cogact(True), "pi0": _pi0, "oft_discrete": _oft_discrete,
                    "navila": _navila}[policy]()
    model.init_weights_(seed=3)
    model.train()
    batch = {k: (v.cuda() if hasattr(v, "cuda") else v) for k, v in batch.items()}
    # ---- what _get_optimizer_grouped_parameters / _freeze_model read (base_exp.py:95-203, 318-330)
    inner = model.model
    names = [n for n, _ in model.named_parameters()]
    assert names and all(isinstance(p, torch.nn.Parameter) for _, p in model.named_parameters())
    assert any(inner.mm_projector_prefix in n for n in names) and any(inner.mm_vision_prefix in n for n in names)
    if policy in ("cogact", "hybrid_cogact"):
        assert any(inner.action_head_prefix in n for n in names) and inner.action_head_module is not None
    assert inner.backbone is not None and inner.mm_projector_module is not None and inner.mm_vision_module is not None
    assert model.supports_gradient_checkpointing and model.dtype == torch.bfloat16
    assert model.to(torch.bfloat16) is model            # the reference's `.to(dtype)` habit is a no-op, not a detach
    model.gradient_checkpointing_enable(gradient_checkpointing_kwargs={"use_reentrant": False})
    # ---- optimizer + scheduler as HF Trainer drives them
    opt = B200AdamW(model, lr=5e-4, mm_projector_lr=2e-4, weight_decay=0.0, max_grad_norm=1.0)
    assert {g["name"] for g in opt.param_groups} <= {"llm", "projector", "vision", "action_head", "lm_head"}
    assert sum(len(g["params"]) for g in opt.param_groups) == sum(1 for _, p in model.named_parameters() if p.requires_grad)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0 if s < 3 else 0.0)
    losses, logged = [], {}
    before = {n: p.detach().clone() for n, p in model.named_parameters() if p.requires_grad}
    for step in range(5):
        opt.zero_grad()
        outputs = model(**batch)                                            # DexboticTrainer.compute_loss -> model(**inputs)
        loss = outputs["loss"] if isinstance(outputs, dict) or hasattr(outputs, "keys") else outputs[0]
        for key in [k for k in outputs.keys() if k.endswith("_loss")]:      # trainer.py:129-137
            if outputs[key] is not None:
                logged[key] = outputs[key].detach().item()
        assert loss.dim() == 0 and loss.requires_grad
        loss.backward()
        if step == 3:
            frozen = {n: p.detach().clone() for n, p in model.named_parameters() if p.requires_grad}
        opt.step()
        sched.step()
        losses.append(loss.item())
    assert all(l == l and abs(l) < 1e4 for l in losses), losses             # finite
    if policy in ("oft_discrete", "navila"):                                # deterministic losses: it trains
        assert losses[2] < losses[0], losses                                # (the diffusion / flow losses draw fresh noise every step)
    moved = sum(int(not torch.equal(before[n], p.detach())) for n, p in model.named_parameters() if p.requires_grad)
    assert moved > 0.5 * len(before)
    # the scheduler's lr = 0 from step 3 on reached the kernels: nothing moved in steps 3 and 4
    assert all(torch.equal(frozen[n], p.detach()) for n, p in model.named_parameters() if p.requires_grad)
    if policy == "hybrid_cogact":
        assert set(logged) == {"text_loss", "action_loss"}
    sd = opt.state_dict()
    assert sd["step"] == 5 and sd["exp_avg"] is not None
    opt.load_state_dict(sd)
