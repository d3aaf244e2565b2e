"""GPU parity of OFT-discrete against the golden vectors produced by the UNMODIFIED reference:
training loss / logits / gradients within bf16 tolerance; the integer path (argmax over the last 255 logits,
bins -> continuous, discretisation) bit-exact."""
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"


def _build(fx):
    from dexbotic_b200.model import OFTDiscreteConfig, OFTDiscreteForCausalLM
    from oracle.weights import seeded_state_dict
    cfg = fx["cfg"]
    c = OFTDiscreteConfig(llm_config=cfg["llm"], mm_vision_tower=cfg["vision"], action_model_type="Discrete",
                          action_dim=cfg["action_dim"], chunk_size=cfg["chunk_size"], num_bins=cfg["num_bins"])
    model = OFTDiscreteForCausalLM(c, device="cuda")
    sd = {k: v for k, v in seeded_state_dict(fx["shapes"], fx["seed"]).items() if "position_ids" not in k}
    model.load_state_dict(sd, strict=True)
    return model


def _rel(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    return ((a - b).norm() / (b.norm() + 1e-12)).item(), torch.nn.functional.cosine_similarity(a, b, dim=0).item()


def test_oft_discrete_training_matches_reference_golden():
    fx = torch.load(GOLDEN / "oft_discrete_tiny.pt", weights_only=False)
    model = _build(fx)
    ours = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert ours == {k: tuple(v) for k, v in fx["shapes"].items() if "position_ids" not in k}
    model.train()
    i = {k: v.cuda() for k, v in fx["inputs"].items()}
    model.zero_grad()
    out = model(input_ids=i["input_ids"], attention_mask=i["attention_mask"], images=i["images"], labels=i["labels"],
                actions=i["actions"])
    ref = fx["outputs"]
    assert abs(out.loss.item() - ref["loss"].item()) < 2e-2 * abs(ref["loss"].item()), (out.loss.item(), ref["loss"].item())
    rel, cos = _rel(out.logits, ref["logits"].cuda())
    assert rel < 4e-2 and cos > 0.999, (rel, cos)
    out.loss.backward()
    for name, gref in ref["grads"].items():
        rel, cos = _rel(model.store.g(name), gref.cuda())
        assert rel < 0.1 and cos > 0.995, f"grad {name}: rel={rel:.4f} cos={cos:.5f}"
    # without `actions` the reference returns loss=None (oft_discrete_arch.py:171)
    out2 = model(input_ids=i["input_ids"], attention_mask=i["attention_mask"], images=i["images"], labels=i["labels"])
    assert out2.loss is None


def test_oft_integer_path_bit_exact():
    from dexbotic_b200 import ops
    fx = torch.load(GOLDEN / "oft_discrete_tiny.pt", weights_only=False)
    ref = fx["outputs"]
    # (1) the reference's own inference logits -> identical integer indices and identical continuous actions
    lg = ref["infer_logits"].cuda()
    B, A, V = lg.shape
    idx = ops.argmax_last(lg.reshape(B * A, V).contiguous(), 255).view(B, A)
    assert torch.equal(idx.cpu(), ref["infer_idx"])
    cont = ops.bins_to_continuous(idx, 256).view(B, 8, 7)
    assert torch.equal(cont.cpu(), ref["infer_cont"])
    # (2) KAT vectors of the tokenizer itself
    k = torch.load(GOLDEN / "oft_integer_kats.pt", weights_only=False)
    assert torch.equal(ops.discretize_actions(k["actions"].cuda().contiguous(), 256).cpu(), k["bins"])
    assert torch.equal(ops.bins_to_continuous(k["bins"].cuda().contiguous(), 256).view_as(k["continuous"]).cpu(),
                       k["continuous"])
    kl = k["logits"].cuda()
    assert torch.equal(ops.argmax_last(kl.reshape(-1, kl.shape[-1]).contiguous(), 255).view(kl.shape[:2]).cpu(),
                       k["argmax"])
    # (3) end to end: our bf16 logits give the reference's indices wherever its top-2 margin is not a near-tie
    model = _build(fx)
    model.eval()
    i = fx["inputs"]
    got = model.predict_action_bins(i["input_ids"][:, :11].cuda(), i["images"].cuda(),
                                    torch.ones(B, 11, dtype=torch.long, device="cuda")).cpu()
    top2 = ref["infer_logits"][:, :, -255:].topk(2, dim=-1).values
    clear = (top2[..., 0] - top2[..., 1]) > 0.05
    assert clear.float().mean().item() > 0.3
    assert torch.equal(got[clear], ref["infer_idx"][clear])
    acts = model.inference_action(i["input_ids"][:1, :11].cuda(), i["images"][:1].cuda(),
                                  {"action_norms": {"min": [-1.0] * 7, "max": [1.0] * 7}})
    assert np.asarray(acts).shape == (8, 7)
