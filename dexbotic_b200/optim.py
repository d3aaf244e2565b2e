"""torch.optim.Optimizer face of the fused flat-buffer AdamW, so that code written for the reference's training stack can
drive this backend: HF `Trainer(optimizers=(optimizer, scheduler))` / `DexboticTrainer.create_optimizer`
(dexbotic/exp/trainer.py:25-36) build a torch optimizer from parameter groups with per-module learning rates
(OptimizerConfig._get_optimizer_grouped_parameters, base_exp.py:95-203) and LR schedulers mutate `param_groups[i]["lr"]`.

The trunk's gradients live in the ParamStore's flat bf16 buffer (written by the wgrad GEMM epilogues), not in `.grad`,
so a foreign optimizer cannot see them; this class is the optimizer: `step()` runs ParamStore.adamw_step (global-norm
clip + AdamW for every group in a handful of launches) with the groups' CURRENT learning rates, `zero_grad()` resets
the store.  One parameter group per module family (llm / projector / vision / action_head / lm_head), exactly the split
the reference makes; weight decay follows the reference's rule inside the store (LayerNorm children and *bias* names
are not decayed)."""
from __future__ import annotations

from typing import Optional

import torch

GROUPS = ("llm", "projector", "vision", "action_head", "lm_head")


class B200AdamW(torch.optim.Optimizer):
    def __init__(self, model, lr: float = 2e-5, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 mm_projector_lr: Optional[float] = None, mm_vision_lr: Optional[float] = None,
                 action_head_lr: Optional[float] = None, max_grad_norm: Optional[float] = 1.0):
        self.model, self.store, self.max_grad_norm = model, model.store, max_grad_norm
        group_lr = {"llm": lr, "projector": mm_projector_lr or lr, "vision": mm_vision_lr or lr,
                    "action_head": action_head_lr or lr, "lm_head": lr}
        by_group = {g: [] for g in GROUPS}
        for name in self.store.order:
            sp = self.store.slots[name].spec
            if sp.trainable:
                by_group[sp.group if sp.group in by_group else "llm"].append(model.get_parameter(name))
        groups = [dict(params=ps, name=g, lr=group_lr[g]) for g, ps in by_group.items() if ps]
        super().__init__(groups, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.last_grad_norm = None

    @torch.no_grad()
    def step(self, closure=None):
        loss = None if closure is None else closure()
        lrs = {g["name"]: float(g["lr"]) for g in self.param_groups}
        lrs.setdefault("llm", self.defaults["lr"])
        d = self.param_groups[0]
        self.last_grad_norm = self.store.adamw_step(lrs, tuple(d["betas"]), d["eps"], d["weight_decay"],
                                                    self.max_grad_norm)
        eng = getattr(self.model, "model_engine", None)
        if eng is not None and hasattr(eng, "refresh"):
            eng.refresh()
        return loss

    def zero_grad(self, set_to_none: bool = True):
        self.store.zero_grad()

    # optimizer state = the store's flat moment buffers (or this rank's shard of them under ZeRO-1)
    def state_dict(self):
        st = self.store
        own = st.sharder if (st.sharder is not None and st.sharder.enabled) else st
        return {"step": st.step_count, "sharded": own is not st,
                "exp_avg": None if own.exp_avg is None else own.exp_avg.detach().cpu(),
                "exp_avg_sq": None if own.exp_avg_sq is None else own.exp_avg_sq.detach().cpu(),
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, sd):
        st = self.store
        own = st.sharder if (st.sharder is not None and st.sharder.enabled) else st
        st.wait_all_params()
        st.step_count = int(sd["step"])
        if sd["exp_avg"] is not None:
            own.exp_avg = sd["exp_avg"].to(st.device).clone()
            own.exp_avg_sq = sd["exp_avg_sq"].to(st.device).clone()
        for g, saved in zip(self.param_groups, sd["param_groups"]):
            g.update({k: v for k, v in saved.items() if k != "name"})
