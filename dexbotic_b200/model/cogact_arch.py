"""Mirror of dexbotic/model/cogact/cogact_arch.py: CogActConfig / CogActModel / CogACTForCausalLM with the
reference's forward(...) keyword signature and CausalLMOutputDexbotic return type (cogact_arch.py:56-147),
running on the B200 kernels.  state_dict keys == the reference's (model.llm.*, model.mm_vision_tower.*,
model.mm_projector.*, model.action_head.*, lm_head.weight).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from .. import ops
from ..functional import CastFn, CrossEntropyFn, GatherRowsFn, Lin, LinearFn
from ..params import ParamSpec
from ._module import B200Module
from .action_model import ActionModel, action_head_specs
from .dexbotic_arch import (CausalLMOutputDexbotic, DexboticConfig, DexboticVLMModel, cfg_get, clip_specs, llm_specs,
                            projector_specs)


class CogActConfig(DexboticConfig):
    """cogact_arch.py:13-17."""
    model_type = "dexbotic_cogact"

    def __init__(self, action_model_type: Optional[str] = None, action_dim: Optional[int] = None,
                 chunk_size: Optional[int] = None, **kwargs):
        super().__init__(**kwargs)
        self.action_model_type = action_model_type
        self.action_dim = action_dim
        self.chunk_size = chunk_size
        self.freeze_action_head = kwargs.get("freeze_action_head", False)


class CogActModel(DexboticVLMModel):
    """cogact_arch.py:20-44 (`model.model`): VLM shell + action head."""
    action_head_prefix = "action_head"

    def __init__(self, store, config: CogActConfig):
        super().__init__(store, config)
        self.action_head = ActionModel(store, config.action_model_type, cfg_get(config.llm_config, "hidden_size"),
                                       config.action_dim, config.chunk_size - 1)   # action_model/builder.py:7-27

    @property
    def action_head_module(self):
        return self.action_head


class CogACTForCausalLM(B200Module):
    """cogact_arch.py:47-198."""
    config_class = CogActConfig
    lm_head_trainable = False      # lm_head exists in the reference (cogact_arch.py:52) but is unused by plain CogACT

    def __init__(self, config: CogActConfig, device="cuda"):
        super().__init__()
        self.config = config
        llm = config.llm_config
        d, V = cfg_get(llm, "hidden_size"), cfg_get(llm, "vocab_size")
        vis = config.mm_vision_tower
        specs = (llm_specs(llm, trainable=not config.freeze_llm)
                 + clip_specs(vis, trainable=not config.freeze_mm_vision)
                 + projector_specs(config.mm_projector_type, cfg_get(vis, "hidden_size"), d,
                                   trainable=not config.freeze_mm_projector)
                 + action_head_specs(config.action_model_type, d, config.action_dim, config.chunk_size,
                                     trainable=not getattr(config, "freeze_action_head", False))
                 # lm_head never receives a gradient in CogACT; the hybrid (text + action) variant trains it
                 + [ParamSpec("lm_head.weight", (V, d), "lm_head", trainable=self.lm_head_trainable, no_decay=False)])
        store = self._materialize(specs, device)
        # fp32 (region B) parameters take their torch .grad directly from the flat gradient buffer
        for name in store.order:
            s = store.slots[name]
            if s.region == "B":
                self.get_parameter(name).grad = store.g(name)
        self.model_engine = CogActModel(store, config)

    # `model.model` must stay the nn.Module that owns the parameters (state-dict keys); the engine with the
    # reference's properties (backbone, mm_projector_prefix, action_head_module ...) is `model_engine`.
    @property
    def engine(self) -> CogActModel:
        return self.model_engine

    def _after_weights_changed(self) -> None:
        self.model_engine.refresh()

    def forward(self,
                input_ids: torch.LongTensor = None,
                attention_mask: Optional[torch.Tensor] = None,
                position_ids: Optional[torch.LongTensor] = None,
                past_key_values: Optional[List[torch.FloatTensor]] = None,
                inputs_embeds: Optional[torch.FloatTensor] = None,
                labels: Optional[torch.LongTensor] = None,
                use_cache: Optional[bool] = None,
                output_attentions: Optional[bool] = None,
                output_hidden_states: Optional[bool] = None,
                images: Optional[torch.FloatTensor] = None,
                return_dict: Optional[bool] = None,
                cache_position: Optional[torch.LongTensor] = None,
                actions: Optional[torch.LongTensor] = None,
                states: Optional[torch.LongTensor] = None,
                repeated_diffusion_steps: int = 4,
                noise: Optional[torch.Tensor] = None,          # parity hooks: inject the reference's random draws
                timesteps: Optional[torch.Tensor] = None,
                drop_mask: Optional[torch.Tensor] = None,
                **kwargs) -> CausalLMOutputDexbotic:
        if images is None or input_ids is None:
            raise NotImplementedError("CogACT training forward needs input_ids and images (cogact_arch.py:75-91)")
        if not input_ids.is_cuda:
            raise RuntimeError("dexbotic_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        eng = self.model_engine
        cfg = self.config
        emb, new_labels, new_mask, pos, S, _ = eng._prepare_inputs_labels_for_multimodal(
            input_ids, attention_mask, labels, images)
        B = input_ids.shape[0]
        hidden2d = eng.llm.forward(emb.view(B * S, -1), B, S, new_mask, pos)           # cogact_arch.py:97-108
        last_hidden_state = hidden2d.view(B, S, -1)

        loss = None
        if actions is not None:
            idx = ops.last_valid_index(new_mask)                                        # :110-120
            cognition = GatherRowsFn.apply(hidden2d, idx)                               # [B, D]
            cog32 = CastFn.apply(cognition, torch.float32)[:, None, :]                  # autocast(float32), :133
            a = actions.reshape(B, -1, cfg.action_dim).to(torch.float32)[:, :cfg.chunk_size, :]
            R = repeated_diffusion_steps
            loss = eng.action_head.loss(self, a.repeat(R, 1, 1), cog32.repeat(R, 1, 1), noise, timesteps, drop_mask,
                                        training=self.training)
        return CausalLMOutputDexbotic(loss=loss, logits=last_hidden_state)

    @torch.no_grad()
    def inference_action(self, input_ids, image_tensor, inference_args={}, noise: Optional[torch.Tensor] = None, **kwargs):
        """cogact_arch.py:149-198: one VLM forward -> cognition token -> CFG + DDIM -> denormalised action chunk."""
        cfg_scale = inference_args.get("cfg_scale", 1.5)
        num_ddim_steps = inference_args.get("num_ddim_steps", 10)
        action_norms = inference_args.get("action_norms")
        out = self.forward(input_ids=input_ids, images=image_tensor)
        cognition = out.logits[:, -1, :].float()[:, None, :].contiguous()            # :158
        B = cognition.shape[0]
        if noise is None:
            noise = torch.randn(B, self.config.chunk_size, self.config.action_dim, device=cognition.device)
        samples = self.model_engine.action_head.sample(self, cognition, noise.float(), cfg_scale, num_ddim_steps)
        actions = np.clip(samples[0].float().cpu().numpy(), -1, 1)                   # _denorm, dexbotic_arch.py:546-563
        mn = np.array(action_norms["min"]).reshape(1, -1)
        mx = np.array(action_norms["max"]).reshape(1, -1)
        return (mn + (actions + 1) * 0.5 * (mx - mn)).tolist()

    # ---- training utilities that the reference delegates to HF Trainer / DeepSpeed ------------------
    def zero_grad(self, set_to_none: bool = False):        # noqa: D401
        self.store.zero_grad()

    def optimizer_step(self, base_lr: float = 2e-5, mm_projector_lr=None, mm_vision_lr=None, action_head_lr=None,
                       betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0, max_grad_norm=1.0):
        """Fused AdamW over the flat buffers with the reference's per-module learning rates
        (OptimizerConfig, base_exp.py:64-203) and max_grad_norm=1.0 (trainer.py:122)."""
        lrs = {"llm": base_lr, "projector": mm_projector_lr or base_lr, "vision": mm_vision_lr or base_lr,
               "action_head": action_head_lr or base_lr, "lm_head": base_lr}
        norm = self.store.adamw_step(lrs, betas, eps, weight_decay, max_grad_norm)
        self.model_engine.refresh()
        return norm


class HybridCogACTForCausalLM(CogACTForCausalLM):
    """cogact/hybrid_cogact_arch.py:51-255: co-training of language and actions.  On top of CogACT's forward:
    text_loss = HF ForCausalLMLoss(lm_head(hidden), labels) * has_text.any() and an action loss weighted per sample by
    has_action; `loss = text_loss + action_loss`, both also returned (DexboticTrainer.compute_loss logs every *_loss,
    trainer.py:126-138).  As in the reference (:131-141), the rows without text are masked only when NO row has text,
    which leaves zero targets: text_loss (and the sum) is NaN for such a batch — reproduced, not repaired."""
    lm_head_trainable = True

    def __init__(self, config: CogActConfig, device="cuda"):
        super().__init__(config, device=device)
        self.lm_head_lin = Lin.of(self.store, "lm_head.weight")

    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None, inputs_embeds=None,
                labels=None, use_cache=None, output_attentions=None, output_hidden_states=None, images=None,
                return_dict=None, cache_position=None, actions=None, states=None, repeated_diffusion_steps: int = 4,
                has_action: Optional[torch.Tensor] = None, has_text: Optional[torch.Tensor] = None,
                noise=None, timesteps=None, drop_mask=None, **kwargs) -> CausalLMOutputDexbotic:
        if images is None or input_ids is None:
            raise NotImplementedError("training forward needs input_ids and images")
        if not input_ids.is_cuda:
            raise RuntimeError("dexbotic_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        eng, cfg = self.model_engine, self.config
        emb, new_labels, new_mask, pos, S, _ = eng._prepare_inputs_labels_for_multimodal(
            input_ids, attention_mask, labels, images)
        B = input_ids.shape[0]
        hidden2d = eng.llm.forward(emb.view(B * S, -1), B, S, new_mask, pos)
        text_loss = action_loss = None
        if labels is not None:
            assert has_action is not None and has_text is not None, "has_action / has_text must be provided"
            ht = has_text.bool().view(-1)
            any_text = bool(ht.any())                              # one host sync, as the reference's `if` (:134)
            tgt = new_labels[:, 1:]
            keep = (tgt != -100) if any_text else torch.zeros_like(tgt, dtype=torch.bool)
            bs = keep.nonzero(as_tuple=False)
            if bs.shape[0] == 0:                                   # mean over zero targets: NaN, times has_text.any()
                text_loss = hidden2d.new_full((), float("nan"), dtype=torch.float32)
            else:
                rows = (bs[:, 0] * S + bs[:, 1]).to(torch.int32).contiguous()
                logits2d = LinearFn.apply(GatherRowsFn.apply(hidden2d, rows), self.lm_head_lin, None, self.store, True,
                                          None)
                text_loss = CrossEntropyFn.apply(logits2d, tgt[keep].contiguous()) * float(any_text)
        if actions is not None:
            assert has_action is not None, "has_action must be provided"
            idx = ops.last_valid_index(new_mask)
            cog32 = CastFn.apply(GatherRowsFn.apply(hidden2d, idx), torch.float32)[:, None, :]
            a = actions.reshape(B, -1, cfg.action_dim).to(torch.float32)[:, :cfg.chunk_size, :]
            R = repeated_diffusion_steps
            w = has_action.reshape(-1).to(torch.float32).repeat(R)
            action_loss = eng.action_head.loss(self, a.repeat(R, 1, 1), cog32.repeat(R, 1, 1), noise, timesteps,
                                               drop_mask, training=self.training, sample_weight=w)
        loss = None
        if text_loss is not None and action_loss is not None:
            loss = text_loss + action_loss
        elif text_loss is not None:
            loss = text_loss
        elif action_loss is not None:
            loss = action_loss
        out = CausalLMOutputDexbotic(loss=loss, logits=hidden2d.view(B, S, -1))
        out.text_loss, out.action_loss = text_loss, action_loss
        return out
