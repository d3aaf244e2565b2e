"""Mirror of dexbotic/model/oft/{oft_arch.py, oft_discrete_arch.py} for the DISCRETE action tokenizer
(SURVEY.md §8a row A9): OFTDiscreteConfig / OFTDiscreteForCausalLM with the reference's forward signature.

Integer semantics (bit-exact, tests/test_gpu_oft.py): inference indices = argmax over the last num_bins-1
vocabulary logits, first maximum wins (oft_discrete_arch.py:222-224); bins -> continuous idx/(num_bins-1)*2-1
(oft/action_model/model.py:314-347); discretisation round-half-even (model.py:303-312).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from .. import ops
from ..functional import CrossEntropyFn, GatherRowsFn, Lin, LinearFn
from ..params import ParamSpec
from ._module import B200Module
from .dexbotic_arch import (CausalLMOutputDexbotic, DexboticConfig, DexboticVLMModel, cfg_get, clip_specs, llm_specs,
                            projector_specs)


class OFTConfig(DexboticConfig):
    """oft_arch.py:13-19."""
    model_type = "dexbotic_oft"

    def __init__(self, action_model_type: Optional[str] = None, action_dim: Optional[int] = None,
                 chunk_size: Optional[int] = None, use_proprio: bool = False, proprio_dim: Optional[int] = None, **kw):
        super().__init__(**kw)
        self.action_model_type, self.action_dim, self.chunk_size = action_model_type, action_dim, chunk_size
        self.use_proprio, self.proprio_dim = use_proprio, proprio_dim


class OFTDiscreteConfig(OFTConfig):
    """oft_discrete_arch.py:11-13."""
    model_type = "dexbotic_oft_discrete"

    def __init__(self, num_bins: int = 256, **kw):
        super().__init__(**kw)
        self.num_bins = num_bins


class DiscreteActionHead:
    """oft/action_model/model.py:273-347 (no parameters unless use_proprio)."""

    def __init__(self, vocab_size: int, action_dim: int, action_chunk: int, num_bins: int = 256):
        self.vocab_size, self.action_dim, self.action_chunk, self.num_bins = vocab_size, action_dim, action_chunk, num_bins

    def discretize_actions(self, actions: torch.Tensor) -> torch.Tensor:
        return ops.discretize_actions(actions.float().contiguous(), self.num_bins)

    def continuous_to_discrete_tokens(self, actions: torch.Tensor) -> torch.Tensor:
        d = self.discretize_actions(actions)
        return d.reshape(d.size(0), -1)

    def discrete_tokens_to_continuous(self, token_ids: torch.Tensor) -> torch.Tensor:
        a = ops.bins_to_continuous(token_ids.contiguous(), self.num_bins)
        return a.reshape(token_ids.size(0), self.action_chunk, self.action_dim)


class OFTDiscreteForCausalLM(B200Module):
    """oft_discrete_arch.py:20-282 on the B200 kernels."""
    config_class = OFTDiscreteConfig

    def __init__(self, config: OFTDiscreteConfig, device="cuda"):
        super().__init__()
        assert "Discrete" in config.action_model_type, "this class mirrors the OFT-discrete forward only"
        if config.use_proprio:
            raise NotImplementedError("use_proprio (ProprioProjector) is not wired yet")
        self.config = config
        llm, vis = config.llm_config, config.mm_vision_tower
        d, V = cfg_get(llm, "hidden_size"), cfg_get(llm, "vocab_size")
        specs = (llm_specs(llm, trainable=not config.freeze_llm)
                 + clip_specs(vis, trainable=not config.freeze_mm_vision)
                 + projector_specs(config.mm_projector_type, cfg_get(vis, "hidden_size"), d,
                                   trainable=not config.freeze_mm_projector)
                 + [ParamSpec("lm_head.weight", (V, d), "lm_head", trainable=True, no_decay=False)])
        store = self._materialize(specs, device)
        self.model_engine = DexboticVLMModel(store, config)
        self.model_engine.action_head = DiscreteActionHead(V, config.action_dim, config.chunk_size, config.num_bins)
        self.lm_head_lin = Lin.of(store, "lm_head.weight")

    def _after_weights_changed(self) -> None:
        self.model_engine.refresh()

    @staticmethod
    def _strip_action_labels(input_ids, attention_mask, labels, A: int):
        """oft_discrete_arch.py:66-106, vectorised: drop the A action-label tokens that precede the last valid
        token of every sample; returns (ids [B, L-A], mask [B, L-A], action labels [B, A])."""
        B, L = input_ids.shape
        npl = attention_mask.sum(dim=1)                                  # non-padding length
        prefix = npl - A - 1
        idx = torch.arange(L - A, device=input_ids.device)[None, :]
        srcpos = torch.where(idx < prefix[:, None], idx, idx + A)
        new_ids = torch.gather(input_ids, 1, srcpos)
        new_mask = (idx < (prefix + 1)[:, None]).to(attention_mask.dtype)
        lab_pos = prefix[:, None] + torch.arange(A, device=input_ids.device)[None, :]
        return new_ids, new_mask, torch.gather(labels, 1, lab_pos)

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
                noisy_dict: Optional[dict] = None,
                **kwargs) -> CausalLMOutputDexbotic:
        if not input_ids.is_cuda:
            raise RuntimeError("dexbotic_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        cfg, eng = self.config, self.model_engine
        A = cfg.chunk_size * cfg.action_dim
        discrete_action_labels = None
        if attention_mask is None:
            attention_mask = torch.ones_like(input_ids)
        if labels is not None:
            input_ids, attention_mask, discrete_action_labels = self._strip_action_labels(input_ids, attention_mask,
                                                                                          labels, A)
        emb, _, new_mask, pos, S, lengths = eng._prepare_inputs_labels_for_multimodal(
            input_ids, attention_mask, None, images, append_tokens=A, append_token_id=1)
        B = input_ids.shape[0]
        hidden2d = eng.llm.forward(emb.view(B * S, -1), B, S, new_mask, pos)
        # extract_action_hidden_states (oft_arch.py:204-210): rows [len, len + A) of every sample
        rows = (torch.arange(B, device=hidden2d.device, dtype=torch.int32)[:, None] * S + lengths[:, None]
                + torch.arange(A, device=hidden2d.device, dtype=torch.int32)[None, :]).reshape(-1).contiguous()
        action_hidden = GatherRowsFn.apply(hidden2d, rows)                       # [B*A, D]
        logits2d = LinearFn.apply(action_hidden, self.lm_head_lin, None, self.store, True, None)   # lm_head, :168
        loss = None
        # the reference computes the loss only when `actions` is passed (labels are None by then): :171
        if discrete_action_labels is not None and actions is not None:
            loss = CrossEntropyFn.apply(logits2d, discrete_action_labels.reshape(-1))               # :169-191
        return CausalLMOutputDexbotic(loss=loss, logits=logits2d.view(B, A, -1))

    @torch.no_grad()
    def predict_action_bins(self, input_ids, images, attention_mask=None) -> torch.Tensor:
        """Parallel decoding (oft_discrete_arch.py:207-224): int64 bin indices [B, chunk*dim] in [0, num_bins-2]."""
        out = self.forward(input_ids=input_ids, attention_mask=attention_mask, images=images)
        B, A, V = out.logits.shape
        return ops.argmax_last(out.logits.reshape(B * A, V).contiguous(), self.config.num_bins - 1).view(B, A)

    @torch.no_grad()
    def inference_action(self, input_ids, image_tensor, inference_args={}, **kwargs):
        """oft_discrete_arch.py:207-235."""
        action_norms = inference_args.get("action_norms")
        idx = self.predict_action_bins(input_ids, image_tensor)
        cont = self.model_engine.action_head.discrete_tokens_to_continuous(idx)
        actions = cont[0].float().cpu().numpy()
        actions = np.clip(actions, -1, 1)                                 # _denorm, dexbotic_arch.py:546-563
        mn, mx = np.array(action_norms["min"]).reshape(1, -1), np.array(action_norms["max"]).reshape(1, -1)
        return (mn + (actions + 1) * 0.5 * (mx - mn)).tolist()

    def zero_grad(self, set_to_none: bool = False):
        self.store.zero_grad()

    def optimizer_step(self, base_lr: float = 2e-5, mm_projector_lr=None, mm_vision_lr=None, betas=(0.9, 0.999),
                       eps: float = 1e-8, weight_decay: float = 0.0, max_grad_norm=1.0):
        lrs = {"llm": base_lr, "projector": mm_projector_lr or base_lr, "vision": mm_vision_lr or base_lr,
               "lm_head": base_lr}
        norm = self.store.adamw_step(lrs, betas, eps, weight_decay, max_grad_norm)
        self.model_engine.refresh()
        return norm
