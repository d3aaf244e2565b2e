"""Mirror of dexbotic/model/oft/{oft_arch.py, oft_discrete_arch.py}: OFTConfig /
OFTForCausalLM with the `Linear` L1-regression head (oft_arch.py:58-166, oft/action_model/model.py:104-165) or the `DiT`
DiffusionActionHead (model.py:197-271; its diffusers scheduler is restated in model/ddim.py), and OFTDiscreteConfig /
OFTDiscreteForCausalLM (the discrete action tokenizer), with the reference's forward signatures.

Integer semantics (bit-exact, tests/test_gpu_oft.py): inference indices = argmax over the last num_bins-1
vocabulary logits, first maximum wins (oft_discrete_arch.py:222-224); bins -> continuous idx/(num_bins-1)*2-1
(oft/action_model/model.py:314-347); discretisation round-half-even (model.py:303-312).
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np
import torch

from .. import ops
from ..functional import CastFn, CrossEntropyFn, GatherRowsFn, Lin, LinearFn, MSELossFn, Norm, NormFn
from ..params import ParamSpec
from ._module import B200Module
from .ddim import DDIMScheduler
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
        self.config = config
        llm, vis = config.llm_config, config.mm_vision_tower
        d, V = cfg_get(llm, "hidden_size"), cfg_get(llm, "vocab_size")
        specs = (llm_specs(llm, trainable=not config.freeze_llm)
                 + clip_specs(vis, trainable=not config.freeze_mm_vision)
                 + projector_specs(config.mm_projector_type, cfg_get(vis, "hidden_size"), d,
                                   trainable=not config.freeze_mm_projector)
                 + [ParamSpec("lm_head.weight", (V, d), "lm_head", trainable=True, no_decay=False)])
        if config.use_proprio:                       # DiscreteActionHead.proprio_projector (model.py:298-301), fp32
            q = "model.action_head.proprio_projector."
            specs += [ParamSpec(q + "fc1.weight", (d, config.proprio_dim), "action_head", "fp32"),
                      ParamSpec(q + "fc1.bias", (d,), "action_head", "fp32"),
                      ParamSpec(q + "fc2.weight", (d, d), "action_head", "fp32"),
                      ParamSpec(q + "fc2.bias", (d,), "action_head", "fp32")]
        store = self._materialize(specs, device)
        for name in store.order:
            if store.slots[name].region == "B":
                self.get_parameter(name).grad = store.g(name)
        self.model_engine = DexboticVLMModel(store, config)
        self.model_engine.action_head = DiscreteActionHead(V, config.action_dim, config.chunk_size, config.num_bins)
        self.proprio = None
        if config.use_proprio:
            q = "model.action_head.proprio_projector."
            self.proprio = (Lin.of(store, q + "fc1.weight", q + "fc1.bias"), Lin.of(store, q + "fc2.weight", q + "fc2.bias"))
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
        B = input_ids.shape[0]
        P0 = 1 if cfg.use_proprio else 0            # a proprio state token precedes the placeholders (:132-137)
        emb, _, new_mask, pos, S, lengths = eng._prepare_inputs_labels_for_multimodal(
            input_ids, attention_mask, None, images, append_tokens=A + P0, append_token_id=[None] * P0 + [1] * A)
        x2d = emb.view(B * S, -1)
        if cfg.use_proprio:
            assert states is not None, "states is required when use_proprio is True"
            h = LinearFn.apply(states.float().contiguous(), self.proprio[0], "gelu", self.store, False, eng.anchor.t)
            st_tok = LinearFn.apply(h, self.proprio[1], None, self.store, True, None).to(emb.dtype)
            srow = (torch.arange(B, device=emb.device, dtype=torch.int32) * S + lengths).contiguous()
            x2d = InsertRowsFn.apply(x2d, st_tok, srow)
        hidden2d = eng.llm.forward(x2d, B, S, new_mask, pos)
        # extract_action_hidden_states (oft_arch.py:204-210): rows [len, len + A) of every sample (minus the state row)
        rows = (torch.arange(B, device=hidden2d.device, dtype=torch.int32)[:, None] * S + lengths[:, None] + P0
                + torch.arange(A, device=hidden2d.device, dtype=torch.int32)[None, :]).reshape(-1).contiguous()
        action_hidden = GatherRowsFn.apply(hidden2d, rows)                       # [B*A, D]
        logits2d = LinearFn.apply(action_hidden, self.lm_head_lin, None, self.store, True, None)   # lm_head, :168
        loss = None
        # the reference computes the loss only when `actions` is passed (labels are None by then): :171
        if discrete_action_labels is not None and actions is not None:
            loss = CrossEntropyFn.apply(logits2d, discrete_action_labels.reshape(-1))               # :169-191
        return CausalLMOutputDexbotic(loss=loss, logits=logits2d.view(B, A, -1))

    @torch.no_grad()
    def predict_action_bins(self, input_ids, images, attention_mask=None, states=None) -> torch.Tensor:
        """Parallel decoding (oft_discrete_arch.py:207-224): int64 bin indices [B, chunk*dim] in [0, num_bins-2]."""
        out = self.forward(input_ids=input_ids, attention_mask=attention_mask, images=images, states=states)
        B, A, V = out.logits.shape
        return ops.argmax_last(out.logits.reshape(B * A, V).contiguous(), self.config.num_bins - 1).view(B, A)

    @torch.no_grad()
    def inference_action(self, input_ids, image_tensor, inference_args={}, **kwargs):
        """oft_discrete_arch.py:207-235."""
        action_norms = inference_args.get("action_norms")
        idx = self.predict_action_bins(input_ids, image_tensor, states=inference_args.get("states"))
        cont = self.model_engine.action_head.discrete_tokens_to_continuous(idx)
        actions = cont[0].float().cpu().numpy()
        actions = np.clip(actions, -1, 1)                                 # _denorm, dexbotic_arch.py:546-563
        mn, mx = np.array(action_norms["min"]).reshape(1, -1), np.array(action_norms["max"]).reshape(1, -1)
        return (mn + (actions + 1) * 0.5 * (mx - mn)).tolist()

    @torch.no_grad()
    def generate_action(self, input_ids, pixel_values, attention_masks, temperature, inference_args={},
                        u: Optional[torch.Tensor] = None, **kwargs):
        """oft_discrete_arch.py:237-282: parallel decoding with temperature sampling instead of argmax — one draw per
        action token from softmax(logits[..., -(num_bins-1):] / temperature).  Returns (de-normalised actions for every
        sample, response token ids = bin index + vocab_size - num_bins + 1).  `u` [B, chunk*dim] injects the uniforms
        (parity / reproducibility); default: torch.rand on the device."""
        cfg = self.config
        assert "Discrete" in cfg.action_model_type, "generate_action is only for discrete action model."
        action_norms = inference_args.get("action_norms")
        out = self.forward(input_ids=input_ids, attention_mask=attention_masks, images=pixel_values,
                           states=inference_args.get("states"))
        B, A, V = out.logits.shape
        if u is None:
            u = torch.rand(B * A, device=out.logits.device, dtype=torch.float32)
        idx = ops.sample_last(out.logits.reshape(B * A, V).contiguous(), cfg.num_bins - 1, float(temperature),
                              u.reshape(-1).float().contiguous()).view(B, A)
        response_ids = idx + cfg.vocab_size - cfg.num_bins + 1
        cont = self.model_engine.action_head.discrete_tokens_to_continuous(idx)
        actions = np.clip(cont.float().cpu().numpy(), -1, 1)                       # _denorm, dexbotic_arch.py:546-563
        mn, mx = np.array(action_norms["min"]).reshape(1, 1, -1), np.array(action_norms["max"]).reshape(1, 1, -1)
        return (mn + (actions + 1) * 0.5 * (mx - mn)).tolist(), response_ids

    def zero_grad(self, set_to_none: bool = False):
        self.store.zero_grad()

    def optimizer_step(self, base_lr: float = 2e-5, mm_projector_lr=None, mm_vision_lr=None, betas=(0.9, 0.999),
                       eps: float = 1e-8, weight_decay: float = 0.0, max_grad_norm=1.0):
        lrs = {"llm": base_lr, "projector": mm_projector_lr or base_lr, "vision": mm_vision_lr or base_lr,
               "lm_head": base_lr}
        norm = self.store.adamw_step(lrs, betas, eps, weight_decay, max_grad_norm)
        self.model_engine.refresh()
        return norm


# ------------------------------------------------------------------------------------------------
# OFT with the L1-regression head
# ------------------------------------------------------------------------------------------------
def oft_linear_head_specs(d: int, action_dim: int, chunk: int, use_proprio: bool, proprio_dim, trainable: bool = True,
                          prefix: str = "model.action_head.") -> list[ParamSpec]:
    """L1RegressionActionHead (oft/action_model/model.py:133-152): fp32 storage like the other action heads."""
    g, c = "action_head", "fp32"
    P = lambda n, s, **k: ParamSpec(prefix + n, s, g, c, trainable=trainable, **k)  # noqa: E731
    din = d * action_dim
    sp = [P("action_query", (1, chunk * action_dim, d)),
          P("model.layer_norm1.weight", (din,)), P("model.layer_norm1.bias", (din,)),
          P("model.fc1.weight", (d, din)), P("model.fc1.bias", (d,))]
    for i in range(2):
        q = f"model.mlp_resnet_blocks.{i}.ffn."
        sp += [P(q + "0.weight", (d,)), P(q + "0.bias", (d,)), P(q + "1.weight", (d, d)), P(q + "1.bias", (d,))]
    sp += [P("model.layer_norm2.weight", (d,)), P("model.layer_norm2.bias", (d,)),
           P("model.fc2.weight", (action_dim, d)), P("model.fc2.bias", (action_dim,))]
    if use_proprio:
        sp += [P("proprio_projector.fc1.weight", (d, proprio_dim)), P("proprio_projector.fc1.bias", (d,)),
               P("proprio_projector.fc2.weight", (d, d)), P("proprio_projector.fc2.bias", (d,))]
    return sp


class InsertRowsFn(torch.autograd.Function):
    """insert_action_embedding (oft_arch.py:169-201): the sequence buffer already has zero rows at the action slots
    (the splice plan leaves them empty); this returns a copy with the action embeddings added there."""

    @staticmethod
    def forward(ctx, emb2d, act2d, idx):
        ctx.save_for_backward(idx)
        out = emb2d.clone()
        ops.scatter_rows_add_(act2d.contiguous(), idx, out)
        return out

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        g = g.contiguous()
        return g, ops.gather_rows(g, idx), None


class L1RegressionActionHead:
    """oft/action_model/model.py:104-165 (MLPResNet with two residual blocks)."""

    def __init__(self, store, d: int, action_dim: int, chunk: int, use_proprio: bool, prefix: str = "model.action_head."):
        self.store, self.d, self.action_dim, self.action_chunk = store, d, action_dim, chunk
        L = lambda n: Lin.of(store, prefix + n + ".weight", prefix + n + ".bias")  # noqa: E731
        N_ = lambda n: Norm("ln", 1e-5, store.w(prefix + n + ".weight"), store.w(prefix + n + ".bias"),  # noqa: E731
                            store.g(prefix + n + ".weight"), store.g(prefix + n + ".bias"))
        self.ln1, self.fc1 = N_("model.layer_norm1"), L("model.fc1")
        self.blocks = [(N_(f"model.mlp_resnet_blocks.{i}.ffn.0"), L(f"model.mlp_resnet_blocks.{i}.ffn.1")) for i in range(2)]
        self.ln2, self.fc2 = N_("model.layer_norm2"), L("model.fc2")
        self.proprio = (L("proprio_projector.fc1"), L("proprio_projector.fc2")) if use_proprio else None

    def predict_action(self, hidden2d_f32: torch.Tensor, B: int) -> torch.Tensor:
        """hidden2d [B*chunk*action_dim, d] fp32 -> [B, chunk, action_dim]."""
        st = self.store
        x = hidden2d_f32.reshape(B * self.action_chunk, self.action_dim * self.d)
        x = LinearFn.apply(NormFn.apply(x, self.ln1, st), self.fc1, "relu", st, True, None)
        for ln, lin in self.blocks:
            x = x + LinearFn.apply(NormFn.apply(x, ln, st), lin, "relu", st, True, None)
        x = LinearFn.apply(NormFn.apply(x, self.ln2, st), self.fc2, None, st, True, None)
        return x.view(B, self.action_chunk, self.action_dim)

    def proprio_projector(self, states: torch.Tensor, anchor) -> torch.Tensor:
        st = self.store
        h = LinearFn.apply(states.float().contiguous(), self.proprio[0], "gelu", st, False, anchor)
        return LinearFn.apply(h, self.proprio[1], None, st, True, None)


def oft_diffusion_head_specs(d: int, action_dim: int, use_proprio: bool, proprio_dim, trainable: bool = True,
                             prefix: str = "model.action_head.") -> list[ParamSpec]:
    """DiffusionActionHead (oft/action_model/model.py:197-226): NoisePredictionModel(MLPResNet) + NoisyActionProjector
    (+ ProprioProjector); the time encoder and the scheduler hold no parameters."""
    g, c = "action_head", "fp32"
    P = lambda n, s, **k: ParamSpec(prefix + n, s, g, c, trainable=trainable, **k)  # noqa: E731
    din, m = d * action_dim, "noise_predictor.mlp_resnet."
    sp = [P(m + "layer_norm1.weight", (din,)), P(m + "layer_norm1.bias", (din,)),
          P(m + "fc1.weight", (d, din)), P(m + "fc1.bias", (d,))]
    for i in range(2):
        q = f"{m}mlp_resnet_blocks.{i}.ffn."
        sp += [P(q + "0.weight", (d,)), P(q + "0.bias", (d,)), P(q + "1.weight", (d, d)), P(q + "1.bias", (d,))]
    sp += [P(m + "layer_norm2.weight", (d,)), P(m + "layer_norm2.bias", (d,)),
           P(m + "fc2.weight", (action_dim, d)), P(m + "fc2.bias", (action_dim,)),
           P("noisy_action_projector.fc1.weight", (d, 1)), P("noisy_action_projector.fc1.bias", (d,)),
           P("noisy_action_projector.fc2.weight", (d, d)), P("noisy_action_projector.fc2.bias", (d,))]
    if use_proprio:
        sp += [P("proprio_projector.fc1.weight", (d, proprio_dim)), P("proprio_projector.fc1.bias", (d,)),
               P("proprio_projector.fc2.weight", (d, d)), P("proprio_projector.fc2.bias", (d,))]
    return sp


def sinusoidal_timestep_encoding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """SinusoidalPositionalEncoding (oft/action_model/model.py:57-80): [sin(t * w) | cos(t * w)],
    w_i = exp(-i * ln(10000) / (dim/2 - 1)); fp32.  B values: host-sized work, evaluated with torch on the device."""
    assert dim % 2 == 0, f"# dimensions must be even but got {dim}"
    half = dim // 2
    w = torch.exp(torch.arange(half, device=t.device) * -math.log(10000) / (half - 1))
    e = t[:, None] * w[None, :]
    return torch.cat((e.sin(), e.cos()), dim=-1)


class DiffusionActionHead(L1RegressionActionHead):
    """oft/action_model/model.py:197-271: the MLPResNet predicts the noise that was added to the action chunk; the noisy
    actions enter the LLM as one token per scalar (NoisyActionProjector) behind a diffusion-timestep token."""

    def __init__(self, store, d: int, action_dim: int, chunk: int, use_proprio: bool, num_diffusion_steps: int = 100,
                 prefix: str = "model.action_head."):
        self.store, self.d, self.action_dim, self.action_chunk = store, d, action_dim, chunk
        L = lambda n: Lin.of(store, prefix + n + ".weight", prefix + n + ".bias")  # noqa: E731
        N_ = lambda n: Norm("ln", 1e-5, store.w(prefix + n + ".weight"), store.w(prefix + n + ".bias"),  # noqa: E731
                            store.g(prefix + n + ".weight"), store.g(prefix + n + ".bias"))
        m = "noise_predictor.mlp_resnet."
        self.ln1, self.fc1 = N_(m + "layer_norm1"), L(m + "fc1")
        self.blocks = [(N_(f"{m}mlp_resnet_blocks.{i}.ffn.0"), L(f"{m}mlp_resnet_blocks.{i}.ffn.1")) for i in range(2)]
        self.ln2, self.fc2 = N_(m + "layer_norm2"), L(m + "fc2")
        self.nap = (L("noisy_action_projector.fc1"), L("noisy_action_projector.fc2"))
        self.proprio = (L("proprio_projector.fc1"), L("proprio_projector.fc2")) if use_proprio else None
        self.noise_scheduler = DDIMScheduler(num_train_timesteps=num_diffusion_steps, beta_schedule="squaredcos_cap_v2")
        self.num_diffusion_steps = num_diffusion_steps

    def time_encoder(self, timesteps: torch.Tensor) -> torch.Tensor:
        return sinusoidal_timestep_encoding(timesteps, self.d)

    def sample_noisy_actions(self, ground_truth_actions: torch.Tensor) -> dict:
        """model.py:227-257: noise ~ N(0,1), one timestep per sample, closed-form forward diffusion, timestep token."""
        B, dev = ground_truth_actions.shape[0], ground_truth_actions.device
        noise = torch.randn(B, self.action_chunk, self.action_dim, device=dev, dtype=ground_truth_actions.dtype)
        timesteps = torch.randint(0, self.noise_scheduler.config.num_train_timesteps, (B,), device=dev)
        noisy = self.noise_scheduler.add_noise(ground_truth_actions, noise, timesteps)
        temb = self.time_encoder(timesteps).to(noisy.dtype).unsqueeze(1)
        return dict(noise=noise, noisy_actions=noisy, diffusion_timestep_embeddings=temb)

    def noisy_action_projector(self, noisy_actions: torch.Tensor, anchor) -> torch.Tensor:
        """[B, chunk, action_dim] -> [B * chunk * action_dim, d]: Linear(1, d), GELU, Linear(d, d) (model.py:33-55)."""
        st = self.store
        x = noisy_actions.reshape(-1, 1).float().contiguous()
        h = LinearFn.apply(x, self.nap[0], "gelu", st, False, anchor)
        return LinearFn.apply(h, self.nap[1], None, st, True, None)

    predict_noise = L1RegressionActionHead.predict_action       # same MLPResNet over [B, chunk, action_dim * d]


class OFTForCausalLM(B200Module):
    """oft_arch.py:50-251 with action_model_type 'Linear' (L1 regression) or 'DiT' (DiffusionActionHead)."""
    config_class = OFTConfig

    def __init__(self, config: OFTConfig, device="cuda"):
        super().__init__()
        amt = config.action_model_type or ""
        if "Linear" not in amt and "DiT" not in amt:
            raise NotImplementedError(f"OFTForCausalLM: action_model_type {amt!r}; 'Discrete' is OFTDiscreteForCausalLM")
        self.diffusion = "Linear" not in amt                    # builder.py:16-35 tests 'Linear' first
        self.config = config
        llm, vis = config.llm_config, config.mm_vision_tower
        d, V = cfg_get(llm, "hidden_size"), cfg_get(llm, "vocab_size")
        specs = (llm_specs(llm, trainable=not config.freeze_llm)
                 + clip_specs(vis, trainable=not config.freeze_mm_vision)
                 + projector_specs(config.mm_projector_type, cfg_get(vis, "hidden_size"), d,
                                   trainable=not config.freeze_mm_projector)
                 + (oft_diffusion_head_specs(d, config.action_dim, config.use_proprio, config.proprio_dim)
                    if self.diffusion else
                    oft_linear_head_specs(d, config.action_dim, config.chunk_size, config.use_proprio, config.proprio_dim))
                 + [ParamSpec("lm_head.weight", (V, d), "lm_head", trainable=False)])   # never used by these heads
        store = self._materialize(specs, device)
        for name in store.order:
            if store.slots[name].region == "B":
                self.get_parameter(name).grad = store.g(name)
        self.model_engine = DexboticVLMModel(store, config)
        head_cls = DiffusionActionHead if self.diffusion else L1RegressionActionHead
        self.model_engine.action_head = head_cls(store, d, config.action_dim, config.chunk_size, config.use_proprio)

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
                noisy_dict: Optional[dict] = None,
                **kwargs) -> CausalLMOutputDexbotic:
        if not input_ids.is_cuda:
            raise RuntimeError("dexbotic_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        cfg, eng, st = self.config, self.model_engine, self.store
        head = eng.action_head
        B, A, T = input_ids.shape[0], cfg.action_dim, cfg.chunk_size
        n_q = T * A
        n_act = n_q + (1 if cfg.use_proprio else 0) + (1 if self.diffusion else 0)
        if attention_mask is None:
            attention_mask = torch.ones_like(input_ids)
        a = None
        if actions is not None:                                                        # :81-83
            a = actions.reshape(B, -1, A)[:, :T, :].to(torch.float32)
        emb, _, new_mask, pos, S, lengths = eng._prepare_inputs_labels_for_multimodal(
            input_ids, attention_mask, None, images, append_tokens=n_act, append_token_id=None)
        dev = emb.device
        d = emb.shape[-1]
        noise = None
        if self.diffusion:
            # [timestep token | one token per noisy action scalar] (:106-116)
            if noisy_dict is None:
                assert a is not None, "actions (or a noisy_dict) are required by the diffusion head"
                noisy_dict = head.sample_noisy_actions(a)
            noise = noisy_dict["noise"]
            temb = noisy_dict["diffusion_timestep_embeddings"].to(device=dev, dtype=torch.float32)
            proj = head.noisy_action_projector(noisy_dict["noisy_actions"].to(dev), eng.anchor.t).view(B, n_q, d)
            act = torch.cat([temb.expand(B, 1, d), proj], dim=1)
        else:
            # the same learned action_query rows for every sample (:104-105)
            aq = self.get_parameter("model.action_head.action_query")                 # fp32 [1, n_q, d]
            act = aq.expand(B, n_q, d)
        if cfg.use_proprio:                                                            # proprio token in front (:118-121)
            assert states is not None, "states is required when use_proprio is True"
            s_tok = head.proprio_projector(states, eng.anchor.t).view(B, 1, d)
            act = torch.cat([s_tok, act], dim=1)
        act2d = act.to(emb.dtype).reshape(B * n_act, d)
        rows = (torch.arange(B, device=dev, dtype=torch.int32)[:, None] * S + lengths[:, None]
                + torch.arange(n_act, device=dev, dtype=torch.int32)[None, :]).reshape(-1).contiguous()
        emb2d = InsertRowsFn.apply(emb.view(B * S, d), act2d, rows)
        hidden2d = eng.llm.forward(emb2d, B, S, new_mask, pos)
        # extract_action_hidden_states (:204-210), minus the proprio row (:139-140) and the timestep row (:147)
        qrows = rows.view(B, n_act)[:, n_act - n_q:].reshape(-1).contiguous()
        ah = CastFn.apply(GatherRowsFn.apply(hidden2d, qrows), torch.float32)         # [B*n_q, d]
        predicted = head.predict_action(ah, B)                                         # actions, or the noise estimate
        loss = None
        if a is not None:                                                              # :149-154, fp32
            if self.diffusion:
                loss = MSELossFn.apply(predicted, noise.to(device=dev, dtype=torch.float32))
            else:
                loss = (a - predicted).abs().mean()
        return CausalLMOutputDexbotic(loss=loss, logits=predicted)

    @torch.no_grad()
    def inference_action(self, input_ids, image_tensor, inference_args={}, noise: Optional[torch.Tensor] = None, **kwargs):
        """oft_arch.py:212-251.  'Linear': one forward.  'DiT': DDIM loop from N(0,1) (`noise` lets a caller fix the
        starting sample), one model call per step with the timestep token of that step (:225-250)."""
        cfg, head = self.config, self.model_engine.action_head
        action_norms = inference_args.get("action_norms")
        states = inference_args.get("states")
        if not self.diffusion:
            predicted = self.forward(input_ids=input_ids, images=image_tensor, states=states).logits
        else:
            sched = head.noise_scheduler
            sched.set_timesteps(inference_args.get("num_ddim_steps", 10))
            dev = input_ids.device
            if noise is None:
                noise = torch.randn(input_ids.size(0), cfg.chunk_size, cfg.action_dim, device=dev, dtype=torch.float32)
            cur = noise.to(device=dev, dtype=torch.float32)
            for t in sched.timesteps.tolist():
                temb = head.time_encoder(torch.tensor([float(t)], device=dev)).unsqueeze(1)
                out = self.forward(input_ids=input_ids, images=image_tensor, states=states,
                                   noisy_dict=dict(noise=noise, noisy_actions=cur, diffusion_timestep_embeddings=temb))
                cur = sched.step(out.logits, t, cur).prev_sample
            predicted = cur
        actions = np.clip(predicted[0].float().cpu().numpy(), -1, 1)         # _denorm, dexbotic_arch.py:546-563
        mn, mx = np.array(action_norms["min"]).reshape(1, -1), np.array(action_norms["max"]).reshape(1, -1)
        return (mn + (actions + 1) * 0.5 * (mx - mn)).tolist()

    def zero_grad(self, set_to_none: bool = False):
        self.store.zero_grad()

    def optimizer_step(self, base_lr: float = 2e-5, mm_projector_lr=None, mm_vision_lr=None, action_head_lr=None,
                       betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0, max_grad_norm=1.0):
        lrs = {"llm": base_lr, "projector": mm_projector_lr or base_lr, "vision": mm_vision_lr or base_lr,
               "action_head": action_head_lr or base_lr, "lm_head": base_lr}
        norm = self.store.adamw_step(lrs, betas, eps, weight_decay, max_grad_norm)
        self.model_engine.refresh()
        return norm
