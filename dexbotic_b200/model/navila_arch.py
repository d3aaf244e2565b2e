"""Mirror of dexbotic/model/navila/navila_arch.py (NaVILA = VLM + (soft) cross entropy over text
tokens): NaVILAConfig / NaVILAForCausalLM with the reference's forward signature and state-dict keys.

Path (navila_arch.py:362-497, training branch without sequence packing — HF decoders take no `seqlens_in_batch`, :415-
417): SigLIP tower, select_layer -2 (siglip_encoder.py:13,64-65) -> mlp_downsample projector (2x2 token merge, LayerNorm,
Linear, GELU, Linear; mm_projector/builder.py:9-33,61-68) -> image-token splice -> Llama / Qwen2 decoder -> lm_head ->
HF ForCausalLMLoss (shifted, mean over labels != -100) or navila/loss.py's soft cross entropy over the time tokens.
Only the rows whose shifted label is a target go through the lm_head GEMM and the fused fp32 cross entropy (the
reference materialises [B, S, V] logits)."""
from __future__ import annotations

from typing import List, Optional

import torch

from .. import ops
from ..functional import CrossEntropyFn, GatherRowsFn, Lin, LinearFn, Norm, NormFn, SpliceFn
from ..params import ParamSpec
from ._module import B200Module
from .dexbotic_arch import (CausalLMOutputDexbotic, Decoder, DexboticConfig, IGNORE_INDEX, _Anchor, cfg_get, llm_specs)
from .pi0_arch import SiglipVisionTower, siglip_specs


class NaVILAConfig(DexboticConfig):
    """navila_arch.py:18-23."""
    model_type = "dexbotic_navila"

    def __init__(self, time_token_ids: Optional[List[int]] = None, soft_ce_std: float = 1.0, **kw):
        kw.setdefault("chat_template", "llama_3")
        kw.setdefault("mm_projector_type", "mlp_downsample")
        super().__init__(**kw)
        self.time_token_ids, self.soft_ce_std = time_token_ids, soft_ce_std


def downsample_projector_specs(d_vis: int, d: int, trainable: bool, prefix: str = "model.mm_projector.") -> list[ParamSpec]:
    """nn.Sequential(DownSampleBlock, LayerNorm(4C), Linear(4C, d), GELU, Linear(d, d)): indices 1, 2, 4 hold weights."""
    g = "projector"
    return [ParamSpec(prefix + "1.weight", (4 * d_vis,), g, trainable=trainable, no_decay=True),   # nn.LayerNorm
            ParamSpec(prefix + "1.bias", (4 * d_vis,), g, trainable=trainable),
            ParamSpec(prefix + "2.weight", (d, 4 * d_vis), g, trainable=trainable),
            ParamSpec(prefix + "2.bias", (d,), g, trainable=trainable),
            ParamSpec(prefix + "4.weight", (d, d), g, trainable=trainable),
            ParamSpec(prefix + "4.bias", (d,), g, trainable=trainable)]


def downsample_2x2(x2d: torch.Tensor, N: int, P: int) -> torch.Tensor:
    """DownSampleBlock (builder.py:9-33) on rows [N*P, C] -> [N*P', 4C], P' = ceil(h/2)^2: the (row-major) token grid
    is zero-padded to even sides and output token j*ceil(h/2)+i concatenates cells (2i,2j), (2i,2j+1), (2i+1,2j),
    (2i+1,2j+1).  Pure data movement of a few MB: torch view ops (their backward is the inverse gather)."""
    C = x2d.shape[-1]
    h = int(P ** 0.5)
    g = x2d.view(N, h, h, C)
    if h % 2:
        g = torch.nn.functional.pad(g, (0, 0, 0, 1, 0, 1))
    H2 = g.shape[1] // 2
    g = g.reshape(N, H2, 2, H2, 2, C).permute(0, 3, 1, 2, 4, 5)
    return g.reshape(N * H2 * H2, 4 * C).contiguous()


class NaVILAModel:
    """The `model.model` object (navila_arch.py:26-219): tower, projector, decoder."""
    mm_projector_prefix = "mm_projector"
    mm_vision_prefix = "mm_vision"

    def __init__(self, store, config: NaVILAConfig):
        self.store, self.config = store, config
        self.anchor = _Anchor(store.device)
        self.mm_vision_tower = SiglipVisionTower(store, config.mm_vision_tower, select_layer=-2)
        config.mm_hidden_size = self.mm_vision_tower.hidden_size
        self.llm = Decoder(store, config.llm_config)
        p = "model.mm_projector."
        self.proj_ln = Norm("ln", 1e-5, store.w(p + "1.weight"), store.w(p + "1.bias"), store.g(p + "1.weight"),
                            store.g(p + "1.bias"))
        self.proj = [Lin.of(store, p + "2.weight", p + "2.bias"), Lin.of(store, p + "4.weight", p + "4.bias")]
        h = int(self.mm_vision_tower.P ** 0.5)
        self.tokens_per_image = ((h + 1) // 2) ** 2

    backbone = property(lambda self: self.llm)
    mm_projector_module = property(lambda self: self.proj)
    mm_vision_module = property(lambda self: self.mm_vision_tower)

    def initialize_model(self, extra_config: dict):
        for key, value in extra_config.items():
            setattr(self.config, key, value)

    def refresh(self):
        self.mm_vision_tower.refresh()

    def encode_images(self, images: torch.Tensor) -> torch.Tensor:
        """navila_arch.py:37-40 (+ the 5-D flatten of dexbotic_arch.py:163-175): rows [n_images * P', d]."""
        images = images.reshape(-1, *images.shape[-3:])
        t = self.mm_vision_tower
        x = t.forward(self.anchor, images)
        x = downsample_2x2(x, images.shape[0], t.P)
        x = NormFn.apply(x, self.proj_ln, self.store)
        x = LinearFn.apply(x, self.proj[0], "gelu", self.store, True, None)
        return LinearFn.apply(x, self.proj[1], None, self.store, True, None)


class NaVILAForCausalLM(B200Module):
    """navila_arch.py:222-506 on the B200 kernels."""
    config_class = NaVILAConfig

    def __init__(self, config: NaVILAConfig, device="cuda"):
        super().__init__()
        self.config = config
        llm, vis = config.llm_config, config.mm_vision_tower
        if config.mm_projector_type != "mlp_downsample":
            raise ValueError("NaVILA is built with the mlp_downsample projector (navila_arch.py:22)")
        d, V = cfg_get(llm, "hidden_size"), cfg_get(llm, "vocab_size")
        specs = (llm_specs(llm, trainable=not config.freeze_llm)
                 + siglip_specs(vis, trainable=not config.freeze_mm_vision, select_layer=-2)
                 + downsample_projector_specs(cfg_get(vis, "hidden_size"), d, trainable=not config.freeze_mm_projector)
                 + [ParamSpec("lm_head.weight", (V, d), "lm_head", trainable=True, no_decay=False)])
        store = self._materialize(specs, device)
        self.model_engine = NaVILAModel(store, config)
        self.lm_head_lin = Lin.of(store, "lm_head.weight")

    def _after_weights_changed(self) -> None:
        self.model_engine.refresh()

    def _splice(self, input_ids, attention_mask, labels, images):
        """navila_arch.py:42-219.  Every <image> token of a row consumes the next tokens_per_image feature rows (the
        reference indexes the row's features by batch_idx and splits them evenly over the row's <image> tokens, :166-
        205 — the same thing whenever all rows carry the same number of image tokens, which is checked)."""
        eng, cfg = self.model_engine, self.config
        n_img = (input_ids == -200).sum(dim=1)
        feats = eng.encode_images(images)
        n_entries = feats.shape[0] // eng.tokens_per_image
        B = input_ids.shape[0]
        if n_entries % B != 0 or not bool((n_img == n_entries // B).all()):
            raise NotImplementedError("NaVILA splice: every row must carry one <image> token per frame of its sample")
        mask_u8 = None if attention_mask is None else attention_mask.to(torch.uint8).contiguous()
        ids = input_ids.contiguous()
        max_len = cfg.tokenizer_model_max_length or 0
        P = eng.tokens_per_image
        lengths = ops.splice_lengths(ids, mask_u8, P, max_len)
        S = int(lengths.max().item())
        src, new_labels, new_mask, pos = ops.splice_plan(ids, mask_u8, labels, P, max_len, S,
                                                         cfg.tokenizer_padding_side == "left")
        self.store.wait_chunk(0)              # the embedding table's update of the previous step
        emb = SpliceFn.apply(feats, src, eng.llm.embed_w, eng.llm.embed_g, self.store)
        return emb, new_labels, new_mask, pos, S

    def forward(self,
                input_ids: torch.LongTensor = None,
                images: Optional[torch.FloatTensor] = None,
                attention_mask: Optional[torch.Tensor] = None,
                position_ids: Optional[torch.LongTensor] = None,
                past_key_values: Optional[List[torch.FloatTensor]] = None,
                seqlens_in_batch: Optional[torch.LongTensor] = None,
                inputs_embeds: Optional[torch.FloatTensor] = None,
                labels: Optional[torch.LongTensor] = None,
                use_cache: Optional[bool] = None,
                output_attentions: Optional[bool] = None,
                output_hidden_states: Optional[bool] = None,
                return_dict: Optional[bool] = None,
                cache_position: Optional[torch.LongTensor] = None,
                image_masks: Optional[torch.BoolTensor] = None,
                **kwargs) -> CausalLMOutputDexbotic:
        if not input_ids.is_cuda:
            raise RuntimeError("dexbotic_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        cfg, eng = self.config, self.model_engine
        B = input_ids.shape[0]
        emb, new_labels, new_mask, pos, S = self._splice(input_ids, attention_mask, labels, images)
        hidden2d = eng.llm.forward(emb.view(B * S, -1), B, S, new_mask, pos)            # final-norm hidden states
        if labels is None:        # inference-style call: the full logits, as the reference returns them
            logits = LinearFn.apply(hidden2d, self.lm_head_lin, None, self.store, True, None)
            return CausalLMOutputDexbotic(loss=None, logits=logits.view(B, S, -1))
        # shifted targets: position s predicts label s+1; only target rows reach the lm_head
        tgt = new_labels[:, 1:]
        keep = (tgt != IGNORE_INDEX)
        bs = keep.nonzero(as_tuple=False)                                              # one host sync (row count)
        rows = (bs[:, 0] * S + bs[:, 1]).to(torch.int32).contiguous()
        targets = tgt[keep].contiguous()
        n = targets.numel()
        if n == 0:
            return CausalLMOutputDexbotic(loss=hidden2d.float().sum() * 0.0, logits=None)
        h_rows = GatherRowsFn.apply(hidden2d, rows)
        logits2d = LinearFn.apply(h_rows, self.lm_head_lin, None, self.store, True, None)   # [n, V]
        soft = cfg.time_token_ids if (self.training and cfg.time_token_ids) else None
        if not soft:
            loss = CrossEntropyFn.apply(logits2d, targets)                                   # mean over the n targets
        else:
            # navila/loss.py:53-70: hard targets as usual; a target that is a time token becomes a Gaussian over the
            # time-token ids.  The few soft rows take torch's fp32 log-softmax (a handful of [V] rows).
            soft_t = torch.tensor(list(soft), device=targets.device, dtype=targets.dtype)
            is_soft = torch.isin(targets, soft_t)
            hard_idx = (~is_soft).nonzero(as_tuple=False).flatten()
            soft_idx = is_soft.nonzero(as_tuple=False).flatten()
            total = logits2d.new_zeros((), dtype=torch.float32)
            if hard_idx.numel():
                lh = GatherRowsFn.apply(logits2d, hard_idx.to(torch.int32).contiguous())
                total = total + CrossEntropyFn.apply(lh, targets[hard_idx].contiguous()) * hard_idx.numel()
            if soft_idx.numel():
                ls = torch.log_softmax(logits2d[soft_idx].float(), dim=-1)
                dist = torch.exp(-((targets[soft_idx][:, None] - soft_t[None, :]) ** 2).float() / (2 * cfg.soft_ce_std ** 2))
                dist = dist / dist.sum(dim=1, keepdim=True)
                total = total - (ls[:, soft_t] * dist).sum()
            loss = total / n
        return CausalLMOutputDexbotic(loss=loss, logits=logits2d)

    def zero_grad(self, set_to_none: bool = False):
        self.store.zero_grad()

    def optimizer_step(self, base_lr: float = 2e-5, mm_projector_lr=None, mm_vision_lr=None, betas=(0.9, 0.999),
                       eps: float = 1e-8, weight_decay: float = 0.0, max_grad_norm=1.0):
        lrs = {"llm": base_lr, "projector": mm_projector_lr or base_lr, "vision": mm_vision_lr or base_lr,
               "lm_head": base_lr}
        norm = self.store.adamw_step(lrs, betas, eps, weight_decay, max_grad_norm)
        self.model_engine.refresh()
        return norm

    @property
    def mm_projector_prefix(self) -> str:
        return "model.mm_projector"

    @property
    def mm_vision_prefix(self) -> str:
        return "model.mm_vision_tower"
