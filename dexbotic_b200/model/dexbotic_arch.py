"""Mirror of dexbotic/model/dexbotic_arch.py (reference lines cited per method): the VLM shell —
vision tower -> mm_projector -> image-token splice -> LLM decoder — executed by the B200 kernels.

Parameter names / shapes are the reference's (model.llm.*, model.mm_vision_tower.*, model.mm_projector.*),
so reference checkpoints load with load_state_dict().
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Any, Optional

import torch

from .. import ops
from ..functional import (AttnEnv, BlockCfg, BlockW, GatherRowsFn, Lin, LinearFn, Norm, NormFn, SpliceFn,
                          TransformerBlockFn)
from ..params import ParamSpec, ParamStore

IGNORE_INDEX = -100        # dexbotic/constants.py
IMAGE_TOKEN_INDEX = -200   # dexbotic/constants.py


def cfg_get(cfg: Any, key: str, default=None):
    if cfg is None:
        return default
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


def rope_theta_of(cfg) -> float:
    t = cfg_get(cfg, "rope_theta")
    if t is None:
        rp = cfg_get(cfg, "rope_parameters") or {}
        t = rp.get("rope_theta") if isinstance(rp, dict) else None
    return float(t if t is not None else 10000.0)


@dataclass
class CausalLMOutputDexbotic:
    """dexbotic_arch.py:26-34 — same field names (the trainer logs every key ending in `_loss`)."""
    loss: Optional[torch.Tensor] = None
    logits: Optional[torch.Tensor] = None
    past_key_values: Any = None
    hidden_states: Any = None
    attentions: Any = None
    text_loss: Optional[torch.Tensor] = None
    action_loss: Optional[torch.Tensor] = None

    def __getitem__(self, k):
        return getattr(self, k) if isinstance(k, str) else tuple(v for v in self.__dict__.values() if v is not None)[k]

    def keys(self):
        return [k for k, v in self.__dict__.items() if v is not None]

    def __iter__(self):
        return iter(self.keys())


class DexboticConfig:
    """dexbotic_arch.py:17-23.  llm_config / mm_vision_tower are HF config objects or plain dicts."""
    model_type = "dexbotic"

    def __init__(self, llm_config=None, mm_projector_type: str = "mlp2x_gelu", mm_vision_tower=None,
                 chat_template: str = "dexbotic", init_llm_weights: bool = False, **kwargs):
        self.llm_config = llm_config
        self.mm_projector_type = mm_projector_type
        self.mm_vision_tower = mm_vision_tower
        self.chat_template = chat_template
        self.init_llm_weights = init_llm_weights
        self.tokenizer_model_max_length = kwargs.pop("tokenizer_model_max_length", None)
        self.tokenizer_padding_side = kwargs.pop("tokenizer_padding_side", "right")
        self.freeze_llm = kwargs.pop("freeze_llm", False)
        self.freeze_mm_projector = kwargs.pop("freeze_mm_projector", False)
        self.freeze_mm_vision = kwargs.pop("freeze_mm_vision", False)
        for k, v in kwargs.items():
            setattr(self, k, v)
        # _merge_llm (dexbotic_arch.py:80-86): expose the LLM's fields on the top-level config
        for key in ("vocab_size", "hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads",
                    "num_key_value_heads", "rms_norm_eps", "hidden_act", "max_position_embeddings"):
            if not hasattr(self, key) and cfg_get(llm_config, key) is not None:
                setattr(self, key, cfg_get(llm_config, key))


# ----------------------------------------------------------------------------------- specs
def llm_specs(cfg, trainable: bool, prefix: str = "model.llm.") -> list[ParamSpec]:
    d, V = cfg_get(cfg, "hidden_size"), cfg_get(cfg, "vocab_size")
    H, KVH = cfg_get(cfg, "num_attention_heads"), cfg_get(cfg, "num_key_value_heads") or cfg_get(cfg, "num_attention_heads")
    hd = cfg_get(cfg, "head_dim") or d // H
    inter = cfg_get(cfg, "intermediate_size")
    mt = cfg_get(cfg, "model_type", "qwen2")
    qkv_bias = bool(cfg_get(cfg, "attention_bias", mt == "qwen2")) or mt == "qwen2"
    sp = [ParamSpec(prefix + "embed_tokens.weight", (V, d), "llm", trainable=trainable, no_decay=False)]
    for i in range(cfg_get(cfg, "num_hidden_layers")):
        q = f"{prefix}layers.{i}."
        sp.append(ParamSpec(q + "input_layernorm.weight", (d,), "llm", trainable=trainable))
        sp.append(ParamSpec(q + "post_attention_layernorm.weight", (d,), "llm", trainable=trainable))
        if qkv_bias:
            for n, rows in (("q", H * hd), ("k", KVH * hd), ("v", KVH * hd)):
                sp.append(ParamSpec(f"{q}self_attn.{n}_proj.bias", (rows,), "llm", fuse=q + "qkvb", trainable=trainable))
        for n, rows in (("q", H * hd), ("k", KVH * hd), ("v", KVH * hd)):
            sp.append(ParamSpec(f"{q}self_attn.{n}_proj.weight", (rows, d), "llm", fuse=q + "qkvw", trainable=trainable))
        sp.append(ParamSpec(q + "self_attn.o_proj.weight", (d, H * hd), "llm", trainable=trainable))
        sp.append(ParamSpec(q + "mlp.gate_proj.weight", (inter, d), "llm", trainable=trainable))
        sp.append(ParamSpec(q + "mlp.up_proj.weight", (inter, d), "llm", trainable=trainable))
        sp.append(ParamSpec(q + "mlp.down_proj.weight", (d, inter), "llm", trainable=trainable))
    sp.append(ParamSpec(prefix + "norm.weight", (d,), "llm", trainable=trainable))
    return sp


def clip_specs(cfg, trainable: bool, prefix: str = "model.mm_vision_tower.vision_tower.vision_model.") -> list[ParamSpec]:
    D, inter, L = cfg_get(cfg, "hidden_size"), cfg_get(cfg, "intermediate_size"), cfg_get(cfg, "num_hidden_layers")
    ps, img = cfg_get(cfg, "patch_size"), cfg_get(cfg, "image_size")
    P = (img // ps) ** 2
    C = cfg_get(cfg, "num_channels", 3)
    g = "vision"
    sp = [ParamSpec(prefix + "embeddings.class_embedding", (D,), g, trainable=trainable),
          ParamSpec(prefix + "embeddings.patch_embedding.weight", (D, C, ps, ps), g, trainable=trainable),
          ParamSpec(prefix + "embeddings.position_embedding.weight", (P + 1, D), g, trainable=trainable, no_decay=False),
          ParamSpec(prefix + "pre_layrnorm.weight", (D,), g, trainable=trainable),
          ParamSpec(prefix + "pre_layrnorm.bias", (D,), g, trainable=trainable)]
    for i in range(L):
        q = f"{prefix}encoder.layers.{i}."
        used = trainable and i < L - 1      # select_layer=-2: the last layer never runs (clip_encoder.py:14,32)
        for n in ("layer_norm1", "layer_norm2"):
            sp.append(ParamSpec(f"{q}{n}.weight", (D,), g, trainable=used))
            sp.append(ParamSpec(f"{q}{n}.bias", (D,), g, trainable=used))
        for n in ("q", "k", "v"):
            sp.append(ParamSpec(f"{q}self_attn.{n}_proj.bias", (D,), g, fuse=q + "qkvb", trainable=used))
        for n in ("q", "k", "v"):
            sp.append(ParamSpec(f"{q}self_attn.{n}_proj.weight", (D, D), g, fuse=q + "qkvw", trainable=used))
        sp.append(ParamSpec(q + "self_attn.out_proj.weight", (D, D), g, trainable=used))
        sp.append(ParamSpec(q + "self_attn.out_proj.bias", (D,), g, trainable=used))
        sp.append(ParamSpec(q + "mlp.fc1.weight", (inter, D), g, trainable=used))
        sp.append(ParamSpec(q + "mlp.fc1.bias", (inter,), g, trainable=used))
        sp.append(ParamSpec(q + "mlp.fc2.weight", (D, inter), g, trainable=used))
        sp.append(ParamSpec(q + "mlp.fc2.bias", (D,), g, trainable=used))
    sp.append(ParamSpec(prefix + "post_layernorm.weight", (D,), g, trainable=False))
    sp.append(ParamSpec(prefix + "post_layernorm.bias", (D,), g, trainable=False))
    return sp


def projector_specs(kind: str, d_in: int, d_out: int, trainable: bool, prefix: str = "model.mm_projector.") -> list[ParamSpec]:
    """mm_projector/builder.py:37-81 (linear and mlpNx_gelu)."""
    g = "projector"
    if kind == "linear":
        return [ParamSpec(prefix + "weight", (d_out, d_in), g, trainable=trainable),
                ParamSpec(prefix + "bias", (d_out,), g, trainable=trainable)]
    import re
    m = re.match(r"^mlp(\d+)x_gelu$", kind)
    if not m:
        raise ValueError(f"Unknown projector type: {kind}")
    sp, din = [], d_in
    for i in range(int(m.group(1))):
        sp.append(ParamSpec(f"{prefix}{2 * i}.weight", (d_out, din), g, trainable=trainable))
        sp.append(ParamSpec(f"{prefix}{2 * i}.bias", (d_out,), g, trainable=trainable))
        din = d_out
    return sp


# ------------------------------------------------------------------------------- engines
class _Anchor:
    """A scalar that requires grad: Functions whose real inputs do not (images) take it as an input so
    autograd still runs their backward (parameter gradients are produced as a side effect)."""

    def __init__(self, device):
        self.t = torch.zeros((), device=device, dtype=torch.float32, requires_grad=True)


class PatchEmbedFn(torch.autograd.Function):
    """HF CLIPVisionEmbeddings + pre_layrnorm input: conv(stride=patch) as a GEMM over im2col rows, then
    [CLS | patches] + position embedding (called from clip_encoder.py:50-54)."""

    @staticmethod
    def forward(ctx, anchor, images, tower: "CLIPVisionTower"):
        B = images.shape[0]
        cols = ops.im2col_patches(images.contiguous(), tower.patch, tower.k_pad)
        patches = ops.gemm(cols, tower.patch_w_pad)                     # [B*P, D]
        out = ops.vit_embed_fwd(patches, tower.cls_w, tower.pos_w, B, tower.P)
        ctx.save_for_backward(cols)
        ctx.tower, ctx.B = tower, B
        return out.view(B * (tower.P + 1), tower.D)

    @staticmethod
    def backward(ctx, dout):
        (cols,) = ctx.saved_tensors
        t, B = ctx.tower, ctx.B
        store = t.store
        if t.g_patch is None:
            return None, None, None
        d_patches = torch.empty((B * t.P, t.D), device=dout.device, dtype=dout.dtype)
        d_cls = store.scratch_f32(t.D)
        d_pos = store.scratch_f32((t.P + 1) * t.D)
        ops.vit_embed_bwd(dout.contiguous().view(B, t.P + 1, t.D), d_patches, d_cls, d_pos, B, t.P)
        store.accumulate_small(d_cls, t.g_cls)
        store.accumulate_small(d_pos, t.g_pos)
        dw_pad = ops.gemm(d_patches, cols, a_mn=True, b_mn=True, out_dtype=torch.float32)   # [D, k_pad]
        K = t.g_patch.numel() // t.D
        ops.copy2d_(dw_pad, t.g_patch.view(t.D, K), t.D, K, accumulate=not store.first_write(t.g_patch))
        return None, None, None


class CLIPVisionTower:
    """modules/mm_vision/clip/clip_encoder.py — forward returns hidden_states[-2][:, 1:] semantics: the rows of
    every image are [CLS, patch_0..patch_{P-1}]; consumers skip the CLS row (feature_select, :31-36)."""

    def __init__(self, store: ParamStore, cfg, prefix: str = "model.mm_vision_tower.vision_tower.vision_model."):
        self.store, self.cfg, self.prefix = store, cfg, prefix
        self.D = cfg_get(cfg, "hidden_size")
        self.patch = cfg_get(cfg, "patch_size")
        self.image_size = cfg_get(cfg, "image_size")
        self.P = (self.image_size // self.patch) ** 2
        self.C = cfg_get(cfg, "num_channels", 3)
        self.heads = cfg_get(cfg, "num_attention_heads")
        self.eps = cfg_get(cfg, "layer_norm_eps", 1e-5)
        self.act = cfg_get(cfg, "hidden_act", "quick_gelu")
        K = self.C * self.patch * self.patch
        self.k_pad = (K + 7) // 8 * 8
        p = prefix
        self.cls_w, self.g_cls = store.w(p + "embeddings.class_embedding"), store.g(p + "embeddings.class_embedding")
        self.pos_w = store.w(p + "embeddings.position_embedding.weight")
        self.g_pos = store.g(p + "embeddings.position_embedding.weight")
        self.g_patch = store.g(p + "embeddings.patch_embedding.weight")
        self.patch_w_pad = torch.zeros((self.D, self.k_pad), device=store.device, dtype=torch.bfloat16)
        self.pre_ln = Norm("ln", self.eps, store.w(p + "pre_layrnorm.weight"), store.w(p + "pre_layrnorm.bias"),
                           store.g(p + "pre_layrnorm.weight"), store.g(p + "pre_layrnorm.bias"))
        L = cfg_get(cfg, "num_hidden_layers")
        bc = BlockCfg(d=self.D, heads=self.heads, kv_heads=self.heads, head_dim=self.D // self.heads,
                      inter=cfg_get(cfg, "intermediate_size"), mlp="mlp", act=self.act, rope=False)
        self.keep_layers = None          # blocks that keep intermediates (None: planned from free HBM at first use)
        self.decoder_reserve_gb = 0.0    # set by the VLM shell: HBM the decoder's kept activations will need
        self.blocks = []
        for i in range(L - 1):                      # select_layer = -2
            q = f"{p}encoder.layers.{i}."
            self.blocks.append(BlockW(
                cfg=bc,
                norm1=Norm("ln", self.eps, store.w(q + "layer_norm1.weight"), store.w(q + "layer_norm1.bias"),
                           store.g(q + "layer_norm1.weight"), store.g(q + "layer_norm1.bias")),
                qkv=Lin.of(store, [f"{q}self_attn.{n}_proj.weight" for n in "qkv"],
                           [f"{q}self_attn.{n}_proj.bias" for n in "qkv"]),
                o=Lin.of(store, q + "self_attn.out_proj.weight", q + "self_attn.out_proj.bias"),
                norm2=Norm("ln", self.eps, store.w(q + "layer_norm2.weight"), store.w(q + "layer_norm2.bias"),
                           store.g(q + "layer_norm2.weight"), store.g(q + "layer_norm2.bias")),
                fc1=Lin.of(store, q + "mlp.fc1.weight", q + "mlp.fc1.bias"),
                fc2=Lin.of(store, q + "mlp.fc2.weight", q + "mlp.fc2.bias")))

    @property
    def hidden_size(self) -> int:
        return self.D

    @property
    def num_patches(self) -> int:
        return self.P

    def refresh(self) -> None:
        """bf16 K-padded copy of the conv weight (588 -> 592 columns: TMA rows must be 16-byte multiples)."""
        K = self.C * self.patch * self.patch
        # read the compute copy (bf16 shadow: complete on every rank), not the fp32 master — under ZeRO-1 a rank's
        # master holds current values only for the pieces it owns
        w = self.store.w(self.prefix + "embeddings.patch_embedding.weight").view(self.D, K)
        ops.copy2d_(w, self.patch_w_pad, self.D, K)

    def forward(self, anchor: _Anchor, images: torch.Tensor) -> torch.Tensor:
        """images [N,3,H,W] -> [N*(P+1), D] (row 0 of every image is the CLS token)."""
        N = images.shape[0]
        x = PatchEmbedFn.apply(anchor.t, images, self)
        x = NormFn.apply(x, self.pre_ln, self.store)
        env = AttnEnv(B=N, S=self.P + 1)
        if self.keep_layers is None and torch.is_grad_enabled():
            # the decoder plans first (it is 95 % of the FLOPs); the tower keeps what still fits
            self.keep_layers = plan_keep_layers(self.store, self.blocks[0].cfg, len(self.blocks), N, self.P + 1,
                                                x.device, reserve_gb=14.0 + self.decoder_reserve_gb)
        keep = self.keep_layers if torch.is_grad_enabled() else 0
        for i, bw in enumerate(self.blocks):
            x = TransformerBlockFn.apply(x, bw, env, self.store, i >= (keep or 0))
        return x


class Decoder:
    """HF Qwen2Model / LlamaModel as instantiated by AutoModel.from_config at dexbotic_arch.py:55-62."""

    def __init__(self, store: ParamStore, cfg, prefix: str = "model.llm."):
        self.store, self.cfg, self.prefix = store, cfg, prefix
        d, H = cfg_get(cfg, "hidden_size"), cfg_get(cfg, "num_attention_heads")
        KVH = cfg_get(cfg, "num_key_value_heads") or H
        self.hd = cfg_get(cfg, "head_dim") or d // H
        self.d = d
        eps = cfg_get(cfg, "rms_norm_eps", 1e-6)
        mt = cfg_get(cfg, "model_type", "qwen2")
        kind = "rms1p" if mt.startswith("gemma") else "rms"
        act = cfg_get(cfg, "hidden_act") or cfg_get(cfg, "hidden_activation") or "silu"
        bc = BlockCfg(d=d, heads=H, kv_heads=KVH, head_dim=self.hd, inter=cfg_get(cfg, "intermediate_size"),
                      mlp="glu", act=act, rope=True)
        self.embed_w = store.w(prefix + "embed_tokens.weight")
        self.embed_g = store.g(prefix + "embed_tokens.weight")
        if self.embed_g is not None:
            store.mark_sparse_grad(prefix + "embed_tokens.weight")
        self.blocks = []
        for i in range(cfg_get(cfg, "num_hidden_layers")):
            q = f"{prefix}layers.{i}."
            has_b = (q + "self_attn.q_proj.bias") in store.slots
            self.blocks.append(BlockW(
                cfg=bc,
                norm1=Norm(kind, eps, store.w(q + "input_layernorm.weight"), None, store.g(q + "input_layernorm.weight")),
                qkv=Lin.of(store, [f"{q}self_attn.{n}_proj.weight" for n in "qkv"],
                           [f"{q}self_attn.{n}_proj.bias" for n in "qkv"] if has_b else None),
                o=Lin.of(store, q + "self_attn.o_proj.weight"),
                norm2=Norm(kind, eps, store.w(q + "post_attention_layernorm.weight"), None,
                           store.g(q + "post_attention_layernorm.weight")),
                gate=Lin.of(store, q + "mlp.gate_proj.weight"), up=Lin.of(store, q + "mlp.up_proj.weight"),
                down=Lin.of(store, q + "mlp.down_proj.weight")))
        for i, bw in enumerate(self.blocks):
            q = f"{prefix}layers.{i}."
            bw.grad_range = store.grad_range([n for n in store.order if n.startswith(q)])
        self.final_norm = Norm(kind, eps, store.w(prefix + "norm.weight"), None, store.g(prefix + "norm.weight"))
        # optimizer / forward overlap: chunk 0 = the embedding table (first read at the image-token splice, after the
        # vision tower has run), chunk i + 1 = decoder block i
        self.embed_chunk = store.grad_range([prefix + "embed_tokens.weight"]) if self.embed_g is not None else None
        store.set_param_chunks([self.embed_chunk] + [bw.grad_range for bw in self.blocks])
        self.theta = rope_theta_of(cfg)
        self._rope_cache = None
        # layers [0, keep_layers) keep their intermediates (no recompute in backward); the rest recompute.
        # None = decide from free HBM at the first forward (1 full-recompute layer costs ~25 % more GEMM work).
        self.keep_layers = None

    def rope_tables(self, n_pos: int, device):
        """cos/sin exactly as HF's rotary embedding computes them (fp32 inv_freq, fp32 outer product)."""
        if self._rope_cache is None or self._rope_cache[0].shape[0] < n_pos:
            n = max(n_pos, 1024)
            inv = 1.0 / (self.theta ** (torch.arange(0, self.hd, 2, dtype=torch.float32, device=device) / self.hd))
            f = torch.arange(n, dtype=torch.float32, device=device)[:, None] * inv[None, :]
            self._rope_cache = (f.cos().contiguous(), f.sin().contiguous())
        return self._rope_cache

    def forward(self, x2d, B: int, S: int, mask_u8, pos_i32):
        dev = x2d.device
        self._last_B = B
        cos, sin = self.rope_tables(S + 1, dev)
        env = AttnEnv(B=B, S=S, keymask=mask_u8, causal=True, pos=pos_i32.reshape(-1), cos=cos, sin=sin)
        keep = self._decide_keep_layers(x2d) if torch.is_grad_enabled() else 0
        for i, bw in enumerate(self.blocks):
            self.store.wait_chunk(i + 1)    # block i's AdamW update of the previous step (ParamStore.async_optimizer)
            x2d = TransformerBlockFn.apply(x2d, bw, env, self.store, i >= keep)
        return NormFn.apply(x2d, self.final_norm, self.store)

    def _decide_keep_layers(self, x2d) -> int:
        if self.keep_layers is None:
            B = self._last_B
            self.keep_layers = plan_keep_layers(self.store, self.blocks[0].cfg, len(self.blocks), B, x2d.shape[0] // B,
                                                x2d.device)
        return self.keep_layers


def kept_bytes_per_block(bc: BlockCfg, B: int, S: int, es: int = 2) -> int:
    """Bytes a block keeps alive when it is NOT recomputed: qkv, attn, x1, the MLP pre-activations (GLU: g, u and
    act(g) * u, which the fused down-projection backward reads instead of recomputing it) and the attention statistics
    (bf16: one fp32 log-sum-exp per row — flash attention; fp32 blocks keep the probabilities)."""
    M = B * S
    W = (bc.heads + 2 * bc.kv_heads) * bc.head_dim
    inter = 3 * bc.inter if bc.mlp == "glu" else bc.inter
    stats = 4 * B * bc.heads * S if es == 2 else es * B * bc.heads * S * ((S + 7) // 8 * 8)
    return es * M * (W + bc.heads * bc.head_dim + bc.d + inter) + stats


def plan_keep_layers(store: ParamStore, bc: BlockCfg, n_layers: int, B: int, S: int, device, reserve_gb: float = 14.0) -> int:
    """How many blocks can keep their intermediates instead of being recomputed, from the HBM that is free
    right now (minus Adam moments still to be allocated, backward transients and a fragmentation reserve)."""
    if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
        reserve_gb += 6.0        # NCCL channel / NVLS buffers grow with the number of peers and are allocated lazily
    free, _ = torch.cuda.mem_get_info(device)
    cached = torch.cuda.memory_reserved(device) - torch.cuda.memory_allocated(device)
    moments = 0 if store.exp_avg is not None else 8 * store.n_train
    budget = free + cached - moments - int(reserve_gb * (1 << 30))
    per = int(kept_bytes_per_block(bc, B, S) * 1.08) + (64 << 20)
    return int(max(0, min(n_layers, budget // per)))


class DexboticVLMModel:
    """dexbotic_arch.py:51-373 (the `model.model` object of every *ForCausalLM)."""

    mm_projector_prefix = "mm_projector"     # dexbotic_arch.py:129-131
    mm_vision_prefix = "mm_vision"           # :137-139

    def __init__(self, store: ParamStore, config: DexboticConfig):
        self.store, self.config = store, config
        self.anchor = _Anchor(store.device)
        self.mm_vision_tower = CLIPVisionTower(store, config.mm_vision_tower)
        config.mm_hidden_size = self.mm_vision_tower.hidden_size           # :114
        self.llm = Decoder(store, config.llm_config)
        self.proj_kind = config.mm_projector_type
        p = "model.mm_projector."
        if self.proj_kind == "linear":
            self.proj = [Lin.of(store, p + "weight", p + "bias")]
        else:
            n = len([k for k in store.slots if k.startswith(p) and k.endswith(".weight")])
            self.proj = [Lin.of(store, f"{p}{2 * i}.weight", f"{p}{2 * i}.bias") for i in range(n)]

    @property
    def backbone(self):                       # dexbotic_arch.py:153-155
        return self.llm

    @property
    def mm_projector_module(self):            # :125-127
        return self.proj

    @property
    def mm_vision_module(self):               # :133-135
        return self.mm_vision_tower

    def initialize_model(self, extra_config: dict):
        """dexbotic_arch.py:122-129.  The reference (re)builds the vision tower / projector / action head here for
        checkpoints that lack them; this backend always builds every module in __init__, so only the config moves."""
        for key, value in extra_config.items():
            setattr(self.config, key, value)

    def refresh(self):
        self.mm_vision_tower.refresh()

    def _extract_vision_features(self, images: torch.Tensor):
        """dexbotic_arch.py:157-180.  Returns (features2d [n_img*(P+1), D_llm], rows_per_entry, views)."""
        views = 1
        if images.dim() == 5:                                   # [B, n_image, C, H, W] -> views share a sample
            views = images.shape[1]
            images = images.reshape(-1, *images.shape[2:])
        tower = self.mm_vision_tower
        if tower.g_patch is None:                                # frozen tower: no backward through it
            with torch.no_grad():
                x = tower.forward(self.anchor, images)
        else:
            x = tower.forward(self.anchor, images)
        for i, lin in enumerate(self.proj):                      # Linear (-> GELU -> Linear)*
            act = "gelu" if i < len(self.proj) - 1 else None
            x = LinearFn.apply(x, lin, act, self.store, tower.g_patch is not None, self.anchor.t if i == 0 else None)
        return x, views

    def _prepare_inputs_labels_for_multimodal(self, input_ids, attention_mask, labels, images, append_tokens: int = 0,
                                              append_token_id: int = 1):
        """dexbotic_arch.py:182-259 + :261-373, as three kernels (lengths, plan, gather) and ONE 4-byte
        device->host read (the padded length), instead of a per-sample Python loop with >= 3 syncs each.

        append_tokens > 0 additionally inserts that many copies of embed_tokens[append_token_id] right after the
        last valid token of every sample (OFT's action placeholders: oft_arch.py:169-201 insert_action_embedding
        applied to embed_tokens(ones), oft_discrete_arch.py:125-130) — folded into the same gather.
        Returns (inputs_embeds [B,S,D], labels, mask u8 [B,S], position ids i32 [B,S], S, lengths i32 [B])."""
        cfg = self.config
        P = self.mm_vision_tower.P
        B, L = input_ids.shape
        if self.mm_vision_tower.keep_layers is None:      # leave room for the decoder's kept activations first
            n_views = images.shape[1] if images.dim() == 5 else 1
            need = kept_bytes_per_block(self.llm.blocks[0].cfg, B, L - 1 + P * n_views + append_tokens) * 1.08 * len(self.llm.blocks)
            self.mm_vision_tower.decoder_reserve_gb = need / (1 << 30)
        feats, views = self._extract_vision_features(images)
        self.last_image_features = feats      # projector output rows [n_img*(P+1), D] (MemVLA reads it back)
        mask_u8 = None if attention_mask is None else attention_mask.to(torch.uint8).contiguous()
        ids = input_ids.contiguous()
        max_len = cfg.tokenizer_model_max_length or 0
        # 5-D images: the n views of a sample form ONE image entry of n*P tokens (dexbotic_arch.py:163-175)
        P_entry = P * views
        lengths = ops.splice_lengths(ids, mask_u8, P_entry, max_len)
        # the padded length is the one device->host read of this path; `static_seq_len` (a config field the caller sets
        # to an upper bound, e.g. L - n_image_tokens + n_image_tokens * P) removes it: rows are padded up to it and the
        # whole step becomes CUDA-graph capturable
        S = int(getattr(cfg, "static_seq_len", 0) or 0) or int(lengths.max().item())
        left = cfg.tokenizer_padding_side == "left"
        src, new_labels, new_mask, pos = ops.splice_plan(ids, mask_u8, labels, P_entry, max_len, S, left)
        # feature rows carry a CLS row per image view: token j of the dense numbering lives at row j + j // P + 1
        src = _shift_image_rows(src, P)
        if append_tokens > 0:
            if left:
                raise NotImplementedError("action-token insertion assumes right padding (oft_arch.py:188-199)")
            A = append_tokens
            idx = torch.arange(S + A, device=src.device, dtype=torch.int32)[None, :]
            ln = lengths[:, None]
            pad = torch.full((B, A), -(2 ** 31), device=src.device, dtype=torch.int32)
            src_ext = torch.cat([src, pad], dim=1)
            shifted = torch.cat([pad, src], dim=1)              # src[b, s - A] for the tail
            # append_token_id: one id for all appended rows, None = zero rows (the caller adds its own embeddings
            # there: OFT's action_query), or a list of per-position ids / None (proprio token + placeholders)
            PAD = -(2 ** 31)
            ids_row = append_token_id if isinstance(append_token_id, (list, tuple)) else [append_token_id] * A
            assert len(ids_row) == A
            fill_row = torch.tensor([PAD if t is None else int(t) for t in ids_row], device=src.device, dtype=torch.int32)
            fill = fill_row[(idx - ln).clamp(0, A - 1).long()]
            src = torch.where(idx < ln, src_ext, torch.where(idx < ln + A, fill, shifted)).contiguous()
            new_mask = (idx < ln + A).to(torch.uint8).contiguous()
            pos = idx.expand(B, S + A).contiguous()             # HF default position ids: arange (position_ids=None)
            new_labels = torch.cat([new_labels, torch.full((B, A), IGNORE_INDEX, device=src.device,
                                                           dtype=new_labels.dtype)], dim=1)
            S = S + A
        self.store.wait_chunk(0)              # the embedding table's update of the previous step
        emb = SpliceFn.apply(feats, src, self.llm.embed_w, self.llm.embed_g, self.store)
        return emb, new_labels, new_mask, pos, S, lengths


def _shift_image_rows(src: torch.Tensor, P: int) -> torch.Tensor:
    """plan rows index a dense [n_entries*P] feature matrix; ours has P+1 rows per entry (CLS first)."""
    img = (src < 0) & (src != -(2 ** 31))
    j = -1 - src
    shifted = -1 - (j + j // P + 1)
    return torch.where(img, shifted, src).to(torch.int32)
