"""Mirror of dexbotic/model/pi0/pi0_arch.py: Pi0Config / Pi0ForCausalLM — SigLIP tower,
linear projector, Gemma LLM + Gemma action expert executed layer by layer with SHARED attention over the joint
[prefix | suffix] sequence (`_inner_forward_mot`, pi0_arch.py:116-228), flow-matching loss (:337-388).

State-dict keys == the reference's (model.llm.*, model.action_expert.*, model.mm_vision_tower.*,
model.mm_projector.*, model.state_proj.*, model.action_in_proj.*, model.action_time_mlp_{in,out}.*,
model.action_out_proj.*).  The block-causal mask (make_attn_mask, :22-33) is evaluated inside the softmax kernel
from two small per-token arrays (validity, cumsum(ar_mask)) instead of a dense fp32 [B,1,S,S] tensor.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch

from .. import ops
from ..functional import (AttnEnv, BlockCfg, BlockW, CastFn, Lin, LinearFn, MSELossFn, Norm, NormFn,
                          TransformerBlockFn, linear_dgrad, linear_fwd, linear_wgrad, norm_bwd, norm_fwd)
from ..params import ParamSpec, ParamStore
from ._module import B200Module
from .dexbotic_arch import CausalLMOutputDexbotic, DexboticConfig, _Anchor, cfg_get, rope_theta_of


class Pi0Config(DexboticConfig):
    """pi0_arch.py:53-83 (llm_config / action_config / vision_config: HF config objects or dicts)."""
    model_type = "dexbotic_pi0"

    def __init__(self, vision_config=None, processor_config=None, action_config=None, action_dim: int = 32,
                 chunk_size: int = 50, **kw):
        kw.setdefault("mm_projector_type", "linear")
        super().__init__(**kw)
        self.vision_config, self.processor_config, self.action_config = vision_config, processor_config, action_config
        self.action_dim, self.chunk_size = action_dim, chunk_size


# ------------------------------------------------------------------------------------ specs
def siglip_specs(cfg, trainable: bool, prefix: str = "model.mm_vision_tower.vision_tower.vision_model.",
                 select_layer=None) -> list[ParamSpec]:
    """select_layer=-2 (the tower's default outside pi0, siglip_encoder.py:13): the last encoder layer and the
    post_layernorm never reach the output — the reference leaves them without gradient, here they are frozen."""
    D, inter, L = cfg_get(cfg, "hidden_size"), cfg_get(cfg, "intermediate_size"), cfg_get(cfg, "num_hidden_layers")
    n_used = L if select_layer is None else L + 1 + select_layer
    all_trainable = trainable
    ps, img, C = cfg_get(cfg, "patch_size"), cfg_get(cfg, "image_size"), cfg_get(cfg, "num_channels", 3)
    P = (img // ps) ** 2
    g = "vision"
    sp = [ParamSpec(prefix + "embeddings.patch_embedding.weight", (D, C, ps, ps), g, trainable=trainable),
          ParamSpec(prefix + "embeddings.patch_embedding.bias", (D,), g, trainable=trainable),
          ParamSpec(prefix + "embeddings.position_embedding.weight", (P, D), g, trainable=trainable, no_decay=False)]
    for i in range(L):
        q = f"{prefix}encoder.layers.{i}."
        trainable = all_trainable and i < n_used
        for n in ("layer_norm1", "layer_norm2"):
            sp.append(ParamSpec(f"{q}{n}.weight", (D,), g, trainable=trainable))
            sp.append(ParamSpec(f"{q}{n}.bias", (D,), g, trainable=trainable))
        for n in "qkv":
            sp.append(ParamSpec(f"{q}self_attn.{n}_proj.bias", (D,), g, fuse=q + "qkvb", trainable=trainable))
        for n in "qkv":
            sp.append(ParamSpec(f"{q}self_attn.{n}_proj.weight", (D, D), g, fuse=q + "qkvw", trainable=trainable))
        sp += [ParamSpec(q + "self_attn.out_proj.weight", (D, D), g, trainable=trainable),
               ParamSpec(q + "self_attn.out_proj.bias", (D,), g, trainable=trainable),
               ParamSpec(q + "mlp.fc1.weight", (inter, D), g, trainable=trainable),
               ParamSpec(q + "mlp.fc1.bias", (inter,), g, trainable=trainable),
               ParamSpec(q + "mlp.fc2.weight", (D, inter), g, trainable=trainable),
               ParamSpec(q + "mlp.fc2.bias", (D,), g, trainable=trainable)]
    trainable = all_trainable and select_layer is None
    sp += [ParamSpec(prefix + "post_layernorm.weight", (D,), g, trainable=trainable),
           ParamSpec(prefix + "post_layernorm.bias", (D,), g, trainable=trainable)]
    # multi-head attention pooling head: part of HF SiglipVisionModel's state dict, never on this path
    h = prefix + "head."
    sp += [ParamSpec(h + "probe", (1, 1, D), g, trainable=False),
           ParamSpec(h + "attention.in_proj_weight", (3 * D, D), g, trainable=False),
           ParamSpec(h + "attention.in_proj_bias", (3 * D,), g, trainable=False),
           ParamSpec(h + "attention.out_proj.weight", (D, D), g, trainable=False),
           ParamSpec(h + "attention.out_proj.bias", (D,), g, trainable=False),
           ParamSpec(h + "layernorm.weight", (D,), g, trainable=False),
           ParamSpec(h + "layernorm.bias", (D,), g, trainable=False),
           ParamSpec(h + "mlp.fc1.weight", (inter, D), g, trainable=False),
           ParamSpec(h + "mlp.fc1.bias", (inter,), g, trainable=False),
           ParamSpec(h + "mlp.fc2.weight", (D, inter), g, trainable=False),
           ParamSpec(h + "mlp.fc2.bias", (D,), g, trainable=False)]
    return sp


def gemma_specs(cfg, prefix: str, group: str, trainable: bool, embed_trainable: bool, skip_tail_of_last: bool):
    """Gemma decoder parameters.  skip_tail_of_last: the prefix stream's last-layer o_proj / MLP / post-norm and the
    final norm never influence the loss (only suffix_out is read, pi0_arch.py:386) -> frozen region, like the
    reference where they simply receive no gradient."""
    d, V = cfg_get(cfg, "hidden_size"), cfg_get(cfg, "vocab_size")
    H, KVH = cfg_get(cfg, "num_attention_heads"), cfg_get(cfg, "num_key_value_heads")
    hd = cfg_get(cfg, "head_dim") or d // H
    inter, L = cfg_get(cfg, "intermediate_size"), cfg_get(cfg, "num_hidden_layers")
    sp = [ParamSpec(prefix + "embed_tokens.weight", (V, d), group, trainable=embed_trainable, no_decay=False)]
    for i in range(L):
        q = f"{prefix}layers.{i}."
        tail = trainable and not (skip_tail_of_last and i == L - 1)
        sp.append(ParamSpec(q + "input_layernorm.weight", (d,), group, trainable=trainable))
        sp.append(ParamSpec(q + "post_attention_layernorm.weight", (d,), group, trainable=tail))
        for n, rows in (("q", H * hd), ("k", KVH * hd), ("v", KVH * hd)):
            sp.append(ParamSpec(f"{q}self_attn.{n}_proj.weight", (rows, d), group, fuse=q + "qkvw", trainable=trainable))
        sp.append(ParamSpec(q + "self_attn.o_proj.weight", (d, H * hd), group, trainable=tail))
        sp.append(ParamSpec(q + "mlp.gate_proj.weight", (inter, d), group, trainable=tail))
        sp.append(ParamSpec(q + "mlp.up_proj.weight", (inter, d), group, trainable=tail))
        sp.append(ParamSpec(q + "mlp.down_proj.weight", (d, inter), group, trainable=tail))
    sp.append(ParamSpec(prefix + "norm.weight", (d,), group, trainable=trainable and not skip_tail_of_last))
    return sp


# ----------------------------------------------------------------------------------- SigLIP
class SiglipEmbedFn(torch.autograd.Function):
    """HF SiglipVisionEmbeddings: conv(stride=patch, +bias) as im2col GEMM, + learned position embedding."""

    @staticmethod
    def forward(ctx, anchor, images, tower: "SiglipVisionTower"):
        N = images.shape[0]
        cols = ops.im2col_patches(images.contiguous(), tower.patch, tower.k_pad)
        patches = ops.gemm(cols, tower.patch_w_pad, bias=tower.patch_b)
        out = ops.add_pos_fwd(patches, tower.pos_w, N, tower.P)
        ctx.save_for_backward(cols)
        ctx.tower, ctx.N = tower, N
        return out

    @staticmethod
    def backward(ctx, dout):
        (cols,) = ctx.saved_tensors
        t, N = ctx.tower, ctx.N
        st = t.store
        if t.g_patch is None:
            return None, None, None
        dout = dout.contiguous()
        d_pos = st.scratch_f32(t.P * t.D)
        ops.add_pos_bwd(dout, d_pos, N, t.P)
        st.accumulate_small(d_pos, t.g_pos)
        sc = st.scratch_f32(t.D)
        ops.colsum_(dout, sc)
        st.accumulate_small(sc, t.g_patch_b)
        dw_pad = ops.gemm(dout, cols, a_mn=True, b_mn=True, out_dtype=torch.float32)
        K = t.g_patch.numel() // t.D
        ops.copy2d_(dw_pad, t.g_patch.view(t.D, K), t.D, K, accumulate=not st.first_write(t.g_patch))
        return None, None, None


class SiglipVisionTower:
    """modules/mm_vision/siglip/siglip_encoder.py with select_layer=None: last_hidden_state (post-LN), no CLS."""

    def __init__(self, store: ParamStore, cfg, prefix: str = "model.mm_vision_tower.vision_tower.vision_model.",
                 select_layer=None):
        self.store, self.cfg, self.prefix = store, cfg, prefix
        self.select_layer = select_layer     # None: last_hidden_state (post-LN); -2: output of the second-to-last layer
        self.D, self.patch = cfg_get(cfg, "hidden_size"), cfg_get(cfg, "patch_size")
        self.P = (cfg_get(cfg, "image_size") // self.patch) ** 2
        self.C = cfg_get(cfg, "num_channels", 3)
        heads, eps = cfg_get(cfg, "num_attention_heads"), cfg_get(cfg, "layer_norm_eps", 1e-6)
        K = self.C * self.patch * self.patch
        self.k_pad = (K + 7) // 8 * 8
        p = prefix
        self.patch_w_pad = torch.zeros((self.D, self.k_pad), device=store.device, dtype=torch.bfloat16)
        self.g_patch = store.g(p + "embeddings.patch_embedding.weight")
        self.patch_b, self.g_patch_b = store.w(p + "embeddings.patch_embedding.bias"), store.g(p + "embeddings.patch_embedding.bias")
        self.pos_w, self.g_pos = store.w(p + "embeddings.position_embedding.weight"), store.g(p + "embeddings.position_embedding.weight")
        bc = BlockCfg(d=self.D, heads=heads, kv_heads=heads, head_dim=self.D // heads,
                      inter=cfg_get(cfg, "intermediate_size"), mlp="mlp",
                      act=cfg_get(cfg, "hidden_act", "gelu_pytorch_tanh"), rope=False)
        ln = lambda n: Norm("ln", eps, store.w(n + ".weight"), store.w(n + ".bias"), store.g(n + ".weight"), store.g(n + ".bias"))  # noqa: E731
        self.blocks = []
        for i in range(cfg_get(cfg, "num_hidden_layers")):
            q = f"{p}encoder.layers.{i}."
            self.blocks.append(BlockW(cfg=bc, norm1=ln(q + "layer_norm1"),
                                      qkv=Lin.of(store, [f"{q}self_attn.{n}_proj.weight" for n in "qkv"],
                                                 [f"{q}self_attn.{n}_proj.bias" for n in "qkv"]),
                                      o=Lin.of(store, q + "self_attn.out_proj.weight", q + "self_attn.out_proj.bias"),
                                      norm2=ln(q + "layer_norm2"),
                                      fc1=Lin.of(store, q + "mlp.fc1.weight", q + "mlp.fc1.bias"),
                                      fc2=Lin.of(store, q + "mlp.fc2.weight", q + "mlp.fc2.bias")))
        self.post_ln = ln(p + "post_layernorm")

    @property
    def hidden_size(self):
        return self.D

    def refresh(self):
        K = self.C * self.patch * self.patch
        # read the compute copy (bf16 shadow: complete on every rank), not the fp32 master — under ZeRO-1 a rank's
        # master holds current values only for the pieces it owns
        w = self.store.w(self.prefix + "embeddings.patch_embedding.weight").view(self.D, K)
        ops.copy2d_(w, self.patch_w_pad, self.D, K)

    def forward(self, anchor: _Anchor, images: torch.Tensor) -> torch.Tensor:
        """[N,3,H,W] -> [N*P, D]"""
        N = images.shape[0]
        x = SiglipEmbedFn.apply(anchor.t, images, self)
        env = AttnEnv(B=N, S=self.P)
        n_used = len(self.blocks) if self.select_layer is None else len(self.blocks) + 1 + self.select_layer
        for bw in self.blocks[:n_used]:
            x = TransformerBlockFn.apply(x, bw, env, self.store, False)
        if self.select_layer is not None:
            return x
        return NormFn.apply(x, self.post_ln, self.store)


class MoTEngine:
    """What `model.model` answers for pi0 / pi0.5 (Pi0Model inherits DexboticVLMModel's accessors, dexbotic_arch.py:
    122-155): the properties the reference's optimizer grouping and freeze logic read (base_exp.py:95-203, 318-330)."""
    mm_projector_prefix = "mm_projector"
    mm_vision_prefix = "mm_vision"

    def __init__(self, owner):
        self._owner = owner

    backbone = property(lambda self: self._owner.layers)
    mm_projector_module = property(lambda self: self._owner.proj)
    mm_vision_module = property(lambda self: self._owner.tower)

    def initialize_model(self, extra_config: dict):
        for key, value in extra_config.items():
            setattr(self._owner.config, key, value)

    def refresh(self):
        self._owner.tower.refresh()


# ------------------------------------------------------------------- mixture of transformers
@dataclass
class StreamW:
    d: int
    inter: int
    norm1: Norm
    qkv: Lin
    o: Lin
    norm2: Norm
    gate: Lin
    up: Lin
    down: Lin


@dataclass
class MoTEnv:
    B: int
    lens: tuple            # (Sp, Ss)
    heads: int
    kv_heads: int
    head_dim: int
    act: str
    keymask: torch.Tensor  # uint8 [B, S]
    bid: torch.Tensor      # int32 [B, S]  cumsum(ar_mask)
    pos: torch.Tensor      # int32 [B*S]
    cos: torch.Tensor
    sin: torch.Tensor


def _stream_w(store: ParamStore, cfg, prefix: str, i: int) -> StreamW:
    q = f"{prefix}layers.{i}."
    eps = cfg_get(cfg, "rms_norm_eps", 1e-6)
    rn = lambda n: Norm("rms1p", eps, store.w(n), None, store.g(n))  # noqa: E731  (GemmaRMSNorm: x * (1 + w))
    return StreamW(d=cfg_get(cfg, "hidden_size"), inter=cfg_get(cfg, "intermediate_size"),
                   norm1=rn(q + "input_layernorm.weight"),
                   qkv=Lin.of(store, [f"{q}self_attn.{n}_proj.weight" for n in "qkv"]),
                   o=Lin.of(store, q + "self_attn.o_proj.weight"), norm2=rn(q + "post_attention_layernorm.weight"),
                   gate=Lin.of(store, q + "mlp.gate_proj.weight"), up=Lin.of(store, q + "mlp.up_proj.weight"),
                   down=Lin.of(store, q + "mlp.down_proj.weight"))


def ada_apply(n2d: torch.Tensor, mod: torch.Tensor, B: int) -> torch.Tensor:
    """AdaRMS modulation (pi05/transformers_pi05/gemma/modeling_gemma.py:80-87): normed * (1 + scale) + shift in fp32,
    scale / shift = the first two thirds of `mod` [B, 3w], shared by all rows of a sample.  A few hundred KB."""
    w = n2d.shape[1]
    sc, sh = mod[:, :w].float(), mod[:, w:2 * w].float()
    y = n2d.view(B, -1, w).float() * (1.0 + sc)[:, None, :] + sh[:, None, :]
    return y.to(n2d.dtype).view(-1, w)


def ada_bwd(dh2d: torch.Tensor, n2d: torch.Tensor, mod: torch.Tensor, B: int):
    """Gradient of ada_apply: (dn [rows, w], dscale [B, w] fp32, dshift [B, w] fp32)."""
    w = n2d.shape[1]
    dhf = dh2d.view(B, -1, w).float()
    dsc = (dhf * n2d.view(B, -1, w).float()).sum(dim=1)
    dsh = dhf.sum(dim=1)
    dn = (dhf * (1.0 + mod[:, :w].float())[:, None, :]).to(dh2d.dtype).view(-1, w).contiguous()
    return dn, dsc, dsh


def gated_residual(x2d, y2d, gate, B: int):
    """_gated_residual (modeling_gemma.py:101-119): x + y * gate, gate [B, w] broadcast over the rows of a sample."""
    w = x2d.shape[1]
    return (x2d.view(B, -1, w) + y2d.view(B, -1, w) * gate[:, None, :]).view(-1, w).contiguous()


class MoTLayerFn(torch.autograd.Function):
    """One joint layer of `_inner_forward_mot` (pi0_arch.py:131-216): per-stream norm + QKV, attention over the
    concatenated sequence with shared RoPE, per-stream O / residual / norm / GeGLU MLP / residual.
    `tail[t] = False` skips stream t's post-attention half (prefix stream of the last layer: its output is unused)."""

    @staticmethod
    def forward(ctx, x_p, x_s, streams, env: MoTEnv, store: ParamStore, tail, mod1=None, mod2=None):
        """mod1 / mod2 [B, 3w]: pi0.5's AdaRMS modulation (scale | shift | gate) of the SUFFIX stream's two norms
        (pi05_arch.py:146-152,217-228); None = pi0's plain norms and residuals."""
        B, (Sp, Ss) = env.B, env.lens
        S = Sp + Ss
        W = (env.heads + 2 * env.kv_heads) * env.head_dim
        C = env.heads * env.head_dim
        xs = (x_p, x_s)
        ada = (None, (mod1, mod2) if mod1 is not None else None)
        joint = torch.empty((B, S, W), device=x_p.device, dtype=x_p.dtype)
        st1 = []
        off = 0
        for x, sw, n, ad in zip(xs, streams, (Sp, Ss), ada):
            h, s1 = norm_fwd(x, sw.norm1)
            if ad is not None:
                h = ada_apply(h, ad[0], B)
            st1.append(s1)
            qkv, _ = linear_fwd(h, sw.qkv)
            ops.copy3d_(qkv, joint, B, n, W, n * W, W, S * W, W, dst_off=off * W)
            off += n
        ops.rope_(joint, env.pos, env.cos, env.sin, env.heads + env.kv_heads, env.head_dim)
        sh = ops.AttnShape(B, S, env.heads, env.kv_heads, env.head_dim, x_p.dtype)
        attn, probs = ops.attention_fwd(joint, sh, keymask=env.keymask, bid_q=env.bid, bid_k=env.bid)
        outs, keep = [], []
        off = 0
        for x, sw, n, has_tail, ad in zip(xs, streams, (Sp, Ss), tail, ada):
            if not has_tail:
                outs.append(torch.zeros(1, device=x.device, dtype=x.dtype))
                keep.append(None)
                off += n
                continue
            a = torch.empty((B * n, C), device=x.device, dtype=x.dtype)
            ops.copy3d_(attn, a, B, n, C, S * C, C, n * C, C, src_off=off * C)
            off += n
            if ad is None:
                x1, _ = linear_fwd(a, sw.o, residual=x)
                h2, s2 = norm_fwd(x1, sw.norm2)
                y_o = y_m = None
            else:
                w_ = x.shape[1]
                y_o, _ = linear_fwd(a, sw.o)
                x1 = gated_residual(x, y_o, ad[0][:, 2 * w_:], B)
                h2, s2 = norm_fwd(x1, sw.norm2)
                h2 = ada_apply(h2, ad[1], B)
            fused = (sw.gate.b is None and sw.up.b is None
                     and ops.glu_fusable(h2, sw.gate.w, sw.up.w, sw.down.w, env.act))
            if fused:         # one GEMM against gate and up, GeGLU in its epilogue (functional.block_forward)
                g = torch.empty((h2.shape[0], sw.gate.w.shape[0]), device=h2.device, dtype=h2.dtype)
                u = torch.empty_like(g)
                hm = ops.gemm_dual(h2, sw.gate.w, sw.up.w, env.act, aux_gate=g, aux_up=u)
            else:
                g, _ = linear_fwd(h2, sw.gate)
                u, _ = linear_fwd(h2, sw.up)
                hm = ops.glu_fwd(g, u, env.act)
            if ad is None:
                y, _ = linear_fwd(hm, sw.down, residual=x1)
            else:
                y_m, _ = linear_fwd(hm, sw.down)
                y = gated_residual(x1, y_m, ad[1][:, 2 * w_:], B)
            outs.append(y)
            keep.append(dict(a=a, x1=x1, s2=s2, g=g, u=u, y_o=y_o, y_m=y_m, hm=hm if fused else None))
        if mod1 is not None:
            ctx.save_for_backward(x_p, x_s, mod1, mod2)
        else:
            ctx.save_for_backward(x_p, x_s)
        ctx.misc = (streams, env, store, tail, joint, probs, st1, keep, sh, attn)
        return outs[0], outs[1]

    @staticmethod
    def backward(ctx, dy_p, dy_s):
        x_p, x_s, *mods = ctx.saved_tensors
        streams, env, store, tail, joint, probs, st1, keep, sh, attn = ctx.misc
        ctx.misc = None
        B, (Sp, Ss) = env.B, env.lens
        S = Sp + Ss
        W = (env.heads + 2 * env.kv_heads) * env.head_dim
        C = env.heads * env.head_dim
        ada = (None, tuple(mods) if mods else None)
        dmods = [None, None]          # [d(scale|shift|gate) of norm1, of norm2] for the suffix stream
        dattn = torch.zeros((B, S, C), device=x_p.device, dtype=x_p.dtype)
        dx1s = []
        off = 0
        for x, dy, sw, n, kp, ad in zip((x_p, x_s), (dy_p, dy_s), streams, (Sp, Ss), keep, ada):
            if kp is None:                      # no post-attention half: nothing flows into this stream's attn rows
                dx1s.append(None)
                off += n
                continue
            dy = dy.contiguous() if dy._base is None and dy.is_contiguous() else dy.clone(memory_format=torch.contiguous_format)
            h2, _ = norm_fwd(kp["x1"], sw.norm2)
            w_ = x.shape[1]
            if ad is not None:                  # y = x1 + y_m * g2
                n2 = h2
                h2 = ada_apply(n2, ad[1], B)
                dgate2 = (dy.view(B, -1, w_).float() * kp["y_m"].view(B, -1, w_).float()).sum(dim=1)
                dym = (dy.view(B, -1, w_) * ad[1][:, None, 2 * w_:]).view(-1, w_).contiguous()
            else:
                dym = dy
            if kp.get("hm") is not None:      # down-projection dgrad with the GLU backward in its epilogue
                linear_wgrad(store, dym, kp["hm"], sw.down)
                dg, du = ops.gemm_glu_bwd(dym, sw.down.w, kp["g"], kp["u"], env.act, dg=kp["g"], du=kp["u"])
            else:
                dhm = linear_dgrad(dym, sw.down)
                dg, du = ops.glu_bwd(dhm, kp["g"], kp["u"], env.act, dg=kp["g"], du=kp["u"], h_out=dhm)
                linear_wgrad(store, dym, dhm, sw.down)
            linear_wgrad(store, dg, h2, sw.gate)
            linear_wgrad(store, du, h2, sw.up)
            dh2 = linear_dgrad(dg, sw.gate)
            linear_dgrad(du, sw.up, out=dh2, residual=dh2)
            if ad is not None:
                dh2, dsc2, dsh2 = ada_bwd(dh2, n2, ad[1], B)
                dmods[1] = torch.cat([dsc2, dsh2, dgate2], dim=-1).to(ad[1].dtype)
            dx1 = norm_bwd(store, dh2, kp["x1"], sw.norm2, kp["s2"], dx=dy, accumulate_dx=True)
            if ad is not None:                  # x1 = x + y_o * g1
                dgate1 = (dx1.view(B, -1, w_).float() * kp["y_o"].view(B, -1, w_).float()).sum(dim=1)
                dyo = (dx1.view(B, -1, w_) * ad[0][:, None, 2 * w_:]).view(-1, w_).contiguous()
            else:
                dyo = dx1
            linear_wgrad(store, dyo, kp["a"], sw.o)
            da = linear_dgrad(dyo, sw.o)
            ops.copy3d_(da, dattn, B, n, C, n * C, C, S * C, C, dst_off=off * C)
            dx1s.append((dx1, dgate1) if ad is not None else dx1)
            off += n
        dqkv = ops.attention_bwd(dattn, joint, probs, sh, out=attn, keymask=env.keymask, bid_q=env.bid, bid_k=env.bid)
        ops.rope_(dqkv, env.pos, env.cos, env.sin, env.heads + env.kv_heads, env.head_dim, inverse=True)
        dxs = []
        off = 0
        for x, sw, n, s1, dx1, ad in zip((x_p, x_s), streams, (Sp, Ss), st1, dx1s, ada):
            dq = torch.empty((B * n, W), device=x.device, dtype=x.dtype)
            ops.copy3d_(dqkv, dq, B, n, W, S * W, W, n * W, W, src_off=off * W)
            off += n
            h, _ = norm_fwd(x, sw.norm1)
            if ad is not None:
                n1 = h
                h = ada_apply(n1, ad[0], B)
            linear_wgrad(store, dq, h, sw.qkv)
            dh = linear_dgrad(dq, sw.qkv)
            if ad is not None:
                dx1, dgate1 = dx1
                dh, dsc1, dsh1 = ada_bwd(dh, n1, ad[0], B)
                dmods[0] = torch.cat([dsc1, dsh1, dgate1], dim=-1).to(ad[0].dtype)
            if dx1 is None:
                dxs.append(norm_bwd(store, dh, x, sw.norm1, s1))
            else:
                dxs.append(norm_bwd(store, dh, x, sw.norm1, s1, dx=dx1, accumulate_dx=True))
        return dxs[0], dxs[1], None, None, None, None, dmods[0], dmods[1]


class PrefixEmbedFn(torch.autograd.Function):
    """embed_prefix (pi0_arch.py:235-269): [camera tokens ... | embed_tokens(ids) * sqrt(hidden)] per sample."""

    @staticmethod
    def forward(ctx, feats, ids32, table, g_table, store, B, n_cam, P, L):
        D = table.shape[1]
        Sp = n_cam * P + L
        out = torch.empty((B * Sp, D), device=feats.device, dtype=feats.dtype)
        for c in range(n_cam):       # feats rows are camera-major: (c*B + b)*P + p
            ops.copy3d_(feats, out, B, P, D, P * D, D, Sp * D, D, src_off=c * B * P * D, dst_off=c * P * D)
        text = ops.splice_gather(ids32, table, None)
        ops.copy3d_(text, out, B, L, D, L * D, D, Sp * D, D, alpha=math.sqrt(D), dst_off=n_cam * P * D)
        ctx.save_for_backward(ids32)
        ctx.misc = (g_table, store, B, n_cam, P, L, D, feats.shape)
        return out

    @staticmethod
    def backward(ctx, dout):
        (ids32,) = ctx.saved_tensors
        g_table, store, B, n_cam, P, L, D, fshape = ctx.misc
        Sp = n_cam * P + L
        dout = dout.contiguous()
        d_feats = torch.empty(fshape, device=dout.device, dtype=dout.dtype)
        for c in range(n_cam):
            ops.copy3d_(dout, d_feats, B, P, D, Sp * D, D, P * D, D, src_off=c * P * D, dst_off=c * B * P * D)
        if g_table is not None:
            d_text = torch.empty((B * L, D), device=dout.device, dtype=dout.dtype)
            ops.copy3d_(dout, d_text, B, L, D, Sp * D, D, L * D, D, alpha=math.sqrt(D), src_off=n_cam * P * D)
            store.begin_sparse_write(g_table)
            ops.splice_scatter(ids32, d_text, g_table, None)
        return d_feats, None, None, None, None, None, None, None, None


def posemb_sincos(time: torch.Tensor, dim: int, min_period: float, max_period: float) -> torch.Tensor:
    """pi0_arch.py:36-50 (float64 periods)."""
    fraction = torch.linspace(0.0, 1.0, dim // 2, dtype=torch.float64, device=time.device)
    period = min_period * (max_period / min_period) ** fraction
    s = time[:, None].float() / period[None, :] * 2 * np.pi
    return torch.cat([torch.sin(s), torch.cos(s)], dim=-1)


class Pi0ForCausalLM(B200Module):
    """pi0_arch.py:109-400 (training forward)."""
    config_class = Pi0Config

    def __init__(self, config: Pi0Config, device="cuda"):
        super().__init__()
        self.config = config
        llm, exp, vis = config.llm_config, config.action_config, config.vision_config
        d, w = cfg_get(llm, "hidden_size"), cfg_get(exp, "hidden_size")
        A = config.action_dim
        lin = lambda n, o, i: [ParamSpec(f"model.{n}.weight", (o, i), "action_head"),  # noqa: E731
                               ParamSpec(f"model.{n}.bias", (o,), "action_head")]
        specs = (gemma_specs(llm, "model.llm.", "llm", not config.freeze_llm, not config.freeze_llm, True)
                 + siglip_specs(vis, trainable=not config.freeze_mm_vision)
                 + [ParamSpec("model.mm_projector.weight", (d, cfg_get(vis, "hidden_size")), "projector",
                              trainable=not config.freeze_mm_projector),
                    ParamSpec("model.mm_projector.bias", (d,), "projector", trainable=not config.freeze_mm_projector)]
                 # the expert's own embedding table is never used (pi0 feeds it the suffix embeddings)
                 + gemma_specs(exp, "model.action_expert.", "action_head", True, False, False)
                 + lin("state_proj", w, A) + lin("action_in_proj", w, A) + lin("action_time_mlp_in", w, 2 * w)
                 + lin("action_time_mlp_out", w, w) + lin("action_out_proj", A, w))
        store = self._materialize(specs, device)
        self.anchor = _Anchor(store.device)
        self.tower = SiglipVisionTower(store, vis)
        self.proj = Lin.of(store, "model.mm_projector.weight", "model.mm_projector.bias")
        self.embed_w = store.w("model.llm.embed_tokens.weight")
        self.embed_g = store.g("model.llm.embed_tokens.weight")
        if self.embed_g is not None:
            store.mark_sparse_grad("model.llm.embed_tokens.weight")
        L = cfg_get(llm, "num_hidden_layers")
        assert L == cfg_get(exp, "num_hidden_layers")
        self.layers = [(_stream_w(store, llm, "model.llm.", i), _stream_w(store, exp, "model.action_expert.", i))
                       for i in range(L)]
        # optimizer / forward overlap (ParamStore.async_optimizer): chunk 2i = LLM layer i, chunk 2i+1 = expert layer i
        store.set_param_chunks([store.grad_range([n for n in store.order if n.startswith(f"model.{m}.layers.{i}.")])
                                for i in range(L) for m in ("llm", "action_expert")])
        self.expert_norm = Norm("rms1p", cfg_get(exp, "rms_norm_eps", 1e-6), store.w("model.action_expert.norm.weight"),
                                None, store.g("model.action_expert.norm.weight"))
        mk = lambda n: Lin.of(store, f"model.{n}.weight", f"model.{n}.bias")  # noqa: E731
        self.state_proj, self.action_in, self.mlp_in = mk("state_proj"), mk("action_in_proj"), mk("action_time_mlp_in")
        self.mlp_out, self.action_out = mk("action_time_mlp_out"), mk("action_out_proj")
        self.H, self.KVH = cfg_get(llm, "num_attention_heads"), cfg_get(llm, "num_key_value_heads")
        self.hd = cfg_get(llm, "head_dim") or d // self.H
        self.act = cfg_get(llm, "hidden_act") or cfg_get(llm, "hidden_activation") or "gelu_pytorch_tanh"
        self.theta = rope_theta_of(llm)
        self._rope = None
        self.d, self.w = d, w
        self.model_engine = MoTEngine(self)

    def _after_weights_changed(self) -> None:
        self.tower.refresh()

    def _rope_tables(self, n_pos: int, device):
        if self._rope is None or self._rope[0].shape[0] < n_pos:
            n = max(n_pos, 1024)
            inv = 1.0 / (self.theta ** (torch.arange(0, self.hd, 2, dtype=torch.float32, device=device) / self.hd))
            f = torch.arange(n, dtype=torch.float32, device=device)[:, None] * inv[None, :]
            self._rope = (f.cos().contiguous(), f.sin().contiguous())
        return self._rope

    def _embed_prefix(self, input_ids, attention_mask, images, image_masks):
        """embed_prefix (pi0_arch.py:235-269): all cameras through SigLIP in one batch (camera-major), linear
        projector, text embeddings * sqrt(hidden).  Returns (prefix [B*Sp, d], prefix_mask [B, Sp] bool, Sp)."""
        st = self.store
        B, n_cam, L = images.shape[0], images.shape[1], input_ids.shape[1]
        imgs = images.transpose(0, 1).reshape(n_cam * B, *images.shape[2:])
        if self.tower.g_patch is None:
            with torch.no_grad():
                f = self.tower.forward(self.anchor, imgs)
        else:
            f = self.tower.forward(self.anchor, imgs)
        f = LinearFn.apply(f, self.proj, None, st, self.tower.g_patch is not None, self.anchor.t)
        P = self.tower.P
        prefix = PrefixEmbedFn.apply(f, input_ids.to(torch.int32).contiguous(), self.embed_w, self.embed_g, st, B, n_cam,
                                     P, L)
        prefix_mask = torch.cat([image_masks.bool()[:, :, None].expand(B, n_cam, P).reshape(B, n_cam * P),
                                 attention_mask.bool()], dim=1)
        return prefix, prefix_mask, n_cam * P + L

    def _embed_suffix(self, states, x_t, time):
        """embed_suffix (pi0_arch.py:271-315): [state token | action_time_mlp(action_in(x_t) ++ posemb(time))]."""
        st, bf = self.store, torch.bfloat16
        B, T, A = x_t.shape
        state_tok = LinearFn.apply(states.to(bf).contiguous(), self.state_proj, None, st, False, self.anchor.t)
        a_tok = LinearFn.apply(x_t.to(bf).reshape(B * T, A).contiguous(), self.action_in, None, st, False, self.anchor.t)
        temb = posemb_sincos(time, self.w, 4e-3, 4.0).to(bf)[:, None, :].expand(B, T, self.w)
        at = torch.cat([a_tok.view(B, T, self.w), temb], dim=-1).reshape(B * T, 2 * self.w).contiguous()
        at = LinearFn.apply(at, self.mlp_in, "silu", st, True, None)
        at = LinearFn.apply(at, self.mlp_out, None, st, True, None)
        suffix = torch.cat([state_tok.view(B, 1, self.w), at.view(B, T, self.w)], dim=1)
        return suffix.reshape(B * (T + 1), self.w).contiguous()

    @staticmethod
    def _stream_tail(x, attn2d, sw: StreamW, act):
        """o_proj + residual, post-attention norm, GeGLU MLP + residual (pi0_arch.py:205-212), inference form."""
        x1, _ = linear_fwd(attn2d, sw.o, residual=x)
        h2, _ = norm_fwd(x1, sw.norm2)
        g, _ = linear_fwd(h2, sw.gate)
        u, _ = linear_fwd(h2, sw.up)
        y, _ = linear_fwd(ops.glu_fwd(g, u, act), sw.down, residual=x1)
        return y

    @torch.no_grad()
    def inference_action(self,
                         input_ids: torch.LongTensor = None,
                         attention_mask: Optional[torch.Tensor] = None,
                         states: Optional[torch.FloatTensor] = None,
                         images: Optional[torch.FloatTensor] = None,
                         image_masks: Optional[torch.BoolTensor] = None,
                         diffusion_steps: int = 10,
                         noise: Optional[torch.Tensor] = None,     # parity hook: the reference draws it inside
                         **kwargs) -> torch.Tensor:
        """pi0_arch.py:402-491.  One prefix pass through the Gemma stream that leaves the RoPE'd K and V of every
        layer in a [B, Sp+Ss, 2*KVH*hd] cache, then `diffusion_steps` Euler steps in which only the Ss = chunk+1
        suffix rows run through the action expert and attend over [cached prefix | own] keys."""
        if not states.is_cuda:
            raise RuntimeError("dexbotic_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        cfg, dev = self.config, states.device
        self.store.wait_all_params()
        B, T, A = states.shape[0], cfg.chunk_size, cfg.action_dim
        H, KVH, hd = self.H, self.KVH, self.hd
        W, C, Wkv = (H + 2 * KVH) * hd, H * hd, 2 * KVH * hd
        if noise is None:
            noise = torch.randn(B, T, A, device=dev)                         # :417-422
        x_t = noise.float()
        prefix, prefix_mask, Sp = self._embed_prefix(input_ids, attention_mask, images, image_masks)
        Ss = T + 1
        S = Sp + Ss
        pm = prefix_mask.to(torch.int32)
        pos_p = (torch.cumsum(pm, dim=1) - 1).clamp_(min=0).to(torch.int32).reshape(-1).contiguous()
        pos_s = (pm.sum(dim=1, keepdim=True) + torch.arange(Ss, device=dev, dtype=torch.int32)[None, :])
        pos_s = pos_s.to(torch.int32).reshape(-1).contiguous()                # :465-469
        cos, sin = self._rope_tables(S + 1, dev)
        keymask_p = prefix_mask.to(torch.uint8).contiguous()
        keymask = torch.cat([keymask_p, torch.ones(B, Ss, dtype=torch.uint8, device=dev)], dim=1).contiguous()
        blk = torch.cat([torch.zeros(Sp, dtype=torch.int32, device=dev),     # make_attn_mask (:22-29): cumsum(ar)
                         torch.ones(1, dtype=torch.int32, device=dev),
                         torch.full((T,), 2, dtype=torch.int32, device=dev)])
        bid_k = blk[None, :].expand(B, S).contiguous()
        bid_q = blk[None, Sp:].expand(B, Ss).contiguous()

        # ---- prefix pass (:432-444): fill the K/V cache; the last layer's post-attention half is never read
        caches = []
        x = prefix
        shp = ops.AttnShape(B, Sp, H, KVH, hd, x.dtype)
        for i, (sw_p, _) in enumerate(self.layers):
            h, _ = norm_fwd(x, sw_p.norm1)
            qkv, _ = linear_fwd(h, sw_p.qkv)
            ops.rope_(qkv.view(B, Sp, W), pos_p, cos, sin, H + KVH, hd)
            kv = torch.empty((B, S, Wkv), device=dev, dtype=x.dtype)
            ops.copy3d_(qkv, kv, B, Sp, Wkv, Sp * W, W, S * Wkv, Wkv, src_off=C)
            caches.append(kv)
            if i == len(self.layers) - 1:
                break
            attn, _ = ops.attention_fwd(qkv.view(B, Sp, W), shp, keymask=keymask_p)
            x = self._stream_tail(x, attn.view(B * Sp, C), sw_p, self.act)

        # ---- Euler steps (:446-489)
        dt = np.float32(-1.0 / diffusion_steps)
        t = np.float32(1.0)
        while t > -dt / 2:
            time = torch.full((B,), float(t), dtype=torch.float32, device=dev)
            xs = self._embed_suffix(states, x_t, time)
            for i, (_, sw_s) in enumerate(self.layers):
                h, _ = norm_fwd(xs, sw_s.norm1)
                qkv, _ = linear_fwd(h, sw_s.qkv)
                ops.rope_(qkv.view(B, Ss, W), pos_s, cos, sin, H + KVH, hd)
                ops.copy3d_(qkv, caches[i], B, Ss, Wkv, Ss * W, W, S * Wkv, Wkv, src_off=C, dst_off=Sp * Wkv)
                attn = ops.attention_cross(qkv.view(B, Ss, W), caches[i], B, Ss, S, H, KVH, hd, keymask=keymask,
                                           bid_q=bid_q, bid_k=bid_k)
                xs = self._stream_tail(xs, attn.view(B * Ss, C), sw_s, self.act)
            out, _ = norm_fwd(xs, self.expert_norm)
            tail = out.view(B, Ss, self.w)[:, -T:].reshape(B * T, self.w).contiguous()
            v_t, _ = linear_fwd(tail, self.action_out)
            x_t = x_t + v_t.float().view(B, T, A) * float(dt)
            t = np.float32(t + dt)
        return x_t

    def forward(self,
                input_ids: torch.LongTensor = None,
                attention_mask: Optional[torch.Tensor] = None,
                position_ids: Optional[torch.LongTensor] = None,
                past_key_values=None,
                inputs_embeds: Optional[torch.FloatTensor] = None,
                labels: Optional[torch.LongTensor] = None,
                use_cache: Optional[bool] = None,
                output_attentions: Optional[bool] = None,
                output_hidden_states: Optional[bool] = None,
                return_dict: Optional[bool] = None,
                actions: Optional[torch.FloatTensor] = None,
                states: Optional[torch.FloatTensor] = None,
                images: Optional[torch.FloatTensor] = None,
                cache_position: Optional[torch.LongTensor] = None,
                repeated_diffusion_steps: int = 4,
                image_masks: Optional[torch.BoolTensor] = None,
                noise: Optional[torch.Tensor] = None,     # parity hooks: inject the reference's random draws
                time: Optional[torch.Tensor] = None,
                **kwargs) -> CausalLMOutputDexbotic:
        if not actions.is_cuda:
            raise RuntimeError("dexbotic_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        cfg, st, dev = self.config, self.store, actions.device
        B, T, A = actions.shape[0], cfg.chunk_size, cfg.action_dim
        actions = actions.float()
        if noise is None:                                                     # pi0_arch.py:337-354
            noise = torch.randn_like(actions)
        if time is None:
            time = torch.distributions.Beta(1.5, 1.0).sample((B,)).to(dev) * 0.999 + 0.001
        te = time[:, None, None].float()
        x_t = te * noise + (1 - te) * actions
        u_t = noise - actions

        prefix, prefix_mask, Sp = self._embed_prefix(input_ids, attention_mask, images, image_masks)
        suffix = self._embed_suffix(states, x_t, time)
        Ss = T + 1

        # ---- joint attention environment (:365-370)
        S = Sp + Ss
        input_mask = torch.cat([prefix_mask, torch.ones(B, Ss, dtype=torch.bool, device=dev)], dim=1)
        ar = torch.zeros(S, dtype=torch.int32, device=dev)
        ar[Sp] = 1
        ar[Sp + 1] = 1
        bid = torch.cumsum(ar, 0).to(torch.int32)[None, :].expand(B, S).contiguous()
        pos = (torch.cumsum(input_mask.to(torch.int32), dim=1) - 1).clamp_(min=0).to(torch.int32).reshape(-1).contiguous()
        cos, sin = self._rope_tables(S + 1, dev)
        env = MoTEnv(B=B, lens=(Sp, Ss), heads=self.H, kv_heads=self.KVH, head_dim=self.hd, act=self.act,
                     keymask=input_mask.to(torch.uint8).contiguous(), bid=bid, pos=pos, cos=cos, sin=sin)
        xp, xs = prefix, suffix
        for i, streams in enumerate(self.layers):
            last = i == len(self.layers) - 1
            st.wait_chunk(2 * i)
            st.wait_chunk(2 * i + 1)
            xp, xs = MoTLayerFn.apply(xp, xs, streams, env, st, (not last, True), None, None)
        suffix_out = NormFn.apply(xs, self.expert_norm, st)
        tail = suffix_out.view(B, Ss, self.w)[:, -T:].reshape(B * T, self.w).contiguous()
        v_t = LinearFn.apply(tail, self.action_out, None, st, True, None)                   # :386
        v32 = CastFn.apply(v_t, torch.float32).view(B, T, A)
        loss = MSELossFn.apply(v32, u_t)                                                     # :387-388
        return CausalLMOutputDexbotic(loss=loss, logits=v_t.view(B, T, A))

    def zero_grad(self, set_to_none: bool = False):
        self.store.zero_grad()

    def optimizer_step(self, base_lr: float = 2e-5, mm_projector_lr=None, mm_vision_lr=None, action_head_lr=None,
                       betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0, max_grad_norm=1.0):
        lrs = {"llm": base_lr, "projector": mm_projector_lr or base_lr, "vision": mm_vision_lr or base_lr,
               "action_head": action_head_lr or base_lr}
        norm = self.store.adamw_step(lrs, betas, eps, weight_decay, max_grad_norm)
        self.tower.refresh()
        return norm
