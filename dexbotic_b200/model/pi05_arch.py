"""Mirror of dexbotic/model/pi05/pi05_arch.py: Pi05Config / Pi05ForCausalLM — pi0's
SigLIP + Gemma mixture-of-transformers, with the action expert conditioned on the flow-matching time through
ADAPTIVE RMSNorm: every expert norm computes scale / shift / gate = chunk(dense(adarms_cond), 3)
(pi05/transformers_pi05/gemma/modeling_gemma.py:38-88), the gates multiply the residual branches
(_gated_residual, :101-119), there is no state token and no action-time MLP (pi05_arch.py:296-331).

State-dict keys == the reference's (model.action_expert.layers.N.{input,post_attention}_layernorm.dense.{weight,bias},
model.action_expert.norm.dense.*, model.time_mlp_{in,out}.*, model.action_{in,out}_proj.*, the rest as pi0).
The joint layer is pi0's MoTLayerFn with the suffix stream's modulation tensors passed in.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .. import ops
from ..functional import CastFn, Lin, LinearFn, MSELossFn, Norm, linear_fwd, norm_bwd, norm_fwd
from ..params import ParamSpec
from ._module import B200Module
from .dexbotic_arch import CausalLMOutputDexbotic, _Anchor, cfg_get, rope_theta_of
from .pi0_arch import (MoTEnv, MoTLayerFn, Pi0Config, PrefixEmbedFn, SiglipVisionTower, StreamW, _stream_w, ada_apply,
                       ada_bwd, gated_residual, gemma_specs, posemb_sincos, siglip_specs)


class Pi05Config(Pi0Config):
    """pi05_arch.py:53-84."""
    model_type = "dexbotic_pi05"


def adarms_expert_specs(cfg, prefix: str = "model.action_expert.") -> list[ParamSpec]:
    """AdaRMSGemmaModel with use_adarms=True (modeling_gemma.py:195-213,279-310): the three norms of every layer
    position are `dense` Linear(cond_dim, 3*hidden) instead of a weight vector."""
    g = "action_head"
    d, V = cfg_get(cfg, "hidden_size"), cfg_get(cfg, "vocab_size")
    H, KVH = cfg_get(cfg, "num_attention_heads"), cfg_get(cfg, "num_key_value_heads")
    hd = cfg_get(cfg, "head_dim") or d // H
    inter, L = cfg_get(cfg, "intermediate_size"), cfg_get(cfg, "num_hidden_layers")
    cd = cfg_get(cfg, "adarms_cond_dim") or d
    # the expert's embedding table is never used (the suffix is action_in_proj(x_t), pi05_arch.py:317)
    sp = [ParamSpec(prefix + "embed_tokens.weight", (V, d), g, trainable=False)]
    for i in range(L):
        q = f"{prefix}layers.{i}."
        for n, rows in (("q", H * hd), ("k", KVH * hd), ("v", KVH * hd)):
            sp.append(ParamSpec(f"{q}self_attn.{n}_proj.weight", (rows, d), g, fuse=q + "qkvw"))
        sp += [ParamSpec(q + "self_attn.o_proj.weight", (d, H * hd), g),
               ParamSpec(q + "mlp.gate_proj.weight", (inter, d), g), ParamSpec(q + "mlp.up_proj.weight", (inter, d), g),
               ParamSpec(q + "mlp.down_proj.weight", (d, inter), g)]
        for n in ("input_layernorm", "post_attention_layernorm"):
            sp += [ParamSpec(f"{q}{n}.dense.weight", (3 * d, cd), g), ParamSpec(f"{q}{n}.dense.bias", (3 * d,), g)]
    sp += [ParamSpec(prefix + "norm.dense.weight", (3 * d, cd), g), ParamSpec(prefix + "norm.dense.bias", (3 * d,), g)]
    return sp


class AdaNormFn(torch.autograd.Function):
    """GemmaRMSNorm.forward with a condition, gate unused (the expert's final norm, pi05_arch.py:232-238)."""

    @staticmethod
    def forward(ctx, x2d, mod, norm: Norm, store, B):
        n, stats = norm_fwd(x2d, norm)
        ctx.save_for_backward(x2d, mod, *stats)
        ctx.misc = (norm, store, B)
        return ada_apply(n, mod, B)

    @staticmethod
    def backward(ctx, dy):
        x2d, mod, *stats = ctx.saved_tensors
        norm, store, B = ctx.misc
        n, _ = norm_fwd(x2d, norm)
        dn, dsc, dsh = ada_bwd(dy.contiguous(), n, mod, B)
        dx = norm_bwd(store, dn, x2d, norm, tuple(stats))
        dmod = torch.cat([dsc, dsh, torch.zeros_like(dsc)], dim=-1).to(mod.dtype)
        return dx, dmod, None, None, None


class Pi05ForCausalLM(B200Module):
    """pi05_arch.py:110-514."""
    config_class = Pi05Config

    def __init__(self, config: Pi05Config, device="cuda"):
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
                 + adarms_expert_specs(exp)
                 + lin("time_mlp_in", w, w) + lin("time_mlp_out", w, w) + lin("action_in_proj", w, A)
                 + lin("action_out_proj", A, w))
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
        eps = cfg_get(exp, "rms_norm_eps", 1e-6)
        # pure normalisation for the expert: (1 + 0) * x * rstd, no parameter (the modulation comes from `dense`)
        self._zero_w = torch.zeros(w, device=store.device, dtype=torch.bfloat16)
        plain = lambda: Norm("rms1p", eps, self._zero_w, None, None)  # noqa: E731
        self.layers, self.mods = [], []
        for i in range(L):
            q = f"model.action_expert.layers.{i}."
            sw = StreamW(d=w, inter=cfg_get(exp, "intermediate_size"), norm1=plain(),
                         qkv=Lin.of(store, [f"{q}self_attn.{n}_proj.weight" for n in "qkv"]),
                         o=Lin.of(store, q + "self_attn.o_proj.weight"), norm2=plain(),
                         gate=Lin.of(store, q + "mlp.gate_proj.weight"), up=Lin.of(store, q + "mlp.up_proj.weight"),
                         down=Lin.of(store, q + "mlp.down_proj.weight"))
            self.layers.append((_stream_w(store, llm, "model.llm.", i), sw))
            self.mods.append((Lin.of(store, q + "input_layernorm.dense.weight", q + "input_layernorm.dense.bias"),
                              Lin.of(store, q + "post_attention_layernorm.dense.weight",
                                     q + "post_attention_layernorm.dense.bias")))
        store.set_param_chunks([store.grad_range([n for n in store.order if n.startswith(f"model.{m}.layers.{i}.")])
                                for i in range(L) for m in ("llm", "action_expert")])
        self.final_mod = Lin.of(store, "model.action_expert.norm.dense.weight", "model.action_expert.norm.dense.bias")
        self.final_norm = plain()
        mk = lambda n: Lin.of(store, f"model.{n}.weight", f"model.{n}.bias")  # noqa: E731
        self.time_in, self.time_out = mk("time_mlp_in"), mk("time_mlp_out")
        self.action_in, self.action_out = mk("action_in_proj"), mk("action_out_proj")
        self.H, self.KVH = cfg_get(llm, "num_attention_heads"), cfg_get(llm, "num_key_value_heads")
        self.hd = cfg_get(llm, "head_dim") or d // self.H
        self.act = cfg_get(llm, "hidden_act") or cfg_get(llm, "hidden_activation") or "gelu_pytorch_tanh"
        self.theta = rope_theta_of(llm)
        self._rope = None
        self.d, self.w = d, w
        from .pi0_arch import MoTEngine
        self.model_engine = MoTEngine(self)

    def _after_weights_changed(self) -> None:
        self.tower.refresh()

    _rope_tables = None  # set below (shared with pi0)

    def _embed_prefix(self, input_ids, attention_mask, images, image_masks):
        """embed_prefix (pi05_arch.py:247-290) — identical to pi0's."""
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

    def _embed_suffix(self, x_t, time):
        """embed_suffix (pi05_arch.py:292-331): action tokens + adarms_cond = silu(time_mlp_out(silu(time_mlp_in(
        posemb(t))))).  Returns (suffix [B*T, w], cond [B, w])."""
        st, bf = self.store, torch.bfloat16
        B, T, A = x_t.shape
        temb = posemb_sincos(time, self.w, 4e-3, 4.0).to(bf).contiguous()
        c = LinearFn.apply(temb, self.time_in, "silu", st, False, self.anchor.t)
        c = LinearFn.apply(c, self.time_out, "silu", st, True, None)
        suffix = LinearFn.apply(x_t.to(bf).reshape(B * T, A).contiguous(), self.action_in, None, st, False, self.anchor.t)
        return suffix, c

    def _mods(self, cond, i):
        st = self.store
        m1, m2 = self.mods[i]
        return (LinearFn.apply(cond, m1, None, st, True, None), LinearFn.apply(cond, m2, None, st, True, None))

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
        if noise is None:                                                     # pi05_arch.py:352-366
            noise = torch.randn_like(actions)
        if time is None:
            time = torch.distributions.Beta(1.5, 1.0).sample((B,)).to(dev) * 0.999 + 0.001
        te = time[:, None, None].float()
        x_t = te * noise + (1 - te) * actions
        u_t = noise - actions
        prefix, prefix_mask, Sp = self._embed_prefix(input_ids, attention_mask, images, image_masks)
        suffix, cond = self._embed_suffix(x_t, time)
        Ss = T
        S = Sp + Ss
        input_mask = torch.cat([prefix_mask, torch.ones(B, Ss, dtype=torch.bool, device=dev)], dim=1)
        ar = torch.zeros(S, dtype=torch.int32, device=dev)
        ar[Sp] = 1                                                            # :329 [True] + [False] * (chunk - 1)
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
            m1, m2 = self._mods(cond, i)
            xp, xs = MoTLayerFn.apply(xp, xs, streams, env, st, (not last, True), m1, m2)
        fm = LinearFn.apply(cond, self.final_mod, None, st, True, None)
        suffix_out = AdaNormFn.apply(xs, fm, self.final_norm, st, B)
        v_t = LinearFn.apply(suffix_out, self.action_out, None, st, True, None)              # :405 (all T rows)
        v32 = CastFn.apply(v_t, torch.float32).view(B, T, A)
        loss = MSELossFn.apply(v32, u_t)
        return CausalLMOutputDexbotic(loss=loss, logits=v_t.view(B, T, A))

    # ------------------------------------------------------------------ inference (pi05_arch.py:424-514)
    def _suffix_layer(self, xs, sw: StreamW, m1, m2, cache, B, Ss, Sp, pos_s, cos, sin, keymask, bid_q, bid_k):
        H, KVH, hd = self.H, self.KVH, self.hd
        W, C, Wkv, w = (H + 2 * KVH) * hd, H * hd, 2 * KVH * hd, self.w
        S = Sp + Ss
        n1, _ = norm_fwd(xs, sw.norm1)
        qkv, _ = linear_fwd(ada_apply(n1, m1, B), sw.qkv)
        ops.rope_(qkv.view(B, Ss, W), pos_s, cos, sin, H + KVH, hd)
        ops.copy3d_(qkv, cache, B, Ss, Wkv, Ss * W, W, S * Wkv, Wkv, src_off=C, dst_off=Sp * Wkv)
        attn = ops.attention_cross(qkv.view(B, Ss, W), cache, B, Ss, S, H, KVH, hd, keymask=keymask, bid_q=bid_q,
                                   bid_k=bid_k)
        y_o, _ = linear_fwd(attn.view(B * Ss, C), sw.o)
        x1 = gated_residual(xs, y_o, m1[:, 2 * w:], B)
        n2, _ = norm_fwd(x1, sw.norm2)
        h2 = ada_apply(n2, m2, B)
        g, _ = linear_fwd(h2, sw.gate)
        u, _ = linear_fwd(h2, sw.up)
        y_m, _ = linear_fwd(ops.glu_fwd(g, u, self.act), sw.down)
        return gated_residual(x1, y_m, m2[:, 2 * w:], B)

    @torch.no_grad()
    def inference_action(self,
                         input_ids: torch.LongTensor = None,
                         attention_mask: Optional[torch.Tensor] = None,
                         states: Optional[torch.FloatTensor] = None,
                         images: Optional[torch.FloatTensor] = None,
                         image_masks: Optional[torch.BoolTensor] = None,
                         diffusion_steps: int = 10,
                         noise: Optional[torch.Tensor] = None,
                         **kwargs) -> torch.Tensor:
        """One prefix pass that fills the per-layer K/V cache, then `diffusion_steps` Euler steps of the suffix through
        the AdaRMS action expert (the time changes every step, so the modulations are recomputed per step)."""
        if not states.is_cuda:
            raise RuntimeError("dexbotic_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        from .pi0_arch import Pi0ForCausalLM
        self.store.wait_all_params()
        cfg, dev = self.config, states.device
        B, T, A = states.shape[0], cfg.chunk_size, cfg.action_dim
        H, KVH, hd = self.H, self.KVH, self.hd
        W, C, Wkv = (H + 2 * KVH) * hd, H * hd, 2 * KVH * hd
        if noise is None:
            noise = torch.randn(B, T, A, device=dev)
        x_t = noise.float()
        prefix, prefix_mask, Sp = self._embed_prefix(input_ids, attention_mask, images, image_masks)
        Ss = T
        S = Sp + Ss
        pm = prefix_mask.to(torch.int32)
        pos_p = (torch.cumsum(pm, dim=1) - 1).clamp_(min=0).to(torch.int32).reshape(-1).contiguous()
        pos_s = (pm.sum(dim=1, keepdim=True) + torch.arange(Ss, device=dev, dtype=torch.int32)[None, :])
        pos_s = pos_s.to(torch.int32).reshape(-1).contiguous()
        cos, sin = self._rope_tables(S + 1, dev)
        keymask_p = prefix_mask.to(torch.uint8).contiguous()
        keymask = torch.cat([keymask_p, torch.ones(B, Ss, dtype=torch.uint8, device=dev)], dim=1).contiguous()
        blk = torch.cat([torch.zeros(Sp, dtype=torch.int32, device=dev), torch.ones(T, dtype=torch.int32, device=dev)])
        bid_k = blk[None, :].expand(B, S).contiguous()
        bid_q = blk[None, Sp:].expand(B, Ss).contiguous()
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
            x = Pi0ForCausalLM._stream_tail(x, attn.view(B * Sp, C), sw_p, self.act)
        dt = np.float32(-1.0 / diffusion_steps)
        t = np.float32(1.0)
        while t > -dt / 2:
            time = torch.full((B,), float(t), dtype=torch.float32, device=dev)
            xs, cond = self._embed_suffix(x_t, time)
            for i, (_, sw_s) in enumerate(self.layers):
                m1, m2 = self._mods(cond, i)
                xs = self._suffix_layer(xs, sw_s, m1, m2, caches[i], B, Ss, Sp, pos_s, cos, sin, keymask, bid_q, bid_k)
            fm, _ = linear_fwd(cond, self.final_mod)
            n, _ = norm_fwd(xs, self.final_norm)
            v_t, _ = linear_fwd(ada_apply(n, fm, B), self.action_out)
            x_t = x_t + v_t.float().view(B, T, A) * float(dt)
            t = np.float32(t + dt)
        return x_t

    def zero_grad(self, set_to_none: bool = False):
        self.store.zero_grad()

    def optimizer_step(self, base_lr: float = 2e-5, mm_projector_lr=None, mm_vision_lr=None, action_head_lr=None,
                       betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0, max_grad_norm=1.0):
        lrs = {"llm": base_lr, "projector": mm_projector_lr or base_lr, "vision": mm_vision_lr or base_lr,
               "action_head": action_head_lr or base_lr}
        norm = self.store.adamw_step(lrs, betas, eps, weight_decay, max_grad_norm)
        self.tower.refresh()
        return norm


from .pi0_arch import Pi0ForCausalLM as _Pi0  # noqa: E402

Pi05ForCausalLM._rope_tables = _Pi0._rope_tables
