"""Mirror of dexbotic/model/cogact/action_model/{dit.py, action_models.py, diffusion.py}: the DiT
epsilon-predictor and its training loss, computed in fp32 storage with TF32 tensor-core GEMMs — what the
reference does under `torch.amp.autocast('cuda', dtype=torch.float32)` (cogact_arch.py:133) with the
trainer's tf32=True default (base_exp.py:254).
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch

from .. import ops
from ..functional import (AttnEnv, BlockCfg, BlockW, CrossAttnFn, Lin, LinearFn, MSELossFn, Norm, NormFn,
                          TransformerBlockFn)
from ..params import ParamSpec, ParamStore

DIT_SIZES = {"DiT-S": (6, 384, 4), "DiT-B": (12, 768, 12), "DiT-L": (24, 1024, 16)}   # action_models.py:48-58


def cosine_schedule(num_steps: int = 100, max_beta: float = 0.999):
    """squaredcos_cap_v2 betas (diffusion.py:205-231) -> sqrt(alphas_cumprod), sqrt(1-alphas_cumprod), float64."""
    def alpha_bar(t):
        return math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
    betas = np.array([min(1 - alpha_bar((i + 1) / num_steps) / alpha_bar(i / num_steps), max_beta)
                      for i in range(num_steps)], dtype=np.float64)
    ac = np.cumprod(1.0 - betas, axis=0)
    return np.sqrt(ac), np.sqrt(1.0 - ac)


def action_head_specs(model_type: str, token_size: int, action_dim: int, chunk_size: int, trainable: bool = True,
                      prefix: str = "model.action_head.net.", per_token_size: Optional[int] = None) -> list[ParamSpec]:
    """per_token_size None: CogACT's DiT (cogact/action_model/dit.py); an int: MemVLA's DiT with the perceptual
    cross-attention (memvla/action_model/dit.py:158-175,241-249) and without the unused history embedder."""
    depth, w, heads = DIT_SIZES[model_type]
    g, c = "action_head", "fp32"
    T = chunk_size - 1 + 2                                         # future_action_window_size + 2 (dit.py:228-236)
    P = lambda n, s, **k: ParamSpec(prefix + n, s, g, c, trainable=trainable, **k)  # noqa: E731
    sp = [P("positional_embedding", (T, w))]
    if per_token_size is None:
        # history_embedder is built but never called (dit.py:205-207 "Action history is not used now")
        sp += [ParamSpec(prefix + "history_embedder.linear.weight", (w, action_dim), g, c, trainable=False),
               ParamSpec(prefix + "history_embedder.linear.bias", (w,), g, c, trainable=False)]
    else:
        sp += [P("per_token_embedder.linear.weight", (w, per_token_size)), P("per_token_embedder.linear.bias", (w,))]
    sp += [P("x_embedder.linear.weight", (w, action_dim)), P("x_embedder.linear.bias", (w,)),
          P("t_embedder.mlp.0.weight", (w, 256)), P("t_embedder.mlp.0.bias", (w,)),
          P("t_embedder.mlp.2.weight", (w, w)), P("t_embedder.mlp.2.bias", (w,)),
          P("z_embedder.uncondition", (1, token_size)),
          P("z_embedder.linear.weight", (w, token_size)), P("z_embedder.linear.bias", (w,))]
    for i in range(depth):
        q = f"blocks.{i}."
        sp += [P(q + "attn.qkv.weight", (3 * w, w)), P(q + "attn.qkv.bias", (3 * w,)),
               P(q + "attn.proj.weight", (w, w)), P(q + "attn.proj.bias", (w,)),
               P(q + "mlp.fc1.weight", (4 * w, w)), P(q + "mlp.fc1.bias", (4 * w,)),
               P(q + "mlp.fc2.weight", (w, 4 * w)), P(q + "mlp.fc2.bias", (w,))]
        if per_token_size is not None:
            sp += [P(q + "per_attn.in_proj_weight", (3 * w, w)), P(q + "per_attn.in_proj_bias", (3 * w,)),
                   P(q + "per_attn.out_proj.weight", (w, w)), P(q + "per_attn.out_proj.bias", (w,)),
                   P(q + "norm3.weight", (w,), no_decay=True), P(q + "norm3.bias", (w,), no_decay=True)]
    sp += [P("final_layer.linear.weight", (action_dim, w)), P("final_layer.linear.bias", (action_dim,))]
    return sp


class _QSampleFn(torch.autograd.Function):
    """x_t = sqrt(ac[t]) x0 + sqrt(1-ac[t]) eps (diffusion.py:308-326); no gradient is needed for x0 / eps."""

    @staticmethod
    def forward(ctx, x, noise, t, sa, sb):
        return ops.q_sample(x.contiguous(), noise.contiguous(), t, sa, sb)

    @staticmethod
    def backward(ctx, g):
        return None, None, None, None, None


class ActionModel:
    """action_models.py:63-135 (ActionModel) wrapping dit.py:181-292 (DiT)."""

    def __init__(self, store: ParamStore, model_type: str, token_size: int, in_channels: int,
                 future_action_window_size: int, diffusion_steps: int = 100, prefix: str = "model.action_head.",
                 use_per_attn: bool = False):
        self.store = store
        self.model_type = model_type
        self.use_per_attn = use_per_attn
        self.depth, self.w, self.heads = DIT_SIZES[model_type]
        self.in_channels = in_channels
        self.T = future_action_window_size + 1
        self.num_timesteps = diffusion_steps
        self.class_dropout_prob = 0.1
        p = prefix + "net."
        self.pos_emb = None
        self._p = p
        L = lambda w, b: Lin.of(store, p + w, p + b)  # noqa: E731
        self.x_emb = L("x_embedder.linear.weight", "x_embedder.linear.bias")
        self.t0 = L("t_embedder.mlp.0.weight", "t_embedder.mlp.0.bias")
        self.t2 = L("t_embedder.mlp.2.weight", "t_embedder.mlp.2.bias")
        self.z_lin = L("z_embedder.linear.weight", "z_embedder.linear.bias")
        self.final = L("final_layer.linear.weight", "final_layer.linear.bias")
        bc = BlockCfg(d=self.w, heads=self.heads, kv_heads=self.heads, head_dim=self.w // self.heads, inter=4 * self.w,
                      mlp="mlp", act="gelu_tanh", rope=False)
        self.blocks = []
        for i in range(self.depth):
            q = f"{p}blocks.{i}."
            self.blocks.append(BlockW(cfg=bc, norm1=Norm("ln_noaffine", 1e-6),
                                      qkv=Lin.of(store, q + "attn.qkv.weight", q + "attn.qkv.bias"),
                                      o=Lin.of(store, q + "attn.proj.weight", q + "attn.proj.bias"),
                                      norm2=Norm("ln_noaffine", 1e-6),
                                      fc1=Lin.of(store, q + "mlp.fc1.weight", q + "mlp.fc1.bias"),
                                      fc2=Lin.of(store, q + "mlp.fc2.weight", q + "mlp.fc2.bias")))
        self.norm_final = Norm("ln_noaffine", 1e-6)
        self.per = []
        if use_per_attn:                       # memvla/action_model/dit.py:158-175: one nn.MultiheadAttention per block
            self.per_emb = L("per_token_embedder.linear.weight", "per_token_embedder.linear.bias")
            w = self.w
            for i in range(self.depth):
                q = f"{p}blocks.{i}."
                wi, bi = store.w(q + "per_attn.in_proj_weight"), store.w(q + "per_attn.in_proj_bias")
                gi, gbi = store.g(q + "per_attn.in_proj_weight"), store.g(q + "per_attn.in_proj_bias")
                sl = lambda t, a, b: None if t is None else t[a:b]  # noqa: E731
                self.per.append(dict(
                    norm3=Norm("ln", 1e-6, store.w(q + "norm3.weight"), store.w(q + "norm3.bias"),
                               store.g(q + "norm3.weight"), store.g(q + "norm3.bias")),
                    q=Lin(wi[:w], bi[:w], sl(gi, 0, w), sl(gbi, 0, w)),
                    kv=Lin(wi[w:], bi[w:], sl(gi, w, 3 * w), sl(gbi, w, 3 * w)),
                    o=Lin.of(store, q + "per_attn.out_proj.weight", q + "per_attn.out_proj.bias")))
        sa, sb = cosine_schedule(diffusion_steps)
        # _extract_into_tensor (diffusion.py:975-987): float64 table entry -> float32
        self.sqrt_ac = torch.from_numpy(sa).float().to(store.device)
        self.sqrt_1mac = torch.from_numpy(sb).float().to(store.device)

    # torch.nn.Parameter views (fp32 master == compute tensor in region B) for the autograd-visible glue
    def _param(self, name: str, module) -> torch.nn.Parameter:
        mod = module
        for part in (self._p + name).split("."):
            mod = getattr(mod, part)
        return mod

    def _block_with_per_attn(self, x2d, bw: BlockW, pa: dict, per_emb, N: int, S: int, groups: int):
        """MemVLA DiTBlock.forward (memvla/action_model/dit.py:176-187) from the small Functions:
        x += attn(norm1 x); x += per_attn(norm3 x, per, per); x += mlp(norm2 x).

        The reference repeats per_token `groups` times along the batch (memvla_arch.py:641-645); cross-attention has
        no mask, so the repeats of one sample are stacked as extra QUERY rows of that sample instead: K/V are
        projected once per sample and no repeated copy exists."""
        st, w, H = self.store, self.w, self.heads
        h = NormFn.apply(x2d, bw.norm1, st)
        qkv = LinearFn.apply(h, bw.qkv, None, st, True, None)
        a = CrossAttnFn.apply(qkv[:, :w], qkv[:, w:2 * w], qkv[:, 2 * w:], N, S, S, H)
        x2d = x2d + LinearFn.apply(a, bw.o, None, st, True, None)
        Bs = N // groups
        P = per_emb.shape[0] // Bs
        h = NormFn.apply(x2d, pa["norm3"], st)
        q = LinearFn.apply(h, pa["q"], None, st, True, None)
        q = q.view(groups, Bs, S, w).transpose(0, 1).reshape(Bs * groups * S, w)         # [(b, r, s), w]
        kv = LinearFn.apply(per_emb, pa["kv"], None, st, True, None)
        c = CrossAttnFn.apply(q, kv[:, :w], kv[:, w:], Bs, groups * S, P, H)
        c = c.view(Bs, groups, S, w).transpose(0, 1).reshape(N * S, w)
        x2d = x2d + LinearFn.apply(c.contiguous(), pa["o"], None, st, True, None)
        h = NormFn.apply(x2d, bw.norm2, st)
        m = LinearFn.apply(h, bw.fc1, "gelu_tanh", st, True, None)
        return x2d + LinearFn.apply(m, bw.fc2, None, st, True, None)

    def net(self, module, x_t, t, z, drop_mask, per_token=None, groups: int = 1):
        """DiT.forward (dit.py:273-292) on fp32 tensors: x_t [N,T,A], t [N] int32, z [N,1,D] -> eps_hat [N,T,A].
        MemVLA: per_token [N/groups, P, D_per] fp32 (one copy per sample; `groups` = how often the batch repeats it)."""
        st = self.store
        N, T, A = x_t.shape
        w = self.w
        anchor = module.model_engine.anchor.t
        x = LinearFn.apply(x_t.reshape(N * T, A), self.x_emb, None, st, False, anchor)            # x_embedder
        te = ops.timestep_embedding(t.float(), 256, torch.float32)                                 # dit.py:37-56
        te = LinearFn.apply(LinearFn.apply(te, self.t0, "silu", st, False, anchor), self.t2, None, st, True, None)
        if drop_mask is not None:                                                                  # token_drop, :80-95
            unc = self._param("z_embedder.uncondition", module)
            z = torch.where(drop_mask[:, None, None], unc[None].to(z.dtype), z)
        ze = LinearFn.apply(z.reshape(N, -1).contiguous(), self.z_lin, None, st, True, None)
        c = te + ze                                                                                # :282
        pos = self._param("positional_embedding", module)
        x = torch.cat([c[:, None, :], x.view(N, T, w)], dim=1) + pos                               # :283-284
        x2d = x.reshape(N * (T + 1), w).contiguous()
        env = AttnEnv(B=N, S=T + 1)
        if self.use_per_attn:
            assert per_token is not None and N % groups == 0
            Bs, P, Dp = per_token.shape
            per_emb = LinearFn.apply(per_token.reshape(Bs * P, Dp).contiguous(), self.per_emb, None, st, True, None)
            for bw, pa in zip(self.blocks, self.per):
                x2d = self._block_with_per_attn(x2d, bw, pa, per_emb, N, T + 1, groups)
        else:
            for bw in self.blocks:
                x2d = TransformerBlockFn.apply(x2d, bw, env, st, False)  # ~30 MB of activations: never recompute
        x2d = NormFn.apply(x2d, self.norm_final, st)
        out = LinearFn.apply(x2d, self.final, None, st, True, None)                                      # FinalLayer
        return out.view(N, T + 1, A)[:, 1:, :]

    def loss(self, module, x, z, noise: Optional[torch.Tensor] = None, timestep: Optional[torch.Tensor] = None,
             drop_mask: Optional[torch.Tensor] = None, training: bool = True, per_token=None, groups: int = 1,
             sample_weight: Optional[torch.Tensor] = None):
        """ActionModel.loss (action_models.py:102-125).  noise / timestep / drop_mask may be injected for parity.
        sample_weight [N] (hybrid co-training, hybrid_cogact_arch.py:182-187: reduction="none", per-sample mean times
        has_action, summed and divided by sum(has_action) + 1e-6): sum_n w_n mean_{t,a}(d_n^2) / (sum w + 1e-6)."""
        N = x.shape[0]
        if noise is None:
            noise = torch.randn_like(x)
        if timestep is None:
            timestep = torch.randint(0, self.num_timesteps, (N,), device=x.device)
        if drop_mask is None and training and self.class_dropout_prob > 0:
            drop_mask = torch.rand(N, device=x.device) < self.class_dropout_prob
        t32 = timestep.to(torch.int32)
        x_t = _QSampleFn.apply(x, noise, t32, self.sqrt_ac, self.sqrt_1mac)
        pred = self.net(module, x_t, t32, z, drop_mask, per_token, groups)
        assert pred.shape == noise.shape == x.shape
        if sample_weight is None:
            return MSELossFn.apply(pred, noise)
        # mean over all N*T*A of w_n d^2 equals (1/N) sum_n w_n mean_ta(d_n^2): scale both sides by sqrt(w_n) (a few KB)
        w = sample_weight.reshape(N).to(torch.float32)
        r = w.sqrt()[:, None, None]
        return MSELossFn.apply(pred * r, noise * r) * (N / (w.sum() + 1e-6))


    # ------------------------------------------------------------------ inference (cogact_arch.py:149-198)
    def ddim_tables(self, ddim_steps: int):
        """space_timesteps('ddimN') + SpacedDiffusion.__init__ (diffusion.py:990-1080): respaced timestep map and
        float64 alphas_cumprod / alphas_cumprod_prev."""
        sa, _ = cosine_schedule(self.num_timesteps)
        ac = sa ** 2
        stride = next(i for i in range(1, self.num_timesteps) if len(range(0, self.num_timesteps, i)) == ddim_steps)
        tmap = list(range(0, self.num_timesteps, stride))
        last, betas = 1.0, []
        for i in tmap:
            betas.append(1 - ac[i] / last)
            last = ac[i]
        ac2 = np.cumprod(1.0 - np.array(betas, dtype=np.float64))
        return tmap, ac2, np.append(1.0, ac2[:-1])

    @torch.no_grad()
    def sample(self, module, cognition, noise, cfg_scale: float = 1.5, num_ddim_steps: int = 10, per_token=None):
        """ddim_sample_loop (diffusion.py:714-795) with eta=0, clip_denoised=False, classifier-free guidance through
        forward_with_cfg (dit.py:294-311).  cognition [B,1,D] fp32, noise [B,T,A] fp32 -> samples [B,T,A]."""
        B = cognition.shape[0]
        use_cfg = cfg_scale > 1.0
        x = torch.cat([noise, noise], 0) if use_cfg else noise
        if use_cfg:
            unc = self._param("z_embedder.uncondition", module)[None].expand(B, 1, -1).to(cognition.dtype)
            z = torch.cat([cognition, unc], 0)
        else:
            z = cognition
        tmap, ac, ac_prev = self.ddim_tables(num_ddim_steps)
        f32 = lambda v: float(np.float32(v))  # noqa: E731   (_extract_into_tensor: float64 table -> float32)
        for i in reversed(range(len(tmap))):
            t = torch.full((x.shape[0],), tmap[i], device=x.device, dtype=torch.int32)
            if use_cfg:
                half = x[: x.shape[0] // 2]
                out = self.net(module, torch.cat([half, half], 0).contiguous(), t, z, None, per_token, 2)
                cond, uncond = out.chunk(2, dim=0)
                e = uncond + cfg_scale * (cond - uncond)
                eps_model = torch.cat([e, e], 0)
            else:
                eps_model = self.net(module, x.contiguous(), t, z, None, per_token, 1)
            sr, srm1 = f32(np.sqrt(1.0 / ac[i])), f32(np.sqrt(1.0 / ac[i] - 1))
            pred_x0 = sr * x - srm1 * eps_model
            eps = (sr * x - pred_x0) / srm1
            x = pred_x0 * f32(np.sqrt(np.float32(ac_prev[i]))) + f32(np.sqrt(1 - np.float32(ac_prev[i]))) * eps
        return x[:B] if use_cfg else x
