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
from ..functional import AttnEnv, BlockCfg, BlockW, Lin, LinearFn, MSELossFn, Norm, NormFn, TransformerBlockFn
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
                      prefix: str = "model.action_head.net.") -> list[ParamSpec]:
    depth, w, heads = DIT_SIZES[model_type]
    g, c = "action_head", "fp32"
    T = chunk_size - 1 + 2                                         # future_action_window_size + 2 (dit.py:228-236)
    P = lambda n, s, **k: ParamSpec(prefix + n, s, g, c, trainable=trainable, **k)  # noqa: E731
    sp = [P("positional_embedding", (T, w), no_decay=True),
          # history_embedder is built but never called (dit.py:205-207 "Action history is not used now")
          ParamSpec(prefix + "history_embedder.linear.weight", (w, action_dim), g, c, trainable=False),
          ParamSpec(prefix + "history_embedder.linear.bias", (w,), g, c, trainable=False),
          P("x_embedder.linear.weight", (w, action_dim)), P("x_embedder.linear.bias", (w,)),
          P("t_embedder.mlp.0.weight", (w, 256)), P("t_embedder.mlp.0.bias", (w,)),
          P("t_embedder.mlp.2.weight", (w, w)), P("t_embedder.mlp.2.bias", (w,)),
          P("z_embedder.uncondition", (1, token_size), no_decay=True),
          P("z_embedder.linear.weight", (w, token_size)), P("z_embedder.linear.bias", (w,))]
    for i in range(depth):
        q = f"blocks.{i}."
        sp += [P(q + "attn.qkv.weight", (3 * w, w)), P(q + "attn.qkv.bias", (3 * w,)),
               P(q + "attn.proj.weight", (w, w)), P(q + "attn.proj.bias", (w,)),
               P(q + "mlp.fc1.weight", (4 * w, w)), P(q + "mlp.fc1.bias", (4 * w,)),
               P(q + "mlp.fc2.weight", (w, 4 * w)), P(q + "mlp.fc2.bias", (w,))]
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
                 future_action_window_size: int, diffusion_steps: int = 100, prefix: str = "model.action_head."):
        self.store = store
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

    def net(self, module, x_t, t, z, drop_mask):
        """DiT.forward (dit.py:273-292) on fp32 tensors: x_t [N,T,A], t [N] int32, z [N,1,D] -> eps_hat [N,T,A]."""
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
        for bw in self.blocks:
            x2d = TransformerBlockFn.apply(x2d, bw, env, st, False)      # ~30 MB of activations: never recompute
        x2d = NormFn.apply(x2d, self.norm_final, st)
        out = LinearFn.apply(x2d, self.final, None, st, True, None)                                      # FinalLayer
        return out.view(N, T + 1, A)[:, 1:, :]

    def loss(self, module, x, z, noise: Optional[torch.Tensor] = None, timestep: Optional[torch.Tensor] = None,
             drop_mask: Optional[torch.Tensor] = None, training: bool = True):
        """ActionModel.loss (action_models.py:102-125).  noise / timestep / drop_mask may be injected for parity."""
        N = x.shape[0]
        if noise is None:
            noise = torch.randn_like(x)
        if timestep is None:
            timestep = torch.randint(0, self.num_timesteps, (N,), device=x.device)
        if drop_mask is None and training and self.class_dropout_prob > 0:
            drop_mask = torch.rand(N, device=x.device) < self.class_dropout_prob
        t32 = timestep.to(torch.int32)
        x_t = _QSampleFn.apply(x, noise, t32, self.sqrt_ac, self.sqrt_1mac)
        pred = self.net(module, x_t, t32, z, drop_mask)
        assert pred.shape == noise.shape == x.shape
        return MSELossFn.apply(pred, noise)
