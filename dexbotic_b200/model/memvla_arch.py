"""Mirror of dexbotic/model/memvla/memvla_arch.py: MemVLAConfig / MemVLAModel /
MemVLAForCausalLM — the CogACT-shaped VLM trunk plus

  * `per_compr`  BottleneckSE (memvla_arch.py:136-173): squeeze-excite + per-token bottleneck on the projected vision
    features (the reference's 1x1 convs on a [B,C,H,W] permute are GEMMs on the channel-last [B,P,C] map as stored);
  * `per_cog_mem_bank`  PerCogMemBank (:195-427): per-episode memory of past fused tokens, CrossTransformerBlock
    retrieval (:82-133), GateFusion (:176-192), FIFO / token-merge consolidation (:247-296);
  * a DiT action head with a perceptual cross-attention per block (memvla/action_model/dit.py:158-187).

State-dict keys and shapes == the reference's (model.per_compr.{excite,reduce}.N.weight keep their conv shape
[out, in, 1, 1]).  The memory bank is host-side STATE (dict episode -> list of (timestep, detached feature)), exactly
as in the reference; its arithmetic (projections, attention, norms, FFN, gate, dropout) runs in the CUDA kernels.
Samples are processed in batch order because sample i reads what samples < i of the same episode wrote (:351-407).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from .. import ops
from ..functional import (CastFn, CrossAttnFn, DropoutFn, GateFuseFn, GatherRowsFn, Lin, LinearFn, Norm, NormFn,
                          SEGateFn)
from ..params import ParamSpec, ParamStore
from ._module import B200Module
from .action_model import ActionModel, action_head_specs
from .dexbotic_arch import (CausalLMOutputDexbotic, DexboticConfig, DexboticVLMModel, cfg_get, clip_specs, llm_specs,
                            projector_specs)

ROLES = ("per", "cog")
MEM_KEYS = ("dataloader_type", "group_size", "mem_length", "retrieval_layers", "use_timestep_pe", "fusion_type",
            "consolidate_type", "per_token_size")


class MemVLAConfig(DexboticConfig):
    """memvla_arch.py:20-24 + the memory keys memvla_exp.py:198-236 writes into the config."""
    model_type = "dexbotic_memvla"

    def __init__(self, action_model_type: Optional[str] = None, action_dim: Optional[int] = None,
                 chunk_size: Optional[int] = None, dataloader_type: str = "group", group_size: int = 16,
                 per_token_size: int = 256, mem_length: int = 16, retrieval_layers: int = 2,
                 use_timestep_pe: bool = True, fusion_type: str = "gate", consolidate_type: str = "tome",
                 update_fused: bool = True, mem_dropout: float = 0.1, **kwargs):
        super().__init__(**kwargs)
        assert dataloader_type in ("stream", "group", "parallel_stream")       # memvla_arch.py:210-212
        assert fusion_type in ("gate", "add") and consolidate_type in ("fifo", "tome")
        self.action_model_type, self.action_dim, self.chunk_size = action_model_type, action_dim, chunk_size
        self.dataloader_type, self.group_size, self.per_token_size = dataloader_type, group_size, per_token_size
        self.mem_length, self.retrieval_layers, self.use_timestep_pe = mem_length, retrieval_layers, use_timestep_pe
        self.fusion_type, self.consolidate_type, self.update_fused = fusion_type, consolidate_type, update_fused
        # CrossTransformerBlock hard-codes dropout=0.1 (memvla_arch.py:83); exposed so parity runs can set 0
        self.mem_dropout = mem_dropout
        self.freeze_action_head = kwargs.get("freeze_action_head", False)


def memory_specs(cfg: MemVLAConfig, d: int, trainable: bool = True) -> list[ParamSpec]:
    """per_compr + per_cog_mem_bank parameters in the reference's registration order."""
    g = "memory"
    P = lambda n, s, **k: ParamSpec(n, s, g, trainable=trainable, **k)  # noqa: E731
    c_out = cfg.per_token_size
    h_se, h_mlp = max(1, d // 16), max(1, int(d * 0.5))                      # memvla_arch.py:143,153
    p = "model.per_compr."
    sp = [P(p + "excite.1.weight", (h_se, d, 1, 1)), P(p + "excite.1.bias", (h_se,)),
          P(p + "excite.3.weight", (d, h_se, 1, 1)), P(p + "excite.3.bias", (d,)),
          P(p + "reduce.0.weight", (h_mlp, d, 1, 1)), P(p + "reduce.0.bias", (h_mlp,)),
          P(p + "reduce.2.weight", (c_out, h_mlp, 1, 1)), P(p + "reduce.2.bias", (c_out,))]
    dims = {"per": c_out, "cog": d}
    m = "model.per_cog_mem_bank."
    for r in ROLES:
        D = dims[r]
        for l in range(cfg.retrieval_layers):
            q = f"{m}retrieval_blocks.{r}.{l}."
            for n in ("q_proj", "k_proj", "v_proj"):
                sp += [P(q + n + ".weight", (D, D)), P(q + n + ".bias", (D,))]
            sp += [P(q + "attn_norm.weight", (D,)), P(q + "attn_norm.bias", (D,)),
                   P(q + "ffn.0.weight", (4 * D, D)), P(q + "ffn.0.bias", (4 * D,)),
                   P(q + "ffn.3.weight", (D, 4 * D)), P(q + "ffn.3.bias", (D,)),
                   P(q + "ffn_norm.weight", (D,)), P(q + "ffn_norm.bias", (D,))]
    if cfg.fusion_type == "gate":
        for r in ROLES:
            q = f"{m}gate_fusion_blocks.{r}.proj."
            sp += [P(q + "weight", (dims[r], 2 * dims[r])), P(q + "bias", (dims[r],))]
    if cfg.use_timestep_pe:
        for r in ROLES:
            q = f"{m}timestep_embedders.{r}.mlp."
            sp += [P(q + "0.weight", (dims[r], 256)), P(q + "0.bias", (dims[r],)),
                   P(q + "2.weight", (dims[r], dims[r])), P(q + "2.bias", (dims[r],))]
    return sp


def _conv_lin(store: ParamStore, name: str) -> Lin:
    """1x1 Conv2d parameters [out, in, 1, 1] seen as the [out, in] GEMM weight they are."""
    w, gw = store.w(name + ".weight"), store.g(name + ".weight")
    o, i = w.shape[0], w.shape[1]
    return Lin(w.view(o, i), store.w(name + ".bias"), None if gw is None else gw.view(o, i), store.g(name + ".bias"))


class BottleneckSE:
    """memvla_arch.py:136-173 on the channel-last token map."""

    def __init__(self, store: ParamStore, prefix: str = "model.per_compr."):
        self.store = store
        self.e1, self.e3 = _conv_lin(store, prefix + "excite.1"), _conv_lin(store, prefix + "excite.3")
        self.r0, self.r2 = _conv_lin(store, prefix + "reduce.0"), _conv_lin(store, prefix + "reduce.2")

    def __call__(self, x3d: torch.Tensor) -> torch.Tensor:
        B, N, Cc = x3d.shape
        assert int(round(N ** 0.5)) ** 2 == N, "Input feature has no spatial structure"   # :160
        x = SEGateFn.apply(x3d, self.e1, self.e3, self.store)                               # :164-165
        h = LinearFn.apply(x.view(B * N, Cc), self.r0, "relu", self.store, True, None)      # :167
        return LinearFn.apply(h, self.r2, None, self.store, True, None).view(B, N, -1)


class _CrossBlock:
    """CrossTransformerBlock (memvla_arch.py:82-133): q/k/v projections, 4-head attention WITHOUT an output
    projection, post-LN residual, erf-GELU FFN with two dropouts, post-LN residual."""

    def __init__(self, store: ParamStore, prefix: str, heads: int = 4):
        L = lambda n: Lin.of(store, prefix + n + ".weight", prefix + n + ".bias")  # noqa: E731
        N_ = lambda n: Norm("ln", 1e-5, store.w(prefix + n + ".weight"), store.w(prefix + n + ".bias"),  # noqa: E731
                            store.g(prefix + n + ".weight"), store.g(prefix + n + ".bias"))
        self.store, self.heads = store, heads
        self.q, self.k, self.v = L("q_proj"), L("k_proj"), L("v_proj")
        self.f0, self.f3 = L("ffn.0"), L("ffn.3")
        self.attn_norm, self.ffn_norm = N_("attn_norm"), N_("ffn_norm")

    def _lin(self, x, lin: Lin, act=None):
        # inputs that carry no gradient (detached memory) still need this backward to run for the weight gradients
        rg = x.requires_grad
        return LinearFn.apply(x, lin, act, self.store, rg, None if rg else self.anchor_t)

    def __call__(self, query, k_in, v_in, p_attn: float, p_ffn: float, seeds):
        """query [N, D]; k_in, v_in [M, D] (one sample).  seeds: (attention, ffn-hidden, ffn-out) dropout seeds."""
        st = self.store
        N, M = query.shape[0], k_in.shape[0]
        q, k, v = self._lin(query, self.q), self._lin(k_in, self.k), self._lin(v_in, self.v)
        a = CrossAttnFn.apply(q, k, v, 1, N, M, self.heads, p_attn, seeds[0])
        x = NormFn.apply(query + a, self.attn_norm, st)
        h = self._lin(x, self.f0, "gelu")
        if p_ffn > 0.0:
            h = DropoutFn.apply(h, p_ffn, seeds[1])
        f = self._lin(h, self.f3)
        if p_ffn > 0.0:
            f = DropoutFn.apply(f, p_ffn, seeds[2])
        return NormFn.apply(x + f, self.ffn_norm, st)


KeyT = Tuple[int, ...]


class PerCogMemBank:
    """memvla_arch.py:195-427.  `training` mirrors nn.Module.training (episode-id handling differs, :325-345)."""

    def __init__(self, store: ParamStore, cfg: MemVLAConfig, d: int, prefix: str = "model.per_cog_mem_bank."):
        self.store, self.cfg = store, cfg
        self.dims = {"per": cfg.per_token_size, "cog": d}
        self.blocks = {r: [_CrossBlock(store, f"{prefix}retrieval_blocks.{r}.{l}.") for l in range(cfg.retrieval_layers)]
                       for r in ROLES}
        self.gates = ({r: Lin.of(store, f"{prefix}gate_fusion_blocks.{r}.proj.weight",
                                 f"{prefix}gate_fusion_blocks.{r}.proj.bias") for r in ROLES}
                      if cfg.fusion_type == "gate" else None)
        self.temb = ({r: (Lin.of(store, f"{prefix}timestep_embedders.{r}.mlp.0.weight",
                                 f"{prefix}timestep_embedders.{r}.mlp.0.bias"),
                          Lin.of(store, f"{prefix}timestep_embedders.{r}.mlp.2.weight",
                                 f"{prefix}timestep_embedders.{r}.mlp.2.bias")) for r in ROLES}
                     if cfg.use_timestep_pe else None)
        self.training = True
        self._seed = 0x5EED
        self.anchor_t = None          # set by MemVLAModel: scalar that makes weight-gradient-only backward passes run
        self.reset()

    def reset(self):                                                         # :238-245
        self.banks: Dict[str, Dict[KeyT, List[Tuple[Optional[torch.Tensor], torch.Tensor]]]] = {r: {} for r in ROLES}
        self.prev_eids: Dict[str, Dict[int, KeyT]] = {r: {} for r in ROLES}
        self.eid_stream: Dict[str, Optional[KeyT]] = {r: None for r in ROLES}

    def clear_episode(self, role: str, episode_id: KeyT):
        self.banks[role].pop(episode_id, None)

    # -------------------------------------------------------------- consolidation (:247-296), no gradient
    @torch.no_grad()
    def _consolidate_with_token_merge(self, role: str, episode_id: KeyT):
        bank = self.banks[role].get(episode_id, [])
        T = len(bank)
        if T < 2:
            return
        feats = torch.stack([f.float() for _, f in bank])                    # [T, N, D]
        sims = torch.nn.functional.cosine_similarity(feats[:-1], feats[1:], dim=-1).mean(dim=-1)
        j = int(sims.argmax().item())                                        # first maximum wins, as torch.argmax
        (ti, fi), (tj, fj) = bank[j], bank[j + 1]
        bank[j] = (0.5 * (ti + tj) if ti is not None else None, (0.5 * (fi + fj)).detach().clone())
        bank.pop(j + 1)

    @torch.no_grad()
    def _memory_consolidate(self, role: str, episode_id: KeyT, feat: torch.Tensor, timestep):
        bank = self.banks[role].setdefault(episode_id, [])
        bank.append((timestep, feat.detach().clone()))
        while len(bank) > self.cfg.mem_length:
            if self.cfg.consolidate_type == "fifo":
                del bank[:-self.cfg.mem_length]
            else:
                self._consolidate_with_token_merge(role, episode_id)

    def _encode_time(self, role: str, t: torch.Tensor) -> torch.Tensor:     # TimestepEmbedder (:37-79)
        l0, l2 = self.temb[role]
        dt = l0.w.dtype
        e = ops.timestep_embedding(t.float().reshape(-1), 256, dt)
        anchor = self.anchor_t
        return LinearFn.apply(LinearFn.apply(e, l0, "silu", self.store, False, anchor), l2, None, self.store, True, None)

    def _next_seeds(self):
        self._seed += 3
        return (self._seed, self._seed + 1, self._seed + 2)

    # -------------------------------------------------------------- :298-407
    def _process_batch(self, role: str, tokens: torch.Tensor, episode_ids, timesteps: List[torch.Tensor]):
        cfg = self.cfg
        B, N, D = tokens.shape
        dl = cfg.dataloader_type
        if self.training:
            if dl == "group":
                self.banks[role].clear()
                self.prev_eids[role].clear()
                self.eid_stream[role] = None
            elif dl == "stream":
                first, prev = episode_ids[0], self.eid_stream[role]
                if prev is not None and prev != first:
                    self.clear_episode(role, prev)
                self.eid_stream[role] = first
            else:
                episode_ids = [(i, e[0], e[1]) for i, e in enumerate(episode_ids)]
        else:
            episode_ids = [(0, 0)] * B if dl in ("group", "stream") else [(i, 0, 0) for i in range(B)]
        p = cfg.mem_dropout
        outs = []
        for i in range(B):
            eid = episode_ids[i]
            if self.training and dl == "stream" and i > 0 and episode_ids[i] != episode_ids[i - 1]:
                self.clear_episode(role, episode_ids[i - 1])
                self.eid_stream[role] = episode_ids[i]
            if self.training and dl == "parallel_stream":
                prev = self.prev_eids[role].get(i)
                if prev is not None and prev != eid:
                    self.clear_episode(role, prev)
                self.prev_eids[role][i] = eid
            work = tokens[i]                                                  # [N, D]
            hist = self.banks[role].get(eid, [])
            if hist:
                mem = torch.stack([f for _, f in hist]).reshape(-1, D)        # [T*N, D], detached
                if cfg.use_timestep_pe:
                    pe = self._encode_time(role, torch.stack([t for t, _ in hist]))       # [T, D]
                    k_in = (mem.view(len(hist), N, D) + pe[:, None, :]).reshape(-1, D)
                else:
                    k_in = mem
            else:
                mem = work
                k_in = work + self._encode_time(role, timesteps[i].reshape(1)) if cfg.use_timestep_pe else work
            q = work
            for blk in self.blocks[role]:
                # F.scaled_dot_product_attention(dropout_p=...) is unconditional in the reference (:122-124): the
                # attention dropout is active in eval too; nn.Dropout inside ffn follows module.training
                q = blk(q, k_in, mem, p, p if self.training else 0.0, self._next_seeds())
            if cfg.fusion_type == "add":
                fused = (work + q) * 0.5
            else:
                z = LinearFn.apply(torch.cat([work, q], dim=-1), self.gates[role], None, self.store, True, None)
                fused = GateFuseFn.apply(z, work, q)
            outs.append(fused)
            self._memory_consolidate(role, eid, fused if cfg.update_fused else tokens[i],
                                     timesteps[i] if cfg.use_timestep_pe else None)
        return torch.stack(outs, dim=0)

    def process_batch_per(self, per_tokens, episode_ids, timesteps):
        return self._process_batch("per", per_tokens, episode_ids, timesteps)

    def process_batch_cog(self, cog_tokens, episode_ids, timesteps):
        return self._process_batch("cog", cog_tokens, episode_ids, timesteps)


class StripClsFn(torch.autograd.Function):
    """The projector runs on [CLS | P patches] rows per image; `vision_proj_feats` (capture_projected_vision,
    memvla_arch.py:748-759) is the patch part: [n_img*(P+1), D] -> [n_img, P, D]."""

    @staticmethod
    def forward(ctx, feats2d, n_img, P):
        D = feats2d.shape[1]
        out = torch.empty((n_img, P, D), device=feats2d.device, dtype=feats2d.dtype)
        ops.copy3d_(feats2d, out, n_img, P, D, (P + 1) * D, D, P * D, D, src_off=D)
        ctx.geom = (n_img, P, D)
        return out

    @staticmethod
    def backward(ctx, g):
        n_img, P, D = ctx.geom
        dx = torch.zeros((n_img * (P + 1), D), device=g.device, dtype=g.dtype)
        ops.copy3d_(g.contiguous(), dx, n_img, P, D, P * D, D, (P + 1) * D, D, dst_off=D)
        return dx, None, None


class MemVLAModel(DexboticVLMModel):
    """memvla_arch.py:430-533 (`model.model`)."""
    action_head_prefix = "action_head"

    def __init__(self, store, config: MemVLAConfig):
        super().__init__(store, config)
        d = cfg_get(config.llm_config, "hidden_size")
        self.per_compr = BottleneckSE(store)
        self.per_cog_mem_bank = PerCogMemBank(store, config, d)
        self.per_cog_mem_bank.anchor_t = self.anchor.t
        for blks in self.per_cog_mem_bank.blocks.values():
            for b in blks:
                b.anchor_t = self.anchor.t
        self.action_head = ActionModel(store, config.action_model_type, d, config.action_dim, config.chunk_size - 1,
                                       use_per_attn=True)

    @property
    def action_head_module(self):
        return self.action_head


class MemVLAForCausalLM(B200Module):
    """memvla_arch.py:536-759."""
    config_class = MemVLAConfig

    def __init__(self, config: MemVLAConfig, device="cuda"):
        super().__init__()
        self.config = config
        llm, vis = config.llm_config, config.mm_vision_tower
        d, V = cfg_get(llm, "hidden_size"), cfg_get(llm, "vocab_size")
        specs = (llm_specs(llm, trainable=not config.freeze_llm)
                 + clip_specs(vis, trainable=not config.freeze_mm_vision)
                 + projector_specs(config.mm_projector_type, cfg_get(vis, "hidden_size"), d,
                                   trainable=not config.freeze_mm_projector)
                 + memory_specs(config, d)
                 + action_head_specs(config.action_model_type, d, config.action_dim, config.chunk_size,
                                     trainable=not getattr(config, "freeze_action_head", False),
                                     per_token_size=config.per_token_size)
                 + [ParamSpec("lm_head.weight", (V, d), "lm_head", trainable=False)])
        store = self._materialize(specs, device)
        for name in store.order:
            s = store.slots[name]
            if s.region == "B":
                self.get_parameter(name).grad = store.g(name)
        self.model_engine = MemVLAModel(store, config)
        self.cur_timestep = 0                                                 # :542, inference only

    @property
    def engine(self) -> MemVLAModel:
        return self.model_engine

    def _after_weights_changed(self) -> None:
        self.model_engine.refresh()

    def train(self, mode: bool = True):
        super().train(mode)
        if hasattr(self, "model_engine"):
            self.model_engine.per_cog_mem_bank.training = mode
        return self

    def _trunk(self, input_ids, attention_mask, labels, images):
        """VLM forward + captured projector output (the reference's forward hook, :566-589)."""
        eng = self.model_engine
        if images.dim() == 5:
            # the hook output is [B*n_view, P, D] there and process_batch_per then indexes episode_ids[i] for
            # i >= B (memvla_arch.py:351-353): multi-view input is not a working configuration of the reference
            raise NotImplementedError("MemVLA takes single-view images [B,3,H,W] (memvla_arch.py:566-626)")
        emb, new_labels, new_mask, pos, S, _ = eng._prepare_inputs_labels_for_multimodal(
            input_ids, attention_mask, labels, images)
        B = input_ids.shape[0]
        P = eng.mm_vision_tower.P
        vision_proj_feats = StripClsFn.apply(eng.last_image_features, B, P)
        hidden2d = eng.llm.forward(emb.view(B * S, -1), B, S, new_mask, pos)
        return hidden2d, new_mask, S, vision_proj_feats

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
                indexes: List[int] = None,
                noise: Optional[torch.Tensor] = None,          # parity hooks: inject the reference's random draws
                timesteps: Optional[torch.Tensor] = None,
                drop_mask: Optional[torch.Tensor] = None,
                **kwargs) -> CausalLMOutputDexbotic:
        if images is None or input_ids is None:
            raise NotImplementedError("MemVLA forward needs input_ids and images (memvla_arch.py:566-589)")
        if not input_ids.is_cuda:
            raise RuntimeError("dexbotic_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        eng, cfg = self.model_engine, self.config
        B = input_ids.shape[0]
        hidden2d, new_mask, S, vision_proj_feats = self._trunk(input_ids, attention_mask, labels, images)
        last_hidden_state = hidden2d.view(B, S, -1)
        loss = None
        if attention_mask is not None and actions is not None:
            idx = ops.last_valid_index(new_mask)                                           # :606-616
            cog_tokens = GatherRowsFn.apply(hidden2d, idx)[:, None, :]                      # [B, 1, D]
            per_tokens = eng.per_compr(vision_proj_feats)                                   # :618
            episode_ids = [tuple(item[:2]) for item in indexes]                             # :620-621
            ts = [torch.tensor(float(item[2]), device=hidden2d.device) for item in indexes]
            bank = eng.per_cog_mem_bank
            cog_tokens = bank.process_batch_cog(cog_tokens, episode_ids, ts)               # :623-633
            per_tokens = bank.process_batch_per(per_tokens, episode_ids, ts)
            R = repeated_diffusion_steps
            a = actions.reshape(B, -1, cfg.action_dim).to(torch.float32)[:, :cfg.chunk_size, :]
            cog32 = CastFn.apply(cog_tokens.contiguous(), torch.float32)
            per32 = CastFn.apply(per_tokens.contiguous(), torch.float32)
            loss = eng.action_head.loss(self, a.repeat(R, 1, 1), cog32.repeat(R, 1, 1), noise, timesteps, drop_mask,
                                        training=self.training, per_token=per32, groups=R)   # :637-652
        out = CausalLMOutputDexbotic(loss=loss, logits=last_hidden_state)
        out.vision_proj_feats = vision_proj_feats                                           # :663
        return out

    @torch.no_grad()
    def inference_action(self, input_ids, image_tensor, episode_first_frame, inference_args={},
                         noise: Optional[torch.Tensor] = None, **kwargs):
        """memvla_arch.py:666-745: stateful across calls (memory bank + cur_timestep); batch size 1 per episode."""
        cfg_scale = inference_args.get("cfg_scale", 1.5)
        num_ddim_steps = inference_args.get("num_ddim_steps", 10)
        action_norms = inference_args.get("action_norms")
        assert episode_first_frame in ("True", "False"), "episode_first_frame must be 'True' or 'False'"
        eng = self.model_engine
        if episode_first_frame == "True":
            eng.per_cog_mem_bank.reset()
            self.cur_timestep = 0
        hidden2d, _, S, vision_proj_feats = self._trunk(input_ids, None, None, image_tensor)
        B = input_ids.shape[0]
        cog_tokens = hidden2d.view(B, S, -1)[:, -1, :][:, None, :].contiguous()             # :687
        per_tokens = eng.per_compr(vision_proj_feats)
        ts = [torch.tensor(float(self.cur_timestep), device=hidden2d.device)]
        self.cur_timestep += 1
        cog_tokens = eng.per_cog_mem_bank.process_batch_cog(cog_tokens, [(0, 0)], ts)
        per_tokens = eng.per_cog_mem_bank.process_batch_per(per_tokens, [(0, 0)], ts)
        if noise is None:
            noise = torch.randn(B, self.config.chunk_size, self.config.action_dim, device=hidden2d.device)
        samples = eng.action_head.sample(self, cog_tokens.float().contiguous(), noise.float(), cfg_scale,
                                         num_ddim_steps, per_token=per_tokens.float().contiguous())
        actions = np.clip(samples[0].float().cpu().numpy(), -1, 1)                          # _denorm
        mn = np.array(action_norms["min"]).reshape(1, -1)
        mx = np.array(action_norms["max"]).reshape(1, -1)
        return (mn + (actions + 1) * 0.5 * (mx - mn)).tolist()

    def zero_grad(self, set_to_none: bool = False):        # noqa: D401
        self.store.zero_grad()

    def optimizer_step(self, base_lr: float = 2e-5, mm_projector_lr=None, mm_vision_lr=None, action_head_lr=None,
                       betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0, max_grad_norm=1.0):
        lrs = {"llm": base_lr, "memory": base_lr, "projector": mm_projector_lr or base_lr,
               "vision": mm_vision_lr or base_lr, "action_head": action_head_lr or base_lr, "lm_head": base_lr}
        norm = self.store.adamw_step(lrs, betas, eps, weight_decay, max_grad_norm)
        self.model_engine.refresh()
        return norm
