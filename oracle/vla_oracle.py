"""TEST INFRASTRUCTURE — CPU restatement (plain fp32 PyTorch / numpy) of the reference's VLA hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module, and only as the checker (or as the timed CPU baseline) — never on the product path.

Parity pin: validated in this container against the UNMODIFIED reference (oracle/ref_loader.py) by
oracle/make_golden.py, which also writes the committed fixtures under tests/golden/.  The third-party
arithmetic the reference calls into (HF transformers Qwen2/Llama/CLIP, timm Attention/Mlp — un-vendored;
pins transformers==4.51.0/4.54.0 in /root/reference/Dockerfile:37, dockerfiles/Dockerfile.c130t28:22; timm
un-pinned) is restated from its published algorithm and checked against transformers 5.5.0 (the version in
this image) + the timm restatement in ref_loader — "parity unpinned" at the timm boundary.

Every function works on a flat state_dict with the reference's parameter names
(model.llm.*, model.mm_vision_tower.*, model.mm_projector.*, model.action_head.*).
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F

IGNORE_INDEX = -100        # dexbotic/constants.py
IMAGE_TOKEN_INDEX = -200   # dexbotic/constants.py


# ----------------------------------------------------------------------------------------------
# diffusion schedule — cogact/action_model/diffusion.py:205-209 (squaredcos_cap_v2), :214-231
# (betas_for_alpha_bar), GaussianDiffusion.__init__ (alphas_cumprod in float64)
# ----------------------------------------------------------------------------------------------
def cosine_schedule(num_steps: int = 100, max_beta: float = 0.999):
    def alpha_bar(t):
        return math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2

    betas = []
    for i in range(num_steps):
        t1, t2 = i / num_steps, (i + 1) / num_steps
        betas.append(min(1 - alpha_bar(t2) / alpha_bar(t1), max_beta))
    betas = np.array(betas, dtype=np.float64)
    alphas_cumprod = np.cumprod(1.0 - betas, axis=0)
    return np.sqrt(alphas_cumprod), np.sqrt(1.0 - alphas_cumprod)


def q_sample(x_start, t, noise, sqrt_ac, sqrt_1mac):
    """diffusion.py:308-326; _extract_into_tensor (:975-987) casts the float64 table entry to float32."""
    a = torch.from_numpy(sqrt_ac)[t].float()[:, None, None]
    b = torch.from_numpy(sqrt_1mac)[t].float()[:, None, None]
    return a * x_start + b * noise


# ----------------------------------------------------------------------------------------------
# CLIP vision tower — modules/mm_vision/clip/clip_encoder.py:31-57 calling HF CLIPVisionModel
# ----------------------------------------------------------------------------------------------
def quick_gelu(x):
    return x * torch.sigmoid(1.702 * x)


ACT = {
    "quick_gelu": quick_gelu,
    "gelu": F.gelu,
    "gelu_pytorch_tanh": lambda x: F.gelu(x, approximate="tanh"),
    "gelu_tanh": lambda x: F.gelu(x, approximate="tanh"),
    "silu": F.silu,
}


def _mha(x, wq, bq, wk, bk, wv, bv, wo, bo, heads, mask=None):
    B, S, D = x.shape
    hd = D // heads
    q = F.linear(x, wq, bq).view(B, S, heads, hd).transpose(1, 2)
    k = F.linear(x, wk, bk).view(B, S, heads, hd).transpose(1, 2)
    v = F.linear(x, wv, bv).view(B, S, heads, hd).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) * hd ** -0.5
    if mask is not None:
        s = s.masked_fill(~mask, float("-inf"))
    p = torch.softmax(s, dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B, S, D)
    return F.linear(o, wo, bo)


def _mha_cross(xq, xk, xv, wq, bq, wk, bk, wv, bv, wo, bo, heads):
    """softmax((xq Wq)(xk Wk)^T / sqrt(hd)) (xv Wv) [Wo]: nn.MultiheadAttention(batch_first) and the SDPA call of
    CrossTransformerBlock (memvla_arch.py:111-124), dropout 0."""
    B, N, D = xq.shape
    M = xk.shape[1]
    hd = D // heads
    q = F.linear(xq, wq, bq).view(B, N, heads, hd).transpose(1, 2)
    k = F.linear(xk, wk, bk).view(B, M, heads, hd).transpose(1, 2)
    v = F.linear(xv, wv, bv).view(B, M, heads, hd).transpose(1, 2)
    p = torch.softmax((q @ k.transpose(-1, -2)) * hd ** -0.5, dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B, N, D)
    return o if wo is None else F.linear(o, wo, bo)


def clip_vision_features(sd, prefix: str, images: torch.Tensor, cfg: dict) -> torch.Tensor:
    """hidden_states[-2][:, 1:] of HF CLIPVisionModel (feature_select, clip_encoder.py:31-36)."""
    p = prefix + "vision_tower.vision_model."
    patch, heads, eps = cfg["patch_size"], cfg["num_attention_heads"], cfg.get("layer_norm_eps", 1e-5)
    act = ACT[cfg.get("hidden_act", "quick_gelu")]
    x = F.conv2d(images, sd[p + "embeddings.patch_embedding.weight"], stride=patch)     # [B, D, h, w]
    x = x.flatten(2).transpose(1, 2)
    cls = sd[p + "embeddings.class_embedding"].expand(x.shape[0], 1, -1)
    x = torch.cat([cls, x], dim=1) + sd[p + "embeddings.position_embedding.weight"][None]
    D = x.shape[-1]
    x = F.layer_norm(x, (D,), sd[p + "pre_layrnorm.weight"], sd[p + "pre_layrnorm.bias"], eps)
    n_layers = cfg["num_hidden_layers"]
    for i in range(n_layers - 1):       # select_layer = -2: the last encoder layer's output is never used
        q = f"{p}encoder.layers.{i}."
        h = F.layer_norm(x, (D,), sd[q + "layer_norm1.weight"], sd[q + "layer_norm1.bias"], eps)
        x = x + _mha(h, sd[q + "self_attn.q_proj.weight"], sd[q + "self_attn.q_proj.bias"],
                     sd[q + "self_attn.k_proj.weight"], sd[q + "self_attn.k_proj.bias"],
                     sd[q + "self_attn.v_proj.weight"], sd[q + "self_attn.v_proj.bias"],
                     sd[q + "self_attn.out_proj.weight"], sd[q + "self_attn.out_proj.bias"], heads)
        h = F.layer_norm(x, (D,), sd[q + "layer_norm2.weight"], sd[q + "layer_norm2.bias"], eps)
        x = x + F.linear(act(F.linear(h, sd[q + "mlp.fc1.weight"], sd[q + "mlp.fc1.bias"])),
                         sd[q + "mlp.fc2.weight"], sd[q + "mlp.fc2.bias"])
    return x[:, 1:]


def mlp_projector(sd, prefix: str, x: torch.Tensor, depth: int = 2) -> torch.Tensor:
    """mlpNx_gelu: Linear -> (GELU(erf) -> Linear)*(N-1)   (mm_projector/builder.py:69-79)."""
    x = F.linear(x, sd[prefix + "0.weight"], sd[prefix + "0.bias"])
    for i in range(1, depth):
        x = F.linear(F.gelu(x), sd[f"{prefix}{2 * i}.weight"], sd[f"{prefix}{2 * i}.bias"])
    return x


# ----------------------------------------------------------------------------------------------
# image-token splice — dexbotic_arch.py:182-373
# ----------------------------------------------------------------------------------------------
def splice(embed_weight, image_features, input_ids, attention_mask, labels, max_len: Optional[int],
           padding_side: str = "right"):
    """Returns inputs_embeds [B,S,D], labels [B,S], attention_mask [B,S] (bool), position_ids [B,S]."""
    B = input_ids.shape[0]
    if attention_mask is None:
        attention_mask = torch.ones_like(input_ids, dtype=torch.bool)
    attention_mask = attention_mask.bool()
    if labels is None:
        labels = torch.full_like(input_ids, IGNORE_INDEX)
    new_embeds, new_labels = [], []
    img_idx = 0
    for b in range(B):
        ids = input_ids[b][attention_mask[b]]
        lab = labels[b][attention_mask[b]]
        pos = (ids == IMAGE_TOKEN_INDEX).nonzero().flatten().tolist()
        if len(pos) == 0:                                            # :264-272
            new_embeds.append(embed_weight[ids])
            new_labels.append(lab)
            img_idx += 1
            continue
        bounds = [-1] + pos + [ids.shape[0]]
        e, l = [], []
        for i in range(len(bounds) - 1):
            seg = slice(bounds[i] + 1, bounds[i + 1])
            e.append(embed_weight[ids[seg]])
            l.append(lab[seg])
            if i < len(pos):
                feat = image_features[img_idx]
                e.append(feat)
                l.append(torch.full((feat.shape[0],), IGNORE_INDEX, dtype=lab.dtype))
                img_idx += 1
        new_embeds.append(torch.cat(e))
        new_labels.append(torch.cat(l))
    if max_len is not None:                                          # :236-243
        new_embeds = [x[:max_len] for x in new_embeds]
        new_labels = [x[:max_len] for x in new_labels]
    S = max(x.shape[0] for x in new_embeds)
    D = new_embeds[0].shape[1]
    emb = torch.zeros(B, S, D, dtype=new_embeds[0].dtype)
    lab = torch.full((B, S), IGNORE_INDEX, dtype=labels.dtype)
    msk = torch.zeros(B, S, dtype=torch.bool)
    pid = torch.zeros(B, S, dtype=torch.long)
    for b in range(B):                                               # :315-373
        n = new_embeds[b].shape[0]
        if n == 0:
            continue
        sl = slice(S - n, S) if padding_side == "left" else slice(0, n)
        emb[b, sl] = new_embeds[b]
        lab[b, sl] = new_labels[b]
        msk[b, sl] = True
        pid[b, sl] = torch.arange(n)
    return emb, lab, msk, pid


# ----------------------------------------------------------------------------------------------
# decoder — HF Qwen2Model / LlamaModel as built by AutoModel at dexbotic_arch.py:55-62
# ----------------------------------------------------------------------------------------------
def rms_norm(x, w, eps, unit_offset=False):
    n = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)
    return n * (1 + w) if unit_offset else w * n


def rope_tables(positions: torch.Tensor, head_dim: int, theta: float):
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.float32) / head_dim))
    f = positions.float()[..., None] * inv
    emb = torch.cat([f, f], dim=-1)
    return emb.cos(), emb.sin()


def rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat([-x[..., h:], x[..., :h]], dim=-1)


def decoder_layer(sd, q: str, x, cos, sin, allow, cfg: dict):
    B, S, D = x.shape
    H, KVH = cfg["num_attention_heads"], cfg["num_key_value_heads"]
    hd = cfg.get("head_dim") or D // H
    eps = cfg["rms_norm_eps"]
    h = rms_norm(x, sd[q + "input_layernorm.weight"], eps)
    qh = F.linear(h, sd[q + "self_attn.q_proj.weight"], sd.get(q + "self_attn.q_proj.bias")).view(B, S, H, hd)
    kh = F.linear(h, sd[q + "self_attn.k_proj.weight"], sd.get(q + "self_attn.k_proj.bias")).view(B, S, KVH, hd)
    vh = F.linear(h, sd[q + "self_attn.v_proj.weight"], sd.get(q + "self_attn.v_proj.bias")).view(B, S, KVH, hd)
    qh, kh, vh = qh.transpose(1, 2), kh.transpose(1, 2), vh.transpose(1, 2)
    c, s = cos[:, None], sin[:, None]
    qh = qh * c + rotate_half(qh) * s
    kh = kh * c + rotate_half(kh) * s
    G = H // KVH
    kh = kh.repeat_interleave(G, dim=1)
    vh = vh.repeat_interleave(G, dim=1)
    sc = (qh @ kh.transpose(-1, -2)) * hd ** -0.5
    sc = sc.masked_fill(~allow, float("-inf"))
    p = torch.nan_to_num(torch.softmax(sc, dim=-1), nan=0.0)
    o = (p @ vh).transpose(1, 2).reshape(B, S, H * hd)
    x = x + F.linear(o, sd[q + "self_attn.o_proj.weight"], sd.get(q + "self_attn.o_proj.bias"))
    h = rms_norm(x, sd[q + "post_attention_layernorm.weight"], eps)
    act = ACT[cfg.get("hidden_act", "silu")]
    m = F.linear(act(F.linear(h, sd[q + "mlp.gate_proj.weight"])) * F.linear(h, sd[q + "mlp.up_proj.weight"]),
                 sd[q + "mlp.down_proj.weight"])
    return x + m


def decoder_forward(sd, prefix: str, inputs_embeds, attention_mask, position_ids, cfg: dict, collect=None):
    """Causal decoder with key-padding mask; returns the final-norm hidden states [B,S,D]."""
    B, S, D = inputs_embeds.shape
    hd = cfg.get("head_dim") or D // cfg["num_attention_heads"]
    cos, sin = rope_tables(position_ids, hd, cfg.get("rope_theta", 10000.0))
    causal = torch.tril(torch.ones(S, S, dtype=torch.bool))
    allow = causal[None, None] & attention_mask.bool()[:, None, None, :]
    x = inputs_embeds
    for i in range(cfg["num_hidden_layers"]):
        x = decoder_layer(sd, f"{prefix}layers.{i}.", x, cos, sin, allow, cfg)
        if collect is not None:
            collect.append(x)
    return rms_norm(x, sd[prefix + "norm.weight"], cfg["rms_norm_eps"])


def cognition_features(last_hidden, attention_mask):
    """cogact_arch.py:110-120: hidden state at the first position where cumsum(mask) reaches its max."""
    cs = attention_mask.long().cumsum(dim=1)
    idx = (cs == cs.max(dim=1, keepdim=True)[0]).float().argmax(dim=1)
    return last_hidden[torch.arange(last_hidden.shape[0]), idx][:, None, :], idx


# ----------------------------------------------------------------------------------------------
# DiT action head — cogact/action_model/dit.py, action_models.py:102-125
# ----------------------------------------------------------------------------------------------
def timestep_embedding(t, dim=256, max_period=10000):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def dit_forward(sd, prefix: str, x, t, z, drop_mask, num_heads: int, per_token=None):
    """dit.py:273-292.  drop_mask[b]=True replaces z[b] by the learned `uncondition` (token_drop, :80-95).
    per_token [N,P,D_per] (MemVLA, memvla/action_model/dit.py:176-187,302-323): embedded once, then every block
    cross-attends to it (nn.MultiheadAttention on norm3(x)) between self-attention and the MLP."""
    p = prefix + "net."
    if per_token is not None:
        per_token = F.linear(per_token, sd[p + "per_token_embedder.linear.weight"],
                             sd[p + "per_token_embedder.linear.bias"])
    x = F.linear(x, sd[p + "x_embedder.linear.weight"], sd[p + "x_embedder.linear.bias"])
    te = timestep_embedding(t)
    te = F.linear(F.silu(F.linear(te, sd[p + "t_embedder.mlp.0.weight"], sd[p + "t_embedder.mlp.0.bias"])),
                  sd[p + "t_embedder.mlp.2.weight"], sd[p + "t_embedder.mlp.2.bias"])
    if drop_mask is not None:
        z = torch.where(drop_mask[:, None, None], sd[p + "z_embedder.uncondition"][None], z)
    ze = F.linear(z, sd[p + "z_embedder.linear.weight"], sd[p + "z_embedder.linear.bias"])
    c = te[:, None] + ze
    x = torch.cat([c, x], dim=1) + sd[p + "positional_embedding"]
    D = x.shape[-1]
    i = 0
    while f"{p}blocks.{i}.attn.qkv.weight" in sd:
        q = f"{p}blocks.{i}."
        h = F.layer_norm(x, (D,), None, None, 1e-6)
        wqkv, bqkv = sd[q + "attn.qkv.weight"], sd[q + "attn.qkv.bias"]
        x = x + _mha(h, wqkv[:D], bqkv[:D], wqkv[D:2 * D], bqkv[D:2 * D], wqkv[2 * D:], bqkv[2 * D:],
                     sd[q + "attn.proj.weight"], sd[q + "attn.proj.bias"], num_heads)
        if per_token is not None:
            h = F.layer_norm(x, (D,), sd[q + "norm3.weight"], sd[q + "norm3.bias"], 1e-6)
            wi, bi = sd[q + "per_attn.in_proj_weight"], sd[q + "per_attn.in_proj_bias"]
            x = x + _mha_cross(h, per_token, per_token, wi[:D], bi[:D], wi[D:2 * D], bi[D:2 * D], wi[2 * D:],
                               bi[2 * D:], sd[q + "per_attn.out_proj.weight"], sd[q + "per_attn.out_proj.bias"],
                               num_heads)
        h = F.layer_norm(x, (D,), None, None, 1e-6)
        x = x + F.linear(F.gelu(F.linear(h, sd[q + "mlp.fc1.weight"], sd[q + "mlp.fc1.bias"]), approximate="tanh"),
                         sd[q + "mlp.fc2.weight"], sd[q + "mlp.fc2.bias"])
        i += 1
    x = F.layer_norm(x, (D,), None, None, 1e-6)
    x = F.linear(x, sd[p + "final_layer.linear.weight"], sd[p + "final_layer.linear.bias"])
    return x[:, 1:]


DIT_HEADS = {384: 4, 768: 12, 1024: 16}   # action_models.py:48-58 (DiT-S/B/L)


# ----------------------------------------------------------------------------------------------
# CogACT training forward — cogact_arch.py:56-147
# ----------------------------------------------------------------------------------------------
def cogact_forward(sd, cfg: dict, input_ids, attention_mask, images, actions, noise, timesteps, drop_mask,
                   repeated_diffusion_steps: int = 4, labels=None):
    """Returns dict(loss, last_hidden, cognition, inputs_embeds, attention_mask, image_features).

    noise [R*B,T,A], timesteps [R*B], drop_mask [R*B] are injected (the reference draws them with
    torch.randn_like / randint / rand: action_models.py:106-109, dit.py:86-88)."""
    # [B, n_view, C, H, W] images: views are encoded in one batch and concatenated per sample (dexbotic_arch.py:163-175)
    tr = cogact_like_trunk(sd, cfg, input_ids, attention_mask, images, labels)
    feats, emb, lab, msk, pid = (tr["image_features"], tr["inputs_embeds"], tr["labels"], tr["attention_mask"],
                                 tr["position_ids"])
    hs, cog, idx = tr["last_hidden"], tr["cognition"], tr["cognition_index"]
    A, T = cfg["action_dim"], cfg["chunk_size"]
    a = actions.reshape(actions.shape[0], -1, A)[:, :T].float()
    a_rep = a.repeat(repeated_diffusion_steps, 1, 1)
    z_rep = cog.repeat(repeated_diffusion_steps, 1, 1)
    sa, sb = cosine_schedule(cfg.get("diffusion_steps", 100))
    x_t = q_sample(a_rep, timesteps, noise, sa, sb)
    width = sd["model.action_head.net.x_embedder.linear.weight"].shape[0]
    pred = dit_forward(sd, "model.action_head.", x_t, timesteps, z_rep, drop_mask, DIT_HEADS[width])
    loss = ((pred - noise) ** 2).mean()
    return dict(loss=loss, last_hidden=hs, cognition=cog, cognition_index=idx, inputs_embeds=emb,
                attention_mask=msk, position_ids=pid, labels=lab, image_features=feats, noise_pred=pred)


def hybrid_cogact_forward(sd, cfg: dict, input_ids, attention_mask, images, actions, labels, has_action, has_text,
                          noise, timesteps, drop_mask, repeated_diffusion_steps: int = 4):
    """HybridCogACTForCausalLM.forward (cogact/hybrid_cogact_arch.py:60-218): co-training of text and actions.
    text_loss = HF ForCausalLMLoss(lm_head(hidden), labels) * has_text.any()  — as written at :131-141, the labels of
    the rows without text are only masked when NO row has text (`~has_text.any()`), which leaves no target at all: the
    mean over zero targets is NaN there (reproduced, not repaired);
    action_loss = sum_n has_action[n] * mean_{t,a}((eps_pred - eps)^2) / (sum_n has_action[n] + 1e-6) over the R
    repeats (:182-187).  loss = text_loss + action_loss."""
    tr = cogact_like_trunk(sd, cfg, input_ids, attention_mask, images, labels)
    hs, cog, lab = tr["last_hidden"], tr["cognition"], tr["labels"]
    logits = F.linear(hs, sd["lm_head.weight"])
    ht = has_text.bool().view(-1)
    text_labels = lab.clone()
    if not bool(ht.any()):
        text_labels[~ht] = IGNORE_INDEX
    sl = logits[:, :-1].reshape(-1, logits.shape[-1]).float()
    st = text_labels[:, 1:].reshape(-1)
    text_loss = F.cross_entropy(sl, st, ignore_index=IGNORE_INDEX, reduction="mean") * ht.any().float()
    A, T, R = cfg["action_dim"], cfg["chunk_size"], repeated_diffusion_steps
    a = actions.reshape(actions.shape[0], -1, A)[:, :T].float()
    sa, sb = cosine_schedule(cfg.get("diffusion_steps", 100))
    x_t = q_sample(a.repeat(R, 1, 1), timesteps, noise, sa, sb)
    width = sd["model.action_head.net.x_embedder.linear.weight"].shape[0]
    pred = dit_forward(sd, "model.action_head.", x_t, timesteps, cog.repeat(R, 1, 1), drop_mask, DIT_HEADS[width])
    w = has_action.reshape(-1).float().repeat(R)
    action_loss = (((pred - noise) ** 2).mean(dim=[1, 2]) * w).sum() / (w.sum() + 1e-6)
    return dict(loss=text_loss + action_loss, text_loss=text_loss, action_loss=action_loss, last_hidden=hs,
                attention_mask=tr["attention_mask"])


# ----------------------------------------------------------------------------------------------
# MemVLA — memvla_arch.py:82-427 (memory modules), :546-664 (training forward)
# ----------------------------------------------------------------------------------------------
def bottleneck_se(sd, p: str, x: torch.Tensor) -> torch.Tensor:
    """BottleneckSE (memvla_arch.py:136-173): squeeze-excite over channels (token mean -> 1x1 conv -> ReLU -> 1x1 conv
    -> sigmoid), rescale, then a per-token 1x1-conv bottleneck MLP.  1x1 convs on the [B,C,H,W] view are per-token
    linears; x [B,N,C] -> [B,N,C_out]."""
    lin = lambda t, n: F.linear(t, sd[p + n + ".weight"].flatten(1), sd[p + n + ".bias"])  # noqa: E731
    w = torch.sigmoid(lin(F.relu(lin(x.mean(dim=1), "excite.1")), "excite.3"))
    x = x * w[:, None, :]
    return lin(F.relu(lin(x, "reduce.0")), "reduce.2")


def cross_transformer_block(sd, p: str, query, k, v, heads: int = 4):
    """CrossTransformerBlock.forward (memvla_arch.py:105-133), dropout 0: attention WITHOUT an output projection,
    post-LN residuals, erf-GELU FFN."""
    D = query.shape[-1]
    a = _mha_cross(query, k, v, sd[p + "q_proj.weight"], sd[p + "q_proj.bias"], sd[p + "k_proj.weight"],
                   sd[p + "k_proj.bias"], sd[p + "v_proj.weight"], sd[p + "v_proj.bias"], None, None, heads)
    x = F.layer_norm(query + a, (D,), sd[p + "attn_norm.weight"], sd[p + "attn_norm.bias"])
    f = F.linear(F.gelu(F.linear(x, sd[p + "ffn.0.weight"], sd[p + "ffn.0.bias"])), sd[p + "ffn.3.weight"],
                 sd[p + "ffn.3.bias"])
    return F.layer_norm(x + f, (D,), sd[p + "ffn_norm.weight"], sd[p + "ffn_norm.bias"])


class MemBankOracle:
    """PerCogMemBank (memvla_arch.py:195-427) for one role, `group` / `stream` / `parallel_stream` training semantics
    and the eval-mode episode-id rewrite.  State: banks[eid] = [(timestep, feat[N,D] detached), ...]."""

    def __init__(self, sd, prefix: str, role: str, mem_length: int, retrieval_layers: int, dataloader_type="group",
                 use_timestep_pe=True, fusion_type="gate", consolidate_type="tome", update_fused=True):
        self.sd, self.p, self.role = sd, prefix, role
        self.mem_length, self.layers, self.dl = mem_length, retrieval_layers, dataloader_type
        self.use_pe, self.fusion, self.consolidate, self.update_fused = (use_timestep_pe, fusion_type, consolidate_type,
                                                                        update_fused)
        self.banks, self.prev_eids, self.eid_stream = {}, {}, None

    def _pe(self, t):
        q = f"{self.p}timestep_embedders.{self.role}.mlp."
        e = timestep_embedding(t, 256)
        return F.linear(F.silu(F.linear(e, self.sd[q + "0.weight"], self.sd[q + "0.bias"])), self.sd[q + "2.weight"],
                        self.sd[q + "2.bias"])

    def _consolidate(self, eid, feat, timestep):
        bank = self.banks.setdefault(eid, [])
        bank.append((timestep, feat.detach().clone()))
        while len(bank) > self.mem_length:
            if self.consolidate == "fifo":
                del bank[:-self.mem_length]
                continue
            # token merge (:247-274): fuse the adjacent pair with the highest mean cosine similarity
            sims = [F.cosine_similarity(bank[i][1].flatten(1) if bank[i][1].dim() > 1 else bank[i][1][None],
                                        bank[i + 1][1].flatten(1) if bank[i + 1][1].dim() > 1 else bank[i + 1][1][None],
                                        dim=1).mean().item() for i in range(len(bank) - 1)]
            j = int(torch.tensor(sims).argmax().item())
            (ti, fi), (tj, fj) = bank[j], bank[j + 1]
            bank[j] = (0.5 * (ti + tj) if ti is not None else None, (0.5 * (fi + fj)).detach().clone())
            bank.pop(j + 1)

    def process_batch(self, tokens, episode_ids, timesteps, training=True):
        B, N, D = tokens.shape
        if training:
            if self.dl == "group":
                self.banks.clear(); self.prev_eids.clear(); self.eid_stream = None
            elif self.dl == "stream":
                if self.eid_stream is not None and self.eid_stream != episode_ids[0]:
                    self.banks.pop(self.eid_stream, None)
                self.eid_stream = episode_ids[0]
            else:
                episode_ids = [(i, e[0], e[1]) for i, e in enumerate(episode_ids)]
        else:
            episode_ids = [(0, 0)] * B if self.dl in ("group", "stream") else [(i, 0, 0) for i in range(B)]
        outs = []
        for i in range(B):
            eid = episode_ids[i]
            if training and self.dl == "stream" and i > 0 and episode_ids[i] != episode_ids[i - 1]:
                self.banks.pop(episode_ids[i - 1], None)
                self.eid_stream = episode_ids[i]
            if training and self.dl == "parallel_stream":
                prev = self.prev_eids.get(i)
                if prev is not None and prev != eid:
                    self.banks.pop(prev, None)
                self.prev_eids[i] = eid
            work = tokens[i][None]
            hist = self.banks.get(eid, [])
            if hist:
                mem = torch.stack([f for _, f in hist]).reshape(-1, D)[None]
                pe = (self._pe(torch.stack([t for t, _ in hist]))[None].repeat_interleave(N, dim=1) if self.use_pe
                      else torch.zeros_like(mem))
            else:
                mem = work
                pe = (self._pe(timesteps[i].reshape(1))[None].repeat_interleave(N, dim=1) if self.use_pe
                      else torch.zeros_like(mem))
            q = work
            for l in range(self.layers):
                q = cross_transformer_block(self.sd, f"{self.p}retrieval_blocks.{self.role}.{l}.", q, mem + pe, mem)
            if self.fusion == "add":
                fused = (work + q) * 0.5
            else:
                g = f"{self.p}gate_fusion_blocks.{self.role}.proj."
                sc = torch.sigmoid(F.linear(torch.cat([work, q], dim=-1), self.sd[g + "weight"], self.sd[g + "bias"]))
                fused = sc * work + (1 - sc) * q
            outs.append(fused)
            self._consolidate(eid, fused[0] if self.update_fused else tokens[i], timesteps[i] if self.use_pe else None)
        return torch.cat(outs, dim=0)


def memvla_forward(sd, cfg: dict, input_ids, attention_mask, images, actions, indexes, noise, timesteps, drop_mask,
                   repeated_diffusion_steps: int = 4, banks=None, training=True):
    """MemVLAForCausalLM.forward (memvla_arch.py:546-664), dropout 0.  indexes[b] = (dataset, episode, frame).
    `banks` (dict role -> MemBankOracle) carries state across calls for the stream modes / inference."""
    out = cogact_like_trunk(sd, cfg, input_ids, attention_mask, images)
    cog, feats = out["cognition"], out["image_features"]
    m = cfg["mem"]
    if banks is None:
        banks = {r: MemBankOracle(sd, "model.per_cog_mem_bank.", r, m["mem_length"], m["retrieval_layers"],
                                  m.get("dataloader_type", "group"), m.get("use_timestep_pe", True),
                                  m.get("fusion_type", "gate"), m.get("consolidate_type", "tome"),
                                  m.get("update_fused", True)) for r in ("per", "cog")}
    per = bottleneck_se(sd, "model.per_compr.", feats)
    eids = [tuple(i[:2]) for i in indexes]
    ts = [torch.tensor(i[2]) for i in indexes]
    cog_f = banks["cog"].process_batch(cog, eids, ts, training)
    per_f = banks["per"].process_batch(per, eids, ts, training)
    A, T = cfg["action_dim"], cfg["chunk_size"]
    R = repeated_diffusion_steps
    a = actions.reshape(actions.shape[0], -1, A)[:, :T].float()
    sa, sb = cosine_schedule(cfg.get("diffusion_steps", 100))
    x_t = q_sample(a.repeat(R, 1, 1), timesteps, noise, sa, sb)
    width = sd["model.action_head.net.x_embedder.linear.weight"].shape[0]
    pred = dit_forward(sd, "model.action_head.", x_t, timesteps, cog_f.repeat(R, 1, 1), drop_mask, DIT_HEADS[width],
                       per_token=per_f.repeat(R, 1, 1))
    out.update(loss=((pred - noise) ** 2).mean(), cog_fused=cog_f, per_fused=per_f, per_tokens=per, noise_pred=pred,
               banks=banks)
    return out


def cogact_like_trunk(sd, cfg: dict, input_ids, attention_mask, images, labels=None):
    """vision tower -> projector -> splice -> decoder -> cognition token (shared by CogACT and MemVLA)."""
    if images.dim() == 5:
        Bv, nv = images.shape[:2]
        feats = clip_vision_features(sd, "model.mm_vision_tower.", images.flatten(0, 1), cfg["vision"])
        feats = mlp_projector(sd, "model.mm_projector.", feats, cfg.get("projector_depth", 2))
        feats = feats.reshape(Bv, nv * feats.shape[1], feats.shape[2])
    else:
        feats = clip_vision_features(sd, "model.mm_vision_tower.", images, cfg["vision"])
        feats = mlp_projector(sd, "model.mm_projector.", feats, cfg.get("projector_depth", 2))
    emb, lab, msk, pid = splice(sd["model.llm.embed_tokens.weight"], feats, input_ids, attention_mask, labels,
                                cfg.get("tokenizer_model_max_length"), cfg.get("tokenizer_padding_side", "right"))
    hs = decoder_forward(sd, "model.llm.", emb, msk, pid, cfg["llm"])
    cog, idx = cognition_features(hs, msk)
    return dict(last_hidden=hs, cognition=cog, cognition_index=idx, inputs_embeds=emb, attention_mask=msk,
                position_ids=pid, labels=lab, image_features=feats)


# ----------------------------------------------------------------------------------------------
# OFT discrete action tokenizer (integer path) — oft/action_model/model.py:303-347,
# oft_discrete_arch.py:207-235, data/dataset/transform/action.py:378-390
# ----------------------------------------------------------------------------------------------
def oft_discretize(actions: np.ndarray, num_bins: int = 256) -> np.ndarray:
    """model.py:303-312: ((clamp(a,-1,1)+1)/2*(num_bins-1)).round().long(); float32, round-half-even."""
    a = np.clip(actions.astype(np.float32), np.float32(-1), np.float32(1))
    y = (a + np.float32(1)) / np.float32(2) * np.float32(num_bins - 1)
    return np.rint(y).astype(np.int64)


def oft_bins_to_continuous(bins: np.ndarray, num_bins: int = 256) -> np.ndarray:
    """model.py:314-347: (ids.float() / (num_bins-1)) * 2 - 1."""
    return (bins.astype(np.float32) / np.float32(num_bins - 1)) * np.float32(2) - np.float32(1)


def oft_argmax_decode(logits: np.ndarray, n_last: int = 255) -> np.ndarray:
    """oft_discrete_arch.py:222-224: argmax over the last 255 vocabulary entries, first maximum wins."""
    return np.argmax(logits[..., -n_last:], axis=-1).astype(np.int64)


def oft_sample_decode(logits: np.ndarray, n_last: int, temperature: float, u: np.ndarray) -> np.ndarray:
    """oft_discrete_arch.py:264-270 (generate_action): softmax(logits[..., -n_last:] / T), one multinomial draw per row.
    torch.multinomial's draw is not reproducible outside torch; the sampling DISTRIBUTION is what is specified, and a
    draw from it is the inverse CDF at a uniform u in [0, 1) — restated here in float64."""
    z = logits[..., -n_last:].astype(np.float64) / temperature
    p = np.exp(z - z.max(axis=-1, keepdims=True))
    cdf = np.cumsum(p, axis=-1)
    target = u.astype(np.float64)[..., None] * cdf[..., -1:]
    return np.minimum((cdf <= target).sum(axis=-1), n_last - 1).astype(np.int64)


def data_action_to_bin(action: np.ndarray, vocab_size: int = 255) -> np.ndarray:
    """data/dataset/transform/action.py:386-390 (_action2bin): np.round((a+1)/2*(V-1)) clipped to [0,V-1].
    NOTE the 254-vs-255 scale mismatch with the model side is the reference's; reproduced, not fixed."""
    a = np.round((action + 1) / 2 * (vocab_size - 1))
    return np.clip(a, 0, vocab_size - 1)


# ----------------------------------------------------------------------------------------------
# OFT-discrete training forward — oft_discrete_arch.py:26-205, oft_arch.py:169-210
# ----------------------------------------------------------------------------------------------
def oft_discrete_forward(sd, cfg: dict, input_ids, attention_mask, images, labels=None, states=None):
    """Returns dict(loss, logits [B, A, V], action_labels).  A = chunk_size * action_dim placeholder tokens
    (embed_tokens of token id 1) are inserted after the last valid token of every sample."""
    A = cfg["chunk_size"] * cfg["action_dim"]
    B = input_ids.shape[0]
    action_labels = None
    if labels is not None:                                            # :66-106
        ids2, mask2, al = [], [], []
        for i in range(B):
            npl = int(attention_mask[i].sum())
            prefix = npl - A - 1
            ids2.append(torch.cat([input_ids[i, :prefix], input_ids[i, npl - 1:]]))
            al.append(labels[i, prefix:prefix + A])
            m = torch.zeros(ids2[-1].shape[0], dtype=attention_mask.dtype)
            m[:prefix + 1] = 1
            mask2.append(m)
        input_ids, attention_mask, action_labels = torch.stack(ids2), torch.stack(mask2), torch.stack(al)
    feats = clip_vision_features(sd, "model.mm_vision_tower.", images, cfg["vision"])
    feats = mlp_projector(sd, "model.mm_projector.", feats, cfg.get("projector_depth", 2))
    emb, _, msk, _ = splice(sd["model.llm.embed_tokens.weight"], feats, input_ids, attention_mask, None,
                            cfg.get("tokenizer_model_max_length"), "right")
    S, D = emb.shape[1], emb.shape[2]
    act_emb = sd["model.llm.embed_tokens.weight"][torch.ones(A, dtype=torch.long)][None].expand(B, A, D)   # :125-130
    if cfg.get("use_proprio"):                                         # :132-137 state token in front
        h = "model.action_head.proprio_projector."
        st = F.linear(F.gelu(F.linear(states, sd[h + "fc1.weight"], sd[h + "fc1.bias"])), sd[h + "fc2.weight"],
                      sd[h + "fc2.bias"])
        act_emb = torch.cat([st.reshape(B, -1, D), act_emb], dim=1)
    n_act = act_emb.shape[1]
    lens = msk.long().sum(dim=1)
    emb2 = torch.zeros(B, S + n_act, D, dtype=emb.dtype)
    msk2 = torch.zeros(B, S + n_act, dtype=torch.bool)
    for i in range(B):                                                 # insert_action_embedding, oft_arch.py:169-201
        n = int(lens[i])
        emb2[i, :n] = emb[i, :n]
        emb2[i, n:n + n_act] = act_emb[i]
        emb2[i, n + n_act:] = emb[i, n:]
        msk2[i, :n + n_act] = True
    pid = torch.arange(S + n_act)[None, :].expand(B, S + n_act)        # position_ids=None -> HF arange
    hs = decoder_forward(sd, "model.llm.", emb2, msk2, pid, cfg["llm"])
    ah = torch.stack([hs[i, int(lens[i]):int(lens[i]) + n_act] for i in range(B)])       # :204-210
    if cfg.get("use_proprio"):
        ah = ah[:, 1:]                                                 # :161-162
    logits = F.linear(ah, sd["lm_head.weight"])
    loss = None
    if action_labels is not None:
        loss = F.cross_entropy(logits.reshape(-1, logits.shape[-1]), action_labels.reshape(-1), reduction="mean")
    return dict(loss=loss, logits=logits, action_labels=action_labels, action_hidden=ah)


def mlp_resnet(sd, p: str, x, num_blocks: int = 2):
    """MLPResNet (oft/action_model/model.py:104-130): LN -> fc1 -> ReLU -> n x [x + ReLU(Linear(LN(x)))] -> LN -> fc2."""
    ln = lambda t, n: F.layer_norm(t, (t.shape[-1],), sd[p + n + ".weight"], sd[p + n + ".bias"])  # noqa: E731
    x = F.relu(F.linear(ln(x, "layer_norm1"), sd[p + "fc1.weight"], sd[p + "fc1.bias"]))
    for i in range(num_blocks):
        q = f"{p}mlp_resnet_blocks.{i}.ffn."
        h = F.layer_norm(x, (x.shape[-1],), sd[q + "0.weight"], sd[q + "0.bias"])
        x = x + F.relu(F.linear(h, sd[q + "1.weight"], sd[q + "1.bias"]))
    return F.linear(ln(x, "layer_norm2"), sd[p + "fc2.weight"], sd[p + "fc2.bias"])


def oft_l1_forward(sd, cfg: dict, input_ids, attention_mask, images, actions=None, states=None):
    """OFTForCausalLM.forward with the `Linear` (L1 regression) head (oft_arch.py:58-166; head
    oft/action_model/model.py:133-165): the learned action_query rows (+ an optional proprio token in front) are
    inserted after the last valid token, the LLM runs causally over them, and the hidden states of the action rows go
    through the MLPResNet, chunk-wise ([B, chunk, action_dim * hidden]).  Returns dict(loss, predicted_actions)."""
    A, T = cfg["action_dim"], cfg["chunk_size"]
    B = input_ids.shape[0]
    feats = clip_vision_features(sd, "model.mm_vision_tower.", images, cfg["vision"])
    feats = mlp_projector(sd, "model.mm_projector.", feats, cfg.get("projector_depth", 2))
    emb, _, msk, _ = splice(sd["model.llm.embed_tokens.weight"], feats, input_ids, attention_mask, None,
                            cfg.get("tokenizer_model_max_length"), "right")
    h = "model.action_head."
    act = sd[h + "action_query"].expand(B, -1, -1)
    if cfg.get("use_proprio"):
        st = F.linear(F.gelu(F.linear(states, sd[h + "proprio_projector.fc1.weight"], sd[h + "proprio_projector.fc1.bias"])),
                      sd[h + "proprio_projector.fc2.weight"], sd[h + "proprio_projector.fc2.bias"])
        act = torch.cat([st.reshape(B, -1, emb.shape[-1]), act], dim=1)
    n_act = act.shape[1]
    S, D = emb.shape[1], emb.shape[2]
    lens = msk.long().sum(dim=1)
    emb2 = torch.zeros(B, S + n_act, D, dtype=emb.dtype)
    msk2 = torch.zeros(B, S + n_act, dtype=torch.bool)
    for i in range(B):                                                 # insert_action_embedding, oft_arch.py:169-201
        n = int(lens[i])
        emb2[i, :n] = emb[i, :n]
        emb2[i, n:n + n_act] = act[i]
        emb2[i, n + n_act:] = emb[i, n:]
        msk2[i, :n + n_act] = True
    pid = torch.arange(S + n_act)[None, :].expand(B, S + n_act)
    hs = decoder_forward(sd, "model.llm.", emb2, msk2, pid, cfg["llm"])
    ah = torch.stack([hs[i, int(lens[i]):int(lens[i]) + n_act] for i in range(B)])       # :204-210
    if cfg.get("use_proprio"):
        ah = ah[:, 1:]
    pred = mlp_resnet(sd, h + "model.", ah.reshape(B, T, -1))
    loss = None
    if actions is not None:
        a = actions.reshape(B, -1, A)[:, :T]
        loss = (a - pred).abs().mean()
    return dict(loss=loss, predicted_actions=pred, action_hidden=ah)


def sinusoidal_timestep_encoding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """SinusoidalPositionalEncoding.forward (oft/action_model/model.py:71-80)."""
    half = dim // 2
    w = torch.exp(torch.arange(half) * -math.log(10000) / (half - 1))
    e = t[:, None] * w[None, :]
    return torch.cat((e.sin(), e.cos()), dim=-1)


def oft_diffusion_forward(sd, cfg: dict, input_ids, attention_mask, images, noisy_dict: dict, actions=None, states=None):
    """OFTForCausalLM.forward with the `DiT` DiffusionActionHead (oft_arch.py:103-154; head model.py:197-271): the
    action rows are [proprio token |] timestep token | NoisyActionProjector(one scalar per token); the MLPResNet reads
    the hidden states of the noisy-action rows chunk-wise and predicts the noise; loss = MSE(noise_pred, noise).
    `noisy_dict` = dict(noise, noisy_actions, diffusion_timestep_embeddings) as sample_noisy_actions returns it."""
    A, T = cfg["action_dim"], cfg["chunk_size"]
    B = input_ids.shape[0]
    feats = clip_vision_features(sd, "model.mm_vision_tower.", images, cfg["vision"])
    feats = mlp_projector(sd, "model.mm_projector.", feats, cfg.get("projector_depth", 2))
    emb, _, msk, _ = splice(sd["model.llm.embed_tokens.weight"], feats, input_ids, attention_mask, None,
                            cfg.get("tokenizer_model_max_length"), "right")
    h = "model.action_head."
    na = noisy_dict["noisy_actions"].reshape(B, -1).unsqueeze(-1)                              # :113
    act = F.linear(F.gelu(F.linear(na, sd[h + "noisy_action_projector.fc1.weight"], sd[h + "noisy_action_projector.fc1.bias"])),
                   sd[h + "noisy_action_projector.fc2.weight"], sd[h + "noisy_action_projector.fc2.bias"])
    act = torch.cat([noisy_dict["diffusion_timestep_embeddings"].expand(B, 1, -1), act], dim=1)   # :115
    if cfg.get("use_proprio"):
        st = F.linear(F.gelu(F.linear(states, sd[h + "proprio_projector.fc1.weight"], sd[h + "proprio_projector.fc1.bias"])),
                      sd[h + "proprio_projector.fc2.weight"], sd[h + "proprio_projector.fc2.bias"])
        act = torch.cat([st.reshape(B, -1, emb.shape[-1]), act], dim=1)
    n_act = act.shape[1]
    S, D = emb.shape[1], emb.shape[2]
    lens = msk.long().sum(dim=1)
    emb2 = torch.zeros(B, S + n_act, D, dtype=emb.dtype)
    msk2 = torch.zeros(B, S + n_act, dtype=torch.bool)
    for i in range(B):                                                 # insert_action_embedding, oft_arch.py:169-201
        n = int(lens[i])
        emb2[i, :n] = emb[i, :n]
        emb2[i, n:n + n_act] = act[i]
        emb2[i, n + n_act:] = emb[i, n:]
        msk2[i, :n + n_act] = True
    pid = torch.arange(S + n_act)[None, :].expand(B, S + n_act)
    hs = decoder_forward(sd, "model.llm.", emb2, msk2, pid, cfg["llm"])
    ah = torch.stack([hs[i, int(lens[i]):int(lens[i]) + n_act] for i in range(B)])       # :204-210
    if cfg.get("use_proprio"):
        ah = ah[:, 1:]                                                                    # :139-140
    ah = ah[:, 1:]                                                                        # timestep row, :147
    pred = mlp_resnet(sd, h + "noise_predictor.mlp_resnet.", ah.reshape(B, T, -1))
    loss = None
    if actions is not None:
        loss = F.mse_loss(pred, noisy_dict["noise"], reduction="mean")
    return dict(loss=loss, predicted_noise=pred, action_hidden=ah)


def oft_diffusion_inference(sd, cfg: dict, input_ids, images, noise, num_ddim_steps: int = 10, states=None,
                            num_diffusion_steps: int = 100):
    """OFTForCausalLM.inference_action, `DiT` branch (oft_arch.py:224-250): DDIM loop, one full model call per step."""
    from oracle.ddim_oracle import DDIMSchedulerOracle
    sched = DDIMSchedulerOracle(num_train_timesteps=num_diffusion_steps, beta_schedule="squaredcos_cap_v2")
    sched.set_timesteps(num_ddim_steps)
    D = sd["model.llm.embed_tokens.weight"].shape[1]
    cur = noise
    mask = torch.ones_like(input_ids)
    for t in sched.timesteps:
        temb = sinusoidal_timestep_encoding(torch.Tensor([t]), D).unsqueeze(1)
        out = oft_diffusion_forward(sd, cfg, input_ids, mask, images,
                                    dict(noise=noise, noisy_actions=cur, diffusion_timestep_embeddings=temb), None, states)
        cur = sched.step(out["predicted_noise"], t, cur).prev_sample
    return cur


# ----------------------------------------------------------------------------------------------
# pi0 — dexbotic/model/pi0/pi0_arch.py (SigLIP tower: modules/mm_vision/siglip/siglip_encoder.py:61-86)
# ----------------------------------------------------------------------------------------------
def siglip_vision_features(sd, prefix: str, images: torch.Tensor, cfg: dict, select_layer=None) -> torch.Tensor:
    """HF SiglipVisionModel(...).last_hidden_state (select_layer=None, siglip_encoder.py:62-63): conv(+bias) patchify,
    learned position embedding, pre-LN encoder with gelu_tanh MLP, post_layernorm.  No CLS token.
    select_layer=-2 (the tower's default, siglip_encoder.py:13,64-65): hidden_states[-2] = the output of the
    second-to-last encoder layer, no post_layernorm."""
    p = prefix + "vision_tower.vision_model."
    patch, heads, eps = cfg["patch_size"], cfg["num_attention_heads"], cfg.get("layer_norm_eps", 1e-6)
    act = ACT[cfg.get("hidden_act", "gelu_pytorch_tanh")]
    x = F.conv2d(images, sd[p + "embeddings.patch_embedding.weight"], sd[p + "embeddings.patch_embedding.bias"],
                 stride=patch).flatten(2).transpose(1, 2)
    x = x + sd[p + "embeddings.position_embedding.weight"][None]
    D = x.shape[-1]
    n_layers = cfg["num_hidden_layers"] if select_layer is None else cfg["num_hidden_layers"] + 1 + select_layer
    for i in range(n_layers):
        q = f"{p}encoder.layers.{i}."
        h = F.layer_norm(x, (D,), sd[q + "layer_norm1.weight"], sd[q + "layer_norm1.bias"], eps)
        x = x + _mha(h, sd[q + "self_attn.q_proj.weight"], sd[q + "self_attn.q_proj.bias"],
                     sd[q + "self_attn.k_proj.weight"], sd[q + "self_attn.k_proj.bias"],
                     sd[q + "self_attn.v_proj.weight"], sd[q + "self_attn.v_proj.bias"],
                     sd[q + "self_attn.out_proj.weight"], sd[q + "self_attn.out_proj.bias"], heads)
        h = F.layer_norm(x, (D,), sd[q + "layer_norm2.weight"], sd[q + "layer_norm2.bias"], eps)
        x = x + F.linear(act(F.linear(h, sd[q + "mlp.fc1.weight"], sd[q + "mlp.fc1.bias"])),
                         sd[q + "mlp.fc2.weight"], sd[q + "mlp.fc2.bias"])
    if select_layer is not None:
        return x
    return F.layer_norm(x, (D,), sd[p + "post_layernorm.weight"], sd[p + "post_layernorm.bias"], eps)


def posemb_sincos(position: torch.Tensor, dim: int, min_period: float, max_period: float) -> torch.Tensor:
    """pi0_arch.py:36-50 (float64 periods; returns float64)."""
    fraction = torch.linspace(0.0, 1.0, dim // 2, dtype=torch.float64)
    period = min_period * (max_period / min_period) ** fraction
    s = position[:, None].float() / period[None, :] * 2 * np.pi
    return torch.cat([torch.sin(s), torch.cos(s)], dim=-1)


def pi0_make_attn_mask(input_mask: torch.Tensor, ar_mask: torch.Tensor) -> torch.Tensor:
    """pi0_arch.py:22-29: attend iff cumsum(ar)[k] <= cumsum(ar)[q], both positions valid."""
    cs = torch.cumsum(ar_mask.broadcast_to(input_mask.shape).long(), dim=1)
    allow = cs[:, None, :] <= cs[:, :, None]
    valid = input_mask[:, None, :] & input_mask[:, :, None]
    return allow & valid


def pi0_forward(sd, cfg: dict, input_ids, attention_mask, images, image_masks, actions, states, noise, time):
    """Pi0ForCausalLM.forward (pi0_arch.py:317-400) with injected noise [B,T,A] and time [B].
    Returns dict(loss, v_t, u_t, suffix_out, prefix_tokens, input_mask)."""
    te = time[:, None, None]
    x_t = te * noise + (1 - te) * actions
    u_t = noise - actions
    out = pi0_velocity(sd, cfg, input_ids, attention_mask, images, image_masks, states, x_t, time)
    out["u_t"] = u_t
    out["loss"] = ((out["v_t"] - u_t) ** 2).mean()
    return out


def pi0_inference(sd, cfg: dict, input_ids, attention_mask, images, image_masks, states, noise,
                  diffusion_steps: int = 10):
    """Pi0ForCausalLM.inference_action (pi0_arch.py:402-491): Euler integration of the flow from t=1 (noise) to t=0
    with dt = -1/steps.  The reference caches the prefix K/V (the prefix never attends to the suffix, :432-444) and
    runs only the suffix through the action expert per step; recomputing the joint forward each step — as done
    here — is the same arithmetic (same masks :452-458, same positions :465-469)."""
    B = states.shape[0]
    dt = -1.0 / diffusion_steps
    x = noise
    time = torch.tensor(1.0)
    while time > -dt / 2:
        v = pi0_velocity(sd, cfg, input_ids, attention_mask, images, image_masks, states, x,
                         time.broadcast_to(B))["v_t"]
        x = x + v * dt
        time = time + dt
    return x


def pi0_velocity(sd, cfg: dict, input_ids, attention_mask, images, image_masks, states, x_t, time):
    """embed_prefix + embed_suffix + _inner_forward_mot + action_out_proj (pi0_arch.py:116-315, 356-386)."""
    L, E, V = cfg["llm"], cfg["expert"], cfg["vision"]
    T = cfg["chunk_size"]
    B = x_t.shape[0]
    # embed_prefix (:235-269): cameras one by one, then text * sqrt(hidden)
    toks, masks = [], []
    for c in range(images.shape[1]):
        f = siglip_vision_features(sd, "model.mm_vision_tower.", images[:, c], V)
        f = F.linear(f, sd["model.mm_projector.weight"], sd["model.mm_projector.bias"])
        toks.append(f)
        masks.append(image_masks[:, c][:, None].expand(B, f.shape[1]))
    toks.append(sd["model.llm.embed_tokens.weight"][input_ids] * L["hidden_size"] ** 0.5)
    masks.append(attention_mask.bool())
    prefix = torch.cat(toks, dim=1)
    prefix_mask = torch.cat(masks, dim=1)
    # embed_suffix (:271-315)
    w = E["hidden_size"]
    state_tok = F.linear(states, sd["model.state_proj.weight"], sd["model.state_proj.bias"])[:, None]
    temb = posemb_sincos(time, w, 4e-3, 4.0)[:, None].expand(B, T, w)
    a_tok = F.linear(x_t, sd["model.action_in_proj.weight"], sd["model.action_in_proj.bias"])
    at = torch.cat([a_tok, temb.to(a_tok.dtype)], dim=-1)
    at = F.linear(F.silu(F.linear(at, sd["model.action_time_mlp_in.weight"], sd["model.action_time_mlp_in.bias"])),
                  sd["model.action_time_mlp_out.weight"], sd["model.action_time_mlp_out.bias"])
    suffix = torch.cat([state_tok, at], dim=1)
    Sp, Ss = prefix.shape[1], suffix.shape[1]
    input_mask = torch.cat([prefix_mask, torch.ones(B, Ss, dtype=torch.bool)], dim=1)
    ar = torch.tensor([False] * Sp + [True, True] + [False] * (T - 1))
    allow = pi0_make_attn_mask(input_mask, ar)[:, None]                      # [B,1,S,S]
    positions = torch.cumsum(input_mask.long(), dim=1) - 1
    H, KVH, hd = L["num_attention_heads"], L["num_key_value_heads"], L["head_dim"]
    cos, sin = rope_tables(positions, hd, L.get("rope_theta", 10000.0))
    c, s = cos[:, None], sin[:, None]
    xs = [prefix, suffix]
    prefixes = ["model.llm.", "model.action_expert."]
    cfgs = [L, E]
    for li in range(L["num_hidden_layers"]):                                  # _inner_forward_mot (:116-228)
        qs, ks, vs = [], [], []
        for x, pf, cc in zip(xs, prefixes, cfgs):
            q = f"{pf}layers.{li}."
            h = rms_norm(x, sd[q + "input_layernorm.weight"], cc["rms_norm_eps"], unit_offset=True)
            n = h.shape[1]
            qs.append(F.linear(h, sd[q + "self_attn.q_proj.weight"]).view(B, n, H, hd).transpose(1, 2))
            ks.append(F.linear(h, sd[q + "self_attn.k_proj.weight"]).view(B, n, KVH, hd).transpose(1, 2))
            vs.append(F.linear(h, sd[q + "self_attn.v_proj.weight"]).view(B, n, KVH, hd).transpose(1, 2))
        qh, kh, vh = torch.cat(qs, dim=2), torch.cat(ks, dim=2), torch.cat(vs, dim=2)
        qh = qh * c + rotate_half(qh) * s
        kh = kh * c + rotate_half(kh) * s
        G = H // KVH
        sc = (qh @ kh.repeat_interleave(G, dim=1).transpose(-1, -2)) * hd ** -0.5
        sc = sc + torch.where(allow, 0.0, -2.3819763e38)
        p = torch.softmax(sc.float(), dim=-1)
        o = (p @ vh.repeat_interleave(G, dim=1)).transpose(1, 2).reshape(B, Sp + Ss, H * hd)
        outs, start = [], 0
        for x, pf, cc in zip(xs, prefixes, cfgs):
            q = f"{pf}layers.{li}."
            n = x.shape[1]
            a = F.linear(o[:, start:start + n], sd[q + "self_attn.o_proj.weight"])
            start += n
            r = x + a
            h = rms_norm(r, sd[q + "post_attention_layernorm.weight"], cc["rms_norm_eps"], unit_offset=True)
            act = ACT[cc.get("hidden_act", "gelu_pytorch_tanh")]
            m = F.linear(act(F.linear(h, sd[q + "mlp.gate_proj.weight"])) * F.linear(h, sd[q + "mlp.up_proj.weight"]),
                         sd[q + "mlp.down_proj.weight"])
            outs.append(r + m)
        xs = outs
    suffix_out = rms_norm(xs[1], sd["model.action_expert.norm.weight"], E["rms_norm_eps"], unit_offset=True)
    v_t = F.linear(suffix_out[:, -T:], sd["model.action_out_proj.weight"], sd["model.action_out_proj.bias"])
    return dict(v_t=v_t, suffix_out=suffix_out, prefix_tokens=prefix, input_mask=input_mask)


# ----------------------------------------------------------------------------------------------
# pi0.5 — dexbotic/model/pi05/pi05_arch.py + pi05/transformers_pi05/gemma/modeling_gemma.py:38-119 (AdaRMS)
# ----------------------------------------------------------------------------------------------
def ada_rms_norm(x, sd, p: str, eps: float, cond=None):
    """GemmaRMSNorm.forward (modeling_gemma.py:38-88): plain `(1 + weight)` RMSNorm when the module has no `dense`
    (or no cond), else scale/shift/gate = chunk(dense(cond), 3): normed * (1 + scale) + shift, and the gate for the
    gated residual.  Returns (y, gate or None)."""
    n = x * torch.rsqrt(x.float().pow(2).mean(-1, keepdim=True) + eps)
    if cond is None or (p + "dense.weight") not in sd:
        return n * (1.0 + sd[p + "weight"].float()), None
    mod = F.linear(cond, sd[p + "dense.weight"], sd[p + "dense.bias"])[:, None, :]
    scale, shift, gate = torch.chunk(mod, 3, dim=-1)
    return n * (1 + scale.float()) + shift.float(), gate


def pi05_velocity(sd, cfg: dict, input_ids, attention_mask, images, image_masks, x_t, time):
    """embed_prefix + embed_suffix + _inner_forward_mot + action_out_proj (pi05_arch.py:117-420).  Differences from
    pi0: no state token; time enters only through adarms_cond = silu(time_mlp_out(silu(time_mlp_in(posemb(t)))))
    (:303-315), which modulates every norm of the action expert and gates its residuals (:146-152,217-228)."""
    L, E, V = cfg["llm"], cfg["expert"], cfg["vision"]
    T = cfg["chunk_size"]
    B = x_t.shape[0]
    toks, masks = [], []
    for c in range(images.shape[1]):
        f = siglip_vision_features(sd, "model.mm_vision_tower.", images[:, c], V)
        toks.append(F.linear(f, sd["model.mm_projector.weight"], sd["model.mm_projector.bias"]))
        masks.append(image_masks[:, c][:, None].expand(B, f.shape[1]))
    toks.append(sd["model.llm.embed_tokens.weight"][input_ids] * L["hidden_size"] ** 0.5)
    masks.append(attention_mask.bool())
    prefix, prefix_mask = torch.cat(toks, dim=1), torch.cat(masks, dim=1)
    w = E["hidden_size"]
    temb = posemb_sincos(time, w, 4e-3, 4.0).to(x_t.dtype)
    cond = F.silu(F.linear(F.silu(F.linear(temb, sd["model.time_mlp_in.weight"], sd["model.time_mlp_in.bias"])),
                           sd["model.time_mlp_out.weight"], sd["model.time_mlp_out.bias"]))
    suffix = F.linear(x_t, sd["model.action_in_proj.weight"], sd["model.action_in_proj.bias"])
    Sp, Ss = prefix.shape[1], suffix.shape[1]
    input_mask = torch.cat([prefix_mask, torch.ones(B, Ss, dtype=torch.bool)], dim=1)
    ar = torch.tensor([False] * Sp + [True] + [False] * (T - 1))
    allow = pi0_make_attn_mask(input_mask, ar)[:, None]
    positions = torch.cumsum(input_mask.long(), dim=1) - 1
    H, KVH, hd = L["num_attention_heads"], L["num_key_value_heads"], L["head_dim"]
    cos, sin = rope_tables(positions, hd, L.get("rope_theta", 10000.0))
    c, s = cos[:, None], sin[:, None]
    xs, prefixes, cfgs, conds = [prefix, suffix], ["model.llm.", "model.action_expert."], [L, E], [None, cond]
    for li in range(L["num_hidden_layers"]):
        qs, ks, vs, gates = [], [], [], []
        for x, pf, cc, cd in zip(xs, prefixes, cfgs, conds):
            q = f"{pf}layers.{li}."
            h, g = ada_rms_norm(x, sd, q + "input_layernorm.", cc["rms_norm_eps"], cd)
            n = h.shape[1]
            gates.append(g)
            qs.append(F.linear(h, sd[q + "self_attn.q_proj.weight"]).view(B, n, H, hd).transpose(1, 2))
            ks.append(F.linear(h, sd[q + "self_attn.k_proj.weight"]).view(B, n, KVH, hd).transpose(1, 2))
            vs.append(F.linear(h, sd[q + "self_attn.v_proj.weight"]).view(B, n, KVH, hd).transpose(1, 2))
        qh, kh, vh = torch.cat(qs, dim=2), torch.cat(ks, dim=2), torch.cat(vs, dim=2)
        qh = qh * c + rotate_half(qh) * s
        kh = kh * c + rotate_half(kh) * s
        G = H // KVH
        sc = (qh @ kh.repeat_interleave(G, dim=1).transpose(-1, -2)) * hd ** -0.5
        sc = sc + torch.where(allow, 0.0, -2.3819763e38)
        p = torch.softmax(sc.float(), dim=-1)
        o = (p @ vh.repeat_interleave(G, dim=1)).transpose(1, 2).reshape(B, Sp + Ss, H * hd)
        outs, start = [], 0
        for x, pf, cc, cd, g1 in zip(xs, prefixes, cfgs, conds, gates):
            q = f"{pf}layers.{li}."
            n = x.shape[1]
            a = F.linear(o[:, start:start + n], sd[q + "self_attn.o_proj.weight"])
            start += n
            r = x + a if g1 is None else x + a * g1
            h, g2 = ada_rms_norm(r, sd, q + "post_attention_layernorm.", cc["rms_norm_eps"], cd)
            act = ACT[cc.get("hidden_act", "gelu_pytorch_tanh")]
            m = F.linear(act(F.linear(h, sd[q + "mlp.gate_proj.weight"])) * F.linear(h, sd[q + "mlp.up_proj.weight"]),
                         sd[q + "mlp.down_proj.weight"])
            outs.append(r + m if g2 is None else r + m * g2)
        xs = outs
    suffix_out, _ = ada_rms_norm(xs[1], sd, "model.action_expert.norm.", E["rms_norm_eps"], cond)
    v_t = F.linear(suffix_out[:, -T:], sd["model.action_out_proj.weight"], sd["model.action_out_proj.bias"])
    return dict(v_t=v_t, suffix_out=suffix_out, prefix_tokens=prefix, input_mask=input_mask, adarms_cond=cond)


def pi05_forward(sd, cfg: dict, input_ids, attention_mask, images, image_masks, actions, noise, time):
    """Pi05ForCausalLM.forward (pi05_arch.py:333-420) with injected noise / time."""
    te = time[:, None, None]
    x_t = te * noise + (1 - te) * actions
    u_t = noise - actions
    out = pi05_velocity(sd, cfg, input_ids, attention_mask, images, image_masks, x_t, time)
    out["u_t"] = u_t
    out["loss"] = ((out["v_t"] - u_t) ** 2).mean()
    return out


def pi05_inference(sd, cfg: dict, input_ids, attention_mask, images, image_masks, noise, diffusion_steps: int = 10):
    """Pi05ForCausalLM.inference_action (pi05_arch.py:424-514), cache-free restatement (see pi0_inference)."""
    B = noise.shape[0]
    dt = -1.0 / diffusion_steps
    x, time = noise, torch.tensor(1.0)
    while time > -dt / 2:
        v = pi05_velocity(sd, cfg, input_ids, attention_mask, images, image_masks, x, time.broadcast_to(B))["v_t"]
        x = x + v * dt
        time = time + dt
    return x


# ----------------------------------------------------------------------------------------------
# CogACT inference — cogact_arch.py:149-198, diffusion.py:626-673 (ddim_sample), :990-1112 (respacing)
# ----------------------------------------------------------------------------------------------
def ddim_tables(num_steps: int = 100, ddim_steps: int = 10):
    """Spaced-diffusion constants: timestep_map and float64 alphas_cumprod / alphas_cumprod_prev of the respaced
    process (space_timesteps 'ddimN' + SpacedDiffusion.__init__)."""
    sa, _ = cosine_schedule(num_steps)
    ac = sa ** 2
    stride = next(i for i in range(1, num_steps) if len(range(0, num_steps, i)) == ddim_steps)
    tmap = list(range(0, num_steps, stride))
    last, betas = 1.0, []
    for i in tmap:
        betas.append(1 - ac[i] / last)
        last = ac[i]
    ac2 = np.cumprod(1.0 - np.array(betas, dtype=np.float64))
    return tmap, ac2, np.append(1.0, ac2[:-1])


def ddim_sample(sd, cfg: dict, cog, noise, cfg_scale: float = 1.5, num_ddim_steps: int = 10, per_token=None):
    """ddim_sample_loop (diffusion.py:714-795) with eta = 0, clip_denoised=False, classifier-free guidance through
    forward_with_cfg (dit.py:294-311).  per_token (MemVLA): repeated twice along the batch like the reference does
    (memvla_arch.py:722)."""
    B = cog.shape[0]
    width = sd["model.action_head.net.x_embedder.linear.weight"].shape[0]
    heads = DIT_HEADS[width]
    use_cfg = cfg_scale > 1.0
    x = noise
    if use_cfg:
        x = torch.cat([noise, noise], 0)
        unc = sd["model.action_head.net.z_embedder.uncondition"][None].expand(B, 1, -1)
        z = torch.cat([cog, unc], 0)
    else:
        z = cog
    pt = None if per_token is None else per_token.repeat(2, 1, 1)
    tmap, ac, ac_prev = ddim_tables(cfg.get("diffusion_steps", 100), num_ddim_steps)
    f32 = lambda v: torch.tensor(v, dtype=torch.float64).float()          # noqa: E731 (_extract_into_tensor)
    for i in reversed(range(len(tmap))):
        t = torch.full((x.shape[0],), tmap[i], dtype=torch.long)
        if use_cfg:                                                          # forward_with_cfg, dit.py:294-311
            half = x[: x.shape[0] // 2]
            out = dit_forward(sd, "model.action_head.", torch.cat([half, half], 0), t, z, None, heads, per_token=pt)
            cond, uncond = out.chunk(2, dim=0)
            e = uncond + cfg_scale * (cond - uncond)
            eps_model = torch.cat([e, e], 0)
        else:
            eps_model = dit_forward(sd, "model.action_head.", x, t, z, None, heads, per_token=pt)
        sr, srm1 = f32(np.sqrt(1.0 / ac[i])), f32(np.sqrt(1.0 / ac[i] - 1))
        pred_x0 = sr * x - srm1 * eps_model                                  # _predict_xstart_from_eps
        eps = (sr * x - pred_x0) / srm1                                      # _predict_eps_from_xstart
        x = pred_x0 * torch.sqrt(f32(ac_prev[i])) + torch.sqrt(1 - f32(ac_prev[i])) * eps   # eta = 0
    return x[:B] if use_cfg else x


def _inference_trunk(sd, cfg: dict, input_ids, images):
    feats = clip_vision_features(sd, "model.mm_vision_tower.", images, cfg["vision"])
    feats = mlp_projector(sd, "model.mm_projector.", feats, cfg.get("projector_depth", 2))
    emb, _, msk, pid = splice(sd["model.llm.embed_tokens.weight"], feats, input_ids, None, None,
                              cfg.get("tokenizer_model_max_length"), "right")
    hs = decoder_forward(sd, "model.llm.", emb, msk, pid, cfg["llm"])
    return hs[:, -1, :][:, None], feats                                     # cognition = last position (:158)


def cogact_inference(sd, cfg: dict, input_ids, images, noise, cfg_scale: float = 1.5, num_ddim_steps: int = 10):
    """CogACTForCausalLM.inference_action up to (excluding) _denorm: returns samples [B, T, A] (normalised actions)."""
    cog, _ = _inference_trunk(sd, cfg, input_ids, images)
    return ddim_sample(sd, cfg, cog, noise, cfg_scale, num_ddim_steps)


def memvla_inference(sd, cfg: dict, banks: dict, input_ids, images, noise, timestep: int, cfg_scale: float = 1.5,
                     num_ddim_steps: int = 10):
    """One MemVLAForCausalLM.inference_action call (memvla_arch.py:666-745), excluding _denorm: the cognition token
    and the compressed perceptual tokens of this frame go through the memory bank in eval mode (every frame of the
    running episode has id (0, 0), :334-337), which also appends them to it; then CFG + DDIM with per_token.
    `banks` = {"per": MemBankOracle, "cog": MemBankOracle} carried across the frames of an episode."""
    cog, feats = _inference_trunk(sd, cfg, input_ids, images)
    per = bottleneck_se(sd, "model.per_compr.", feats)
    ts = [torch.tensor(timestep)]
    cog = banks["cog"].process_batch(cog, [(0, 0)], ts, training=False)
    per = banks["per"].process_batch(per, [(0, 0)], ts, training=False)
    return ddim_sample(sd, cfg, cog, noise, cfg_scale, num_ddim_steps, per_token=per)


# ----------------------------------------------------------------------------------------------
# NaVILA — dexbotic/model/navila/navila_arch.py (VLM + (soft) cross entropy), navila/loss.py,
# mm_projector/builder.py:9-33,61-68 (mlp_downsample)
# ----------------------------------------------------------------------------------------------
def downsample_2x2(x: torch.Tensor) -> torch.Tensor:
    """DownSampleBlock (mm_projector/builder.py:9-33): [N, h*w, C] -> [N, ceil(h/2)*ceil(w/2), 4C].  The token grid
    (row-major, rows = first axis) is zero-padded to even sides; output token t = j * ceil(h/2) + i concatenates the
    cells (2i, 2j), (2i, 2j+1), (2i+1, 2j), (2i+1, 2j+1) in that order — the two view/permute steps of flat_square
    written as one gather (checked equal to the reference module in oracle/make_golden.py:make_navila_tiny)."""
    N, T, C = x.shape
    h = w = int(T ** 0.5)
    g = x.reshape(N, h, w, C)
    g = F.pad(g, (0, 0, 0, w % 2, 0, h % 2))
    H2, W2 = g.shape[1] // 2, g.shape[2] // 2
    g = g.reshape(N, H2, 2, W2, 2, C)                     # [N, i, di, j, dj, C]
    out = g.permute(0, 3, 1, 2, 4, 5)                     # [N, j, i, di, dj, C]
    return out.reshape(N, W2 * H2, 4 * C)


def navila_projector(sd, prefix: str, feats: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """mlp_downsample: DownSampleBlock, LayerNorm(4C), Linear(4C, d), GELU(erf), Linear(d, d)."""
    x = downsample_2x2(feats)
    x = F.layer_norm(x, (x.shape[-1],), sd[prefix + "1.weight"], sd[prefix + "1.bias"], eps)
    x = F.gelu(F.linear(x, sd[prefix + "2.weight"], sd[prefix + "2.bias"]))
    return F.linear(x, sd[prefix + "4.weight"], sd[prefix + "4.bias"])


def soft_cross_entropy(logits: torch.Tensor, targets: torch.Tensor, soft_tokens, std: float = 1.0) -> torch.Tensor:
    """navila/loss.py:11-70: shifted CE where a target that is one of `soft_tokens` (the time tokens) is replaced by a
    Gaussian over the soft-token IDs, exp(-(target_id - soft_id)^2 / (2 std^2)) normalised; sum / number of valid
    targets."""
    out = logits[..., :-1, :].reshape(-1, logits.shape[-1])
    tgt = targets[..., 1:].reshape(-1)
    keep = tgt != IGNORE_INDEX
    out, tgt = out[keep], tgt[keep]
    if out.numel() == 0:
        return torch.zeros((), dtype=logits.dtype)
    soft = torch.as_tensor(soft_tokens, dtype=tgt.dtype)
    is_soft = torch.isin(tgt, soft)
    lsm = F.log_softmax(out.float(), dim=-1)
    loss = -lsm[~is_soft].gather(1, tgt[~is_soft][:, None]).sum()
    if is_soft.any():
        d = torch.exp(-((tgt[is_soft][:, None] - soft[None, :]) ** 2).float() / (2 * std ** 2))
        d = d / d.sum(dim=1, keepdim=True)
        loss = loss - (lsm[is_soft][:, soft] * d).sum()
    return loss / tgt.shape[0]


def navila_forward(sd, cfg: dict, input_ids, attention_mask, images, labels, time_token_ids=None, soft_ce_std=1.0):
    """NaVILAForCausalLM.forward (navila_arch.py:362-497), training path without sequence packing (HF decoders do not
    take seqlens_in_batch, :415-417): SigLIP (select_layer -2) -> mlp_downsample -> splice (every row carries the same
    number of <image> tokens, each consuming its share of the row's features, :166-205) -> decoder -> lm_head ->
    HF ForCausalLMLoss (shifted, mean over labels != -100) or soft_cross_entropy when time_token_ids is set (:474-489).
    images [B, 3, H, W] or [B, n, 3, H, W].  Returns dict(loss, logits, labels, attention_mask)."""
    B = input_ids.shape[0]
    imgs = images.reshape(-1, *images.shape[-3:])
    feats = siglip_vision_features(sd, "model.mm_vision_tower.", imgs, cfg["vision"], select_layer=-2)
    feats = navila_projector(sd, "model.mm_projector.", feats)                      # [B*n, P', d]
    emb, lab, msk, pid = splice(sd["model.llm.embed_tokens.weight"], feats, input_ids, attention_mask, labels,
                                cfg.get("tokenizer_model_max_length"), cfg.get("tokenizer_padding_side", "right"))
    hidden = decoder_forward(sd, "model.llm.", emb, msk, pid, cfg["llm"])
    logits = F.linear(hidden, sd["lm_head.weight"])
    if time_token_ids:
        loss = soft_cross_entropy(logits, lab, time_token_ids, soft_ce_std)
    else:
        sl = logits[:, :-1].reshape(-1, logits.shape[-1]).float()
        st = lab[:, 1:].reshape(-1)
        loss = F.cross_entropy(sl, st, ignore_index=IGNORE_INDEX, reduction="mean")
    return dict(loss=loss, logits=logits, labels=lab, attention_mask=msk)
