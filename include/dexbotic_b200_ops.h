/*
 * dexbotic_b200 — C-ABI, part 2: the HBM-bound operators of the VLA hot path.
 * Same conventions as dexbotic_b200.h (device pointers, stream as void*, 0 = ok).
 * `dtype` is B200_BF16 or B200_F32 and names the activation/storage type; all arithmetic
 * accumulates in fp32.  Citations are the reference Python call sites (under
 * /root/reference) whose arithmetic each entry point replaces; where the arithmetic lives in
 * an un-vendored third-party module (HF transformers 4.51/4.54, timm — SURVEY.md §8c) the
 * reference line is the call into it.
 */
#ifndef DEXBOTIC_B200_OPS_H
#define DEXBOTIC_B200_OPS_H

#include <stdint.h>

#include "dexbotic_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* RMSNorm  y = w * round(x * rsqrt(mean(x^2)+eps))            (HF Qwen2RMSNorm / LlamaRMSNorm,
 * unit_offset=1: y = (x*rstd) * (1+w)  (HF GemmaRMSNorm);  decoder built at dexbotic_arch.py:55-62,
 * final norm read at cogact_arch.py:108.  rstd[M] (fp32) is saved for backward (may be NULL). */
int b200_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd, int64_t M, int64_t D, float eps,
                     int unit_offset, int dtype, void* stream);
/* dx (=, or += if accumulate_dx) and dw[D] (fp32, +=; may be NULL).  workspace: optional fp32 scratch of
 * b200_norm_bwd_workspace_rows(M, D) x D floats (x 2D for layernorm) for contention-free partial sums; NULL falls
 * back to fp32 atomics. */
int64_t b200_norm_bwd_workspace_rows(int64_t M, int64_t D);
/* Bit mask of the RMSNorm kernels that stage their rows in shared memory through the bulk-copy engine (cp.async.bulk +
 * mbarrier ring) instead of register prefetch: bit 0 = backward (default on), bit 1 = forward (default off: measured
 * slower).  Bit 2 (default off, opt-in): LayerNorm forward / backward with one warp per row and a column kernel for
 * dw / db (D <= 1280) instead of one block per row.  For A/B measurements and tests. */
int b200_set_norm_staged(int mask);
int b200_rmsnorm_bwd(const void* dy, const void* x, const void* w, const float* rstd, void* dx, float* dw,
                     float* workspace, int64_t M, int64_t D, int unit_offset, int accumulate_dx, int dtype,
                     void* stream);

/* LayerNorm (w, b optional): HF CLIPEncoderLayer.layer_norm1/2, pre_layrnorm (clip_encoder.py:50-54);
 * DiT norm1/norm2/norm_final, elementwise_affine=False (dit.py:141,147,167). */
int b200_layernorm_fwd(const void* x, const void* w, const void* b, void* y, float* mean, float* rstd, int64_t M,
                       int64_t D, float eps, int dtype, void* stream);
int b200_layernorm_bwd(const void* dy, const void* x, const void* w, const float* mean, const float* rstd, void* dx,
                       float* dw, float* db, float* workspace, int64_t M, int64_t D, int accumulate_dx, int dtype,
                       void* stream);

/* rotate_half RoPE, in place, on the first n_rot_heads heads (q then k) of each row of a packed
 * [M, row_stride] qkv buffer.  cos/sin: fp32 [n_pos, head_dim/2] tables, pos: int32 [M].
 * inverse=1 applies the transpose (backward).  HF apply_rotary_pos_emb inside Qwen2/Llama
 * attention; position_ids come from dexbotic_arch.py:368-371. */
int b200_rope(void* qkv, const int32_t* pos, const float* cos_t, const float* sin_t, int64_t M, int n_rot_heads,
              int head_dim, int64_t row_stride, int inverse, int dtype, void* stream);

/* y = act(x) ; dx = dy * act'(x).  act codes: B200_ACT_*. */
int b200_act_fwd(const void* x, void* y, int64_t n, int act, int dtype, void* stream);
int b200_act_bwd(const void* dy, const void* x, void* dx, int64_t n, int act, int dtype, void* stream);
/* h = act(g) * u  (HF Qwen2MLP: down(silu(gate(x)) * up(x)); GemmaMLP with gelu_tanh) */
int b200_glu_fwd(const void* g, const void* u, void* h, int64_t n, int act, int dtype, void* stream);
/* dg = dh*u*act'(g), du = dh*act(g) (may alias g/u); h_out optional recompute of h. */
int b200_glu_bwd(const void* dh, const void* g, const void* u, void* dg, void* du, void* h_out, int64_t n, int act,
                 int dtype, void* stream);

/* Masked softmax over fp32 score rows [Z, Sq, s_ld] -> P [Z, Sq, p_ld] (p_dtype).  Batch b = z / heads.
 * allowed(q,k) = (keymask==NULL || keymask[b,k]) && (bid_q==NULL || bid_k[b,k] <= bid_q[b,q]):
 *   causal decoder: causal=1 (k <= q by index; bid_* ignored and keys beyond q never read) — HF SDPA causal +
 *   padding mask;
 *   pi0 block-causal: bid = cumsum(ar_mask) (pi0_arch.py:22-33);  ViT / DiT: no mask. */
int b200_softmax_fwd(const float* scores, void* p, int64_t Z, int64_t Sq, int64_t Sk, int64_t s_ld, int64_t p_ld,
                     int heads, const uint8_t* keymask, const int32_t* bid_q, const int32_t* bid_k, int causal,
                     int p_dtype, void* stream);
/* dS = scale * P * (dP - rowsum(P*dP)) */
int b200_softmax_bwd(const void* p, const float* dp, void* ds, int64_t rows, int64_t Sk, int64_t p_ld, int64_t dp_ld,
                     int64_t ds_ld, float scale, int p_dtype, void* stream);

/* Fused attention scores (bf16, head_dim <= 128, Sk <= 512): the [128 x Sk] score block of a (batch, head, query
 * tile) stays in TMEM.  mode 0: P = masked softmax(scale * A B^T) with A = Q, B = K;  mode 1: dS = scale * P *
 * (A B^T - rowsum(P * A B^T)) with A = dO, B = V.  A: (hd, Sq, H, B) / B: (hd, Sk, KVH, B) strided views (element
 * strides a_ld / a_s_head / a_s_batch ...), out / p_in: [B*H, Sq, p_ld] bf16.  Mask rule as b200_softmax_fwd.
 * Replaces QK^T + softmax (and dP + softmax backward) of HF attention without an fp32 score round trip. */
int b200_attn_scores(const void* a, const void* b, const void* p_in, void* out, int64_t B, int64_t H, int64_t KVH,
                     int64_t Sq, int64_t Sk, int64_t head_dim, int64_t a_ld, int64_t a_s_head, int64_t a_s_batch,
                     int64_t b_ld, int64_t b_s_head, int64_t b_s_batch, int64_t p_ld, float scale, int causal,
                     const uint8_t* keymask, const int32_t* bid_q, const int32_t* bid_k, int mode, void* stream);

/* Flash attention (bf16, head_dim % 8 == 0 and <= 256, self-attention): O = softmax_mask(scale * Q K^T) V with the score
 * blocks resident in TMEM — no [B, H, S, S] tensor is written; lse[B, H, S] (fp32, log2 domain: scale*log2(e)*rowmax +
 * log2(rowsum)) is all the backward needs besides O.  q / k / v: (head_dim, S, heads, B) strided views of a packed qkv
 * buffer (element strides qkv_ld per row, qkv_s_head per head, qkv_s_batch per batch; H heads for q, KVH for k / v);
 * out: same form with its own strides.  Mask rule as b200_softmax_fwd (causal / key padding / pi0 block ids).
 * mask_ws: int32 workspace [B, ceil(S/64), 8] (16-byte aligned) for the per-block mask summary the kernels read;
 * required when keymask or block ids are given, may be NULL otherwise.
 * Replaces F.scaled_dot_product_attention (HF Qwen2 / Llama / CLIP / SigLIP attention, dexbotic_arch.py:55-62,
 * clip_encoder.py:50-54, siglip_encoder.py:79-84) and pi0's eager joint attention (pi0_arch.py:22-33,185-192). */
int b200_flash_attn_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int64_t B, int64_t H,
                        int64_t KVH, int64_t S, int64_t head_dim, int64_t qkv_ld, int64_t qkv_s_head,
                        int64_t qkv_s_batch, int64_t o_ld, int64_t o_s_head, int64_t o_s_batch, float scale, int causal,
                        const uint8_t* keymask, const int32_t* bid_q, const int32_t* bid_k, int32_t* mask_ws, void* stream);
/* Debugging hook (tools/trace_flash.py): when buf != NULL, CTA 0 of the flash kernels writes clock64() stamps of its
 * pipeline events into buf (int64 [64 slots][64 steps]); NULL (the default) switches it off. */
int b200_flash_attn_set_trace(void* buf);

/* Its backward: recomputes P from lse; dq / dk / dv use the strides (g_ld, g_s_head, g_s_batch) (the packed dqkv
 * buffer), dout the strides of out.  dK / dV are reduced over the query heads of a GQA group inside the kernel's
 * TMEM accumulators (no atomics: deterministic).  delta: fp32 workspace [B, H, S]. */
int b200_flash_attn_bwd(const void* q, const void* k, const void* v, const void* out, const void* dout,
                        const float* lse, float* delta, void* dq, void* dk, void* dv, int64_t B, int64_t H, int64_t KVH,
                        int64_t S, int64_t head_dim, int64_t qkv_ld, int64_t qkv_s_head, int64_t qkv_s_batch,
                        int64_t o_ld, int64_t o_s_head, int64_t o_s_batch, int64_t g_ld, int64_t g_s_head,
                        int64_t g_s_batch, float scale, int causal, const uint8_t* keymask, const int32_t* bid_q,
                        const int32_t* bid_k, int32_t* mask_ws, void* stream);

/* out[N] (fp32) += column sums of x[M,N]  (bias gradients) */
int b200_colsum(const void* x, float* out, int64_t M, int64_t N, int dtype, void* stream);
/* *out (fp32) += sum(x^2)  (global grad-norm, trainer.py:122 max_grad_norm=1.0) */
int b200_sumsq(const void* x, int64_t n, float* out, int dtype, void* stream);
int b200_clip_coef(const float* sumsq, float max_norm, float* clip, float* norm_out, void* stream);

/* *out += mean((a-b)^2)   (action_models.py:119-121; pi0_arch.py:388) ; da = 2(a-b)/n * (*gscale) */
int b200_mse_fwd(const void* a, const void* b, int64_t n, float* out, int dtype, void* stream);
int b200_mse_bwd(const void* a, const void* b, int64_t n, const float* gscale, void* da, int dtype, void* stream);

/* torch.optim.AdamW step on a flat fp32 master buffer (trainer.py:25-36 create_optimizer); gradient
 * in g (g_dtype), optional bf16 shadow of the updated weights, device-side clip coefficient.  Hyper-parameters are
 * doubles: torch evaluates 1-beta, the bias corrections and lr/bc1 in Python doubles and rounds to fp32 once. */
int b200_adamw(float* p, const void* g, float* m, float* v, void* shadow_bf16, int64_t n, double lr, double beta1,
               double beta2, double eps, double weight_decay, int64_t step, const float* clip, int g_dtype, void* stream);

/* The same update with its seven scalars read from device memory (hyper8: b2, 1-b1, 1-b2, eps, 1-lr*wd, lr/bc1,
 * sqrt(bc2), pad): a CUDA graph that contains the optimizer is replayed with new learning rates and step counts by
 * rewriting that block.  b200_adamw_hyper fills a host block with exactly the values b200_adamw would use (step <= 0:
 * the identity update). */
int b200_adamw_hyper(double lr, double beta1, double beta2, double eps, double weight_decay, int64_t step, float* out8);
int b200_adamw_dev(float* p, const void* g, float* m, float* v, void* shadow_bf16, int64_t n, const float* hyper8,
                   const float* clip, int g_dtype, void* stream);

int b200_cast(const void* src, void* dst, int64_t n, int src_dtype, int dst_dtype, void* stream);
int b200_add(const void* a, const void* b, void* y, int64_t n, int dtype, void* stream);

/* ---- image-token splice (dexbotic_arch.py:182-373) -------------------------------------------------
 * Plan: for every sample, drop padding (attention_mask), replace each IMAGE_TOKEN_INDEX (-200) by the
 * next n_img_tokens rows of the image features, truncate to max_len, pad to S (right or left).
 * Outputs: src[B*S] (>=0: token id; -1-j: image-feature row j; INT32_MIN: pad), new_labels[B*S]
 * (int64, -100 on image span / pad), new_mask[B*S] (uint8), pos[B*S] (int32), lengths[B] (int32).
 * S must be >= max length (b200_splice_lengths computes lengths first; host reads the max). */
int b200_splice_lengths(const int64_t* input_ids, const uint8_t* attention_mask, int64_t B, int64_t L,
                        int n_img_tokens, int64_t max_len, int32_t* lengths, void* stream);
int b200_splice_plan(const int64_t* input_ids, const uint8_t* attention_mask, const int64_t* labels, int64_t B,
                     int64_t L, int n_img_tokens, int64_t max_len, int64_t S, int left_pad, int32_t* src,
                     int64_t* new_labels, uint8_t* new_mask, int32_t* pos, void* stream);
/* out[r,:] = src[r]>=0 ? table[src[r]] : (src[r]==INT32_MIN ? 0 : feats[-1-src[r]])   rows of D elements */
int b200_splice_gather(const int32_t* src, const void* table, const void* feats, void* out, int64_t rows, int64_t D,
                       int dtype, void* stream);
/* backward: d_table[id] += sum of dout rows whose src == id (fp32 sum, one deterministic add per id);
 * d_feats[-1-src[r]] = dout[r].  Either output may be NULL (frozen embedding / frozen tower). */
int b200_splice_scatter(const int32_t* src, const void* dout, void* d_table, void* d_feats, int64_t rows, int64_t D,
                        int dtype, void* stream);

/* out[i,:] = x[idx[i],:] ; backward dx[idx[i],:] += dout[i,:]   (cognition token: cogact_arch.py:110-120;
 * OFT action-position extract: oft_arch.py:204-210) */
int b200_gather_rows(const void* x, const int32_t* idx, void* out, int64_t n_idx, int64_t D, int dtype, void* stream);
int b200_scatter_rows_add(const void* dout, const int32_t* idx, void* dx, int64_t n_idx, int64_t D, int dtype,
                          void* stream);
/* idx[b] = b*S + (index of last position with mask[b,:]!=0)   (cogact_arch.py:112-117) */
int b200_last_valid_index(const uint8_t* mask, int64_t B, int64_t S, int32_t* idx, void* stream);

/* x_t = sqrt_ac[t]*x + sqrt_1mac[t]*noise   (diffusion.py:308-326 q_sample; tables from diffusion.py:187-231) */
int b200_q_sample(const void* x, const void* noise, const int32_t* t, const float* sqrt_ac, const float* sqrt_1mac,
                  void* x_t, int64_t B, int64_t per_sample, int dtype, void* stream);
/* sinusoidal timestep embedding [B, dim] = [cos(t*f) | sin(t*f)]  (dit.py:37-56) */
int b200_timestep_embedding(const float* t, void* out, int64_t B, int dim, float max_period, int dtype, void* stream);

/* ---- OFT discrete action tokenizer (integer path; bit-exact) ---------------------------------------
 * bins[i] = (int64) round_half_even((clamp(a,-1,1)+1)/2 * (n_bins-1))          oft/action_model/model.py:303-312
 * cont[i] = bins[i] / (n_bins-1) * 2 - 1                                        model.py:314-323
 * idx[r]  = argmax_j logits[r, V-n_last+j], first maximum wins                 oft_discrete_arch.py:222-224 */
int b200_discretize_actions(const float* actions, int64_t n, int n_bins, int64_t* bins, void* stream);
int b200_bins_to_continuous(const int64_t* bins, int64_t n, int n_bins, float* out, void* stream);
int b200_argmax_last(const void* logits, int64_t rows, int64_t V, int n_last, int64_t* idx, int dtype, void* stream);
/* One temperature-scaled draw per row from softmax(logits[row, V - n_last :] / temperature)
 * (OFTDiscreteForCausalLM.generate_action, oft_discrete_arch.py:264-270: softmax + torch.multinomial(.., 1)).
 * u[rows]: uniforms in [0, 1) supplied by the caller (torch's generator on the host side); idx = inverse CDF at u. */
int b200_sample_last(const void* logits, int64_t rows, int64_t V, int n_last, float temperature, const float* u,
                     int64_t* idx, int dtype, void* stream);
/* Cross entropy over fp32-upcast logits rows with int64 labels (ignore_index=-100), mean over valid rows:
 * *loss += sum(-log softmax[label]) / n_valid ; lse[rows] saved.  Backward: dlogits = (softmax - onehot) * g / n_valid.
 * oft_discrete_arch.py:187-191 (F.cross_entropy under fp32 autocast); DexboticForCausalLM loss (dexbotic_arch.py:489). */
int b200_cross_entropy_fwd(const void* logits, const int64_t* labels, int64_t rows, int64_t V, float* lse,
                           float* loss_sum, int32_t* n_valid, int dtype, void* stream);
int b200_cross_entropy_bwd(const void* logits, const int64_t* labels, const float* lse, const int32_t* n_valid,
                           const float* gscale, void* dlogits, int64_t rows, int64_t V, int dtype, void* stream);

/* ---- ViT front end (HF CLIPVisionEmbeddings, called from clip_encoder.py:50-54) ----------------------
 * Patchify: the stride=patch Conv2d is a GEMM over im2col rows; out[B*P, Kpad] with columns
 * c*ps*ps + i*ps + j (the conv weight's own flattening) and zero padding up to Kpad (16-byte rows). */
int b200_im2col_patches(const void* images, void* out, int64_t B, int C, int H, int W, int patch, int Kpad,
                        int in_dtype, int out_dtype, void* stream);
/* out[b,0] = class_embedding + pos[0]; out[b,1+p] = patches[b,p] + pos[1+p] */
int b200_vit_embed_fwd(const void* patches, const void* cls, const void* pos, void* out, int64_t B, int64_t P,
                       int64_t D, int dtype, void* stream);
int b200_vit_embed_bwd(const void* dout, void* d_patches, float* d_cls, float* d_pos, int64_t B, int64_t P, int64_t D,
                       int dtype, void* stream);
/* dst (+)= src (fp32 scratch -> parameter gradient) ; 2-D strided copy/accumulate with dtype conversion */
int b200_cast_add(const float* src, void* dst, int64_t n, int dst_dtype, int accumulate, void* stream);
int b200_copy2d(const void* src, void* dst, int64_t rows, int64_t cols, int64_t lds, int64_t ldd, int src_dtype,
                int dst_dtype, int accumulate, void* stream);
/* dst[b, r, :cols] (+)= alpha * src[b, r, :cols] with independent batch / row strides (elements): moves row blocks
 * between per-stream buffers and the joint [prefix | suffix] sequence of pi0's mixture-of-transformers layers
 * (torch.cat / slicing in pi0_arch.py:163-165,196-206,258-262), alpha = sqrt(hidden) for the text embeddings. */
int b200_copy3d(const void* src, void* dst, int64_t B, int64_t rows, int64_t cols, int64_t src_bs, int64_t src_ld,
                int64_t dst_bs, int64_t dst_ld, float alpha, int accumulate, int dtype, void* stream);
/* SigLIP embeddings: out[b,p] = patches[b,p] + position_embedding[p] (HF SiglipVisionEmbeddings via
 * siglip_encoder.py:79-84); backward d_pos[p] += sum_b dout[b,p]. */
int b200_add_pos_fwd(const void* x, const void* pos, void* out, int64_t B, int64_t P, int64_t D, int dtype,
                     void* stream);
int b200_add_pos_bwd(const void* dout, float* d_pos, int64_t B, int64_t P, int64_t D, int dtype, void* stream);

/* ---- MemVLA memory path (memvla_arch.py) ---------------------------------------------------------------
 * Stateless counter-based dropout: out[r,c] = u(seed, r*cols+c) >= p ? x[r,c]/(1-p) : 0.  Replaces the dropout
 * inside F.scaled_dot_product_attention(dropout_p=...) and the two nn.Dropout of CrossTransformerBlock.ffn
 * (memvla_arch.py:97-103,122-124); applying it to the incoming gradient with the same (seed, shape) IS the
 * backward pass, so no mask is stored.  The random stream is this library's own (the reference's is cuRAND Philox
 * driven by torch's global generator): parity holds in distribution, bit parity at p = 0. */
int b200_dropout(const void* x, void* out, int64_t rows, int64_t cols, int64_t x_ld, int64_t out_ld, float p,
                 uint64_t seed, int dtype, void* stream);
/* BottleneckSE (memvla_arch.py:136-173).  se_reduce: out_f32[b,c] += scale * sum_p x[b,p,c] * (y ? y[b,p,c] : 1) —
 * AdaptiveAvgPool2d(1) with scale = 1/P, and the channel-gate gradient sum_p dout*x.  se_scale:
 * out[b,p,c] = x[b,p,c]*w[b,c] + (add ? add[b,c]*add_scale : 0) — `x * w` forward, and
 * dx = dout*w + dpool/P backward.  x is the projected vision feature map [B, P, C] (channel-last view of the
 * reference's [B,C,H,W] permute; the 1x1 convs are GEMMs on this layout). */
int b200_se_reduce(const void* x, const void* y, float* out_f32, int64_t B, int64_t P, int64_t C, float scale, int dtype,
                   void* stream);
int b200_se_scale(const void* x, const void* w, const void* add, void* out, int64_t B, int64_t P, int64_t C,
                  float add_scale, int dtype, void* stream);
/* GateFusion (memvla_arch.py:176-192): s = sigmoid(z); out = s*x1 + (1-s)*x2; backward gives dz, dx1, dx2. */
int b200_gate_fuse_fwd(const void* z, const void* x1, const void* x2, void* out, int64_t n, int dtype, void* stream);
int b200_gate_fuse_bwd(const void* dout, const void* z, const void* x1, const void* x2, void* dz, void* dx1, void* dx2,
                       int64_t n, int dtype, void* stream);

/* GPU-side input pipeline.
 * b200_image_preprocess: PreprocessRGB.__call__ with image_aspect_ratio='pad' (dexbotic/data/dataset/rgb_preprocess.py:
 * 13-44): expand2square with background (bg_r, bg_g, bg_b) -> PIL bicubic resize to out_size x out_size (HF
 * CLIPImageProcessor.preprocess) -> rescale / normalise.  src: uint8 [B, H, W, 3] device frames; coeff_* / bounds_*:
 * Pillow's fixed-point (22-bit) coefficient tables and (first, count) tap windows for the horizontal (square side ->
 * out_size) and the vertical pass, computed on the host exactly as Resample.c does; lut: float [3, 256] = the
 * processor's ((v * rescale) - mean_c) / std_c per channel; tmp: uint8 workspace [B, max(H, W), out_size, 3]; dst:
 * [B, 3, out_size, out_size] (dtype); dst_u8 (optional): the resized uint8 image [B, out_size, out_size, 3].  The uint8
 * image is bit-exact vs Pillow, the tensor bit-exact vs the processor. */
int b200_image_preprocess(const uint8_t* src, int64_t B, int64_t H, int64_t W, int64_t out_size, const int32_t* coeff_h,
                          const int32_t* bounds_h, int ksize_h, const int32_t* coeff_v, const int32_t* bounds_v,
                          int ksize_v, int bg_r, int bg_g, int bg_b, const float* lut, uint8_t* tmp, void* dst,
                          uint8_t* dst_u8, int dtype, void* stream);
/* ActionNorm._normalize (dexbotic/data/dataset/transform/action.py:268-275): quantile=1: (x - a) / (b - a + 1e-6) * 2 - 1
 * with a = min, b = max; quantile=0: (x - a) / (b + 1e-6) with a = mean, b = std.  float64 in, fp32 out. */
int b200_action_normalize(const double* x, const double* a, const double* b, float* out, int64_t rows, int64_t D,
                          int quantile, void* stream);

/* ---- data-parallel exchange over peer memory (ZeRO-1 reduce-scatter, parallel.ShardedDataParallel) -------------
 * own[0..n) = scale * (own + sum_r peer_ptrs[r][0..n))   bf16, fp32 accumulate in the order of peer_ptrs (deterministic).
 * peer_ptrs: HOST array of n_peers device pointers that are peer-mapped into this process (symmetric memory); the
 * caller orders the launch after a cross-rank barrier.  ctas: grid size (<= 0: 16) — kept small on purpose, the
 * persistent GEMM owns the SMs.  Replaces DeepSpeed's reduce-scatter of script/deepspeed/zero2.json. */
int b200_reduce_scatter_p2p(void* own, const void* const* peer_ptrs, int n_peers, int64_t n, float scale, int ctas,
                            void* stream);

#ifdef __cplusplus
}
#endif
#endif
