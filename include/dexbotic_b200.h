/*
 * dexbotic_b200 — C-ABI of the B200 (sm_100a) kernel library behind Dexbotic's VLA
 * training hot path (ViT encoder -> projector -> image-token splice -> LLM decoder ->
 * action head, forward + backward + optimizer).
 *
 * The reference (dexmal/dexbotic) is pure Python and has NO native interface for this
 * path; every entry point below therefore cites the reference *Python*
 * call site (file:line under /root/reference) whose arithmetic it replaces.  A
 * replacement backend binds these symbols with ctypes.
 *
 * Conventions
 *   - plain pointers + sizes; no torch types cross this boundary
 *   - all pointers are DEVICE pointers unless the name ends in _host
 *   - every call enqueues on `stream` (a cudaStream_t passed as void*) and returns
 *     immediately; nothing here allocates, synchronises or creates streams
 *   - return value: 0 = ok, non-zero = error; b200_last_error() gives the message
 *   - dtype codes: B200_BF16 = 0, B200_F32 = 1
 */
#ifndef DEXBOTIC_B200_H
#define DEXBOTIC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_BF16 0
#define B200_F32 1

/* epilogue activations (applied to alpha*acc + bias, before the residual add) */
#define B200_ACT_NONE 0
#define B200_ACT_GELU_ERF 1   /* nn.GELU()            — mm_projector/builder.py:75 */
#define B200_ACT_GELU_TANH 2  /* nn.GELU("tanh")      — cogact/action_model/dit.py:150; Gemma/SigLIP MLP */
#define B200_ACT_QUICK_GELU 3 /* x*sigmoid(1.702x)    — HF CLIP MLP (clip_encoder.py:26) */
#define B200_ACT_SILU 4       /* nn.SiLU()            — dit.py:31, Qwen2/Llama MLP */
#define B200_ACT_RELU 5

int b200_version(void);
const char* b200_last_error(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
int64_t b200_launch_count(void);
/* Number of SMs the persistent GEMM grid may occupy (0 = all of them).  Host-side knob for the data-parallel overlap
 * (HF Trainer's DDP all-reduce during backward, trainer.py:110,121): while an NCCL collective holds R SMs, a grid
 * sized for all SMs queues its last CTAs behind the collective. */
int b200_set_gemm_sm_limit(int sms);

/* ------------------------------------------------------------------ GEMM ----------
 * D[z] = epilogue( alpha * sum_seg A[zA(seg)] x B[zB(seg)] )      (tcgen05 / TMEM / TMA)
 *
 * Replaces every nn.Linear / matmul on the path:  HF Qwen2/Llama/Gemma q,k,v,o,gate,up,
 * down projections (dexbotic_arch.py:55-62 builds them via AutoModel), CLIP/SigLIP
 * encoder GEMMs (clip_encoder.py:26,50-54), mm_projector (mm_projector/builder.py:71-79),
 * timm Attention/Mlp inside DiT (dit.py:145-157), QK^T / PV of every attention, and all
 * their dgrad / wgrad forms in backward.
 *
 * Operands are described as up-to-4-D strided tensors so that batched attention GEMMs
 * (per batch x head, GQA head mapping, GQA-group reduction in dK/dV) use the same kernel:
 *   A  K-major : dims (K, M, a_z2, z_hi)   element strides (1, a_ld, a_s2, a_s3)
 *   A  MN-major: dims (M, K, a_z2, z_hi)   element strides (1, a_ld, a_s2, a_s3)
 *   B  likewise with N.   D: dims (N, M, z_lo, z_hi) strides (1, d_ld, d_s2, d_s3)
 * For output batch index (zl, zh) and K-segment seg (0..k_segs-1):
 *   A's dim-2 coordinate = (zl / a_div) * a_mul + seg * a_seg ; dim-3 coordinate = zh.
 * A plain linear layer uses z_lo = z_hi = k_segs = 1.
 *
 * dual_b: B and B2 are two [N, K] K-major matrices (gate_proj / up_proj); the kernel
 * accumulates both against the same A tile and writes act(A B^T) * (A B2^T)  (SwiGLU /
 * GeGLU, HF Qwen2MLP / GemmaMLP), optionally also storing both pre-activations to
 * aux / aux2 for the backward pass.
 */
typedef struct {
  const void* a;
  const void* b;
  const void* b2; /* dual_b only */
  void* d;
  int32_t ab_dtype; /* B200_BF16 (kind::f16) or B200_F32 (kind::tf32) */
  int32_t d_dtype;
  int32_t a_mn_major; /* 0: K contiguous, 1: M contiguous */
  int32_t b_mn_major; /* 0: K contiguous, 1: N contiguous */
  int64_t m, n, k;    /* per batch entry, per K segment */
  int64_t a_ld, a_s2, a_s3;
  int64_t b_ld, b_s2, b_s3;
  int64_t d_ld, d_s2, d_s3;
  int32_t a_z2, b_z2; /* extents of dim 2 of A and B */
  int32_t z_lo, z_hi;
  int32_t a_div, a_mul, a_seg;
  int32_t b_div, b_mul, b_seg;
  int32_t k_segs;
  float alpha;       /* scale of the accumulator; 0 (a zero-initialised struct) means 1.0 -- a product scaled by zero
                        is a memset, not a GEMM, so the value is free to mean "unset" */
  const void* bias; /* [n] or NULL */
  int32_t bias_dtype;
  const void* residual; /* same logical shape as D, or NULL; may alias d (accumulate) */
  int32_t res_dtype;
  int64_t res_ld, res_s2, res_s3;
  void* aux;  /* optional pre-activation copy (dtype = d_dtype), z-strides as D */
  void* aux2; /* dual_b: second pre-activation */
  int64_t aux_ld;
  int32_t act;
  int32_t dual_b;
  int32_t block_n; /* 0 = auto; 64 / 128 / 256 */
  /* glu_bwd: the accumulator is dh = d(act(g) * u); the epilogue writes D = dg = dh * u * act'(g) and d2 = du =
   * dh * act(g) (bf16, rows of glu_ld elements; g / u may alias D / d2: every element is read before it is written
   * by the same thread).  Replaces the glu_bwd kernel after the down-projection dgrad (HF Qwen2MLP / GemmaMLP). */
  int32_t glu_bwd;
  const void* glu_g;
  const void* glu_u;
  void* d2;
  int64_t glu_ld;
} b200_gemm_args;

int b200_gemm(const b200_gemm_args* args, void* stream);

/* Reference-quality SIMT GEMM for shapes TMA cannot describe (K or N not a multiple of
 * 16 bytes: DiT x_embedder K=7, final_layer N=7 — dit.py:110,172).  Row-major, fp32
 * accumulate: D[M,N] = act(A[M,K] * B[N,K]^T (or B[K,N] if b_mn_major) + bias) + res. */
int b200_gemm_simt(const b200_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEXBOTIC_B200_H */
