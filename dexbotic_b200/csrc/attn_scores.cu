// Fused attention-score kernels (bf16, head_dim <= 128, Sk <= 512):
//   mode 0 (forward):   P  = softmax_mask(alpha * Q K^T)                       -> bf16 probabilities
//   mode 1 (backward):  dS = alpha * P * (dO V^T - rowsum(P * dO V^T))         -> bf16
// Persistent: one CTA per SM walks a list of (batch*head, 128-row query tile) items.  The [128 x Sk] score block of an
// item lives in TMEM (<= 512 fp32 columns), so fp32 scores never touch HBM: one thread stages the Q (or dO) tile and
// the K (or V) rows with TMA and issues the tcgen05.mma sequence; 16 epilogue warps (four threads per query row) do
// the masked softmax (or its backward) straight out of TMEM and write bf16 rows.  Two items are in flight: item i+1
// is loaded and multiplied into a second TMEM region ([0, n) / [512 - n, 512) alternate) while the epilogue warps drain
// item i — whenever the two regions fit side by side (causal tiles see 128 / 256 / 320 keys, and the per-CTA order
// pairs them so that they do); otherwise the MMA waits for the previous epilogue.
// Replaces, per attention call, the fp32 score GEMM epilogue (write), the softmax kernel (read fp32 + write bf16)
// and, in backward, the fp32 dP round trip.  Causal tiles only load / multiply the keys they can see.
//
// Reference arithmetic replaced: F.scaled_dot_product_attention / eager softmax inside HF Qwen2/CLIP attention
// (called from dexbotic_arch.py:55-62, clip_encoder.py:50-54) and its autograd backward.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/dexbotic_b200_ops.h"
#include "common.h"
#include "ptx.cuh"

namespace b200 {

using bf16 = __nv_bfloat16;
// warp 0: TMA + MMA + TMEM alloc; warps 1-16: epilogue — warp w reads TMEM lane quarter (w & 3) and the 32-column
// chunks c with c % 4 == (w - 1) / 4, so every query row is shared by four threads (softmax is issue-bound: exp +
// masking per score; 4 warps per SM were the bottleneck)
constexpr int kAttnThreads = 32 + 16 * 32;
constexpr int kColGroups = 4;
constexpr int kAttnMaxSk = 512;

struct AttnKParams {
  CUtensorMap tmA, tmB;
  int Sq, Sk, H, G;
  int kblocks;  // ceil(head_dim / 64)
  float scale;
  int causal;
  int mode;
  const uint8_t* keymask;  // [B, Sk] or null
  const int* bid_q;        // [B, Sq] or null
  const int* bid_k;        // [B, Sk] or null
  const bf16* p_in;        // mode 1: probabilities
  bf16* out;               // mode 0: P, mode 1: dS
  long long ld;            // row stride of P / dS (elements); rows are [z, q]
  int m_tiles;
  int Z;                   // batch * heads
};

constexpr int kMaxItems = 1024;   // (z, tile) items one CTA may own

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void store8_bf16(bf16* p, const float* v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
  uint4 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  u.z = *reinterpret_cast<uint32_t*>(&c);
  u.w = *reinterpret_cast<uint32_t*>(&d);
  *reinterpret_cast<uint4*>(p) = u;
}

__device__ __forceinline__ int visible_keys(const AttnKParams& p, int tile) {
  return p.causal ? min(p.Sk, tile * 128 + 128) : p.Sk;
}

__global__ void __launch_bounds__(kAttnThreads, 1) attn_scores_kernel(const __grid_constant__ AttnKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar_load, bar_mma[2], bar_free[2];
  __shared__ uint32_t tmem_slot;
  __shared__ int n_items_s;
  __shared__ uint32_t items[kMaxItems];   // (z << 3) | tile, in this CTA's processing order
  __shared__ float red[2][kColGroups][128], red2[2][kColGroups][128];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0) {
    if (lane == 0) {
      mbar_init(&bar_load, 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bar_mma[i], 1);
        mbar_init(&bar_free[i], 16);       // one arrive per epilogue warp
      }
      mbar_fence_init();
      fence_proxy_async_smem();
      // ---- this CTA's items: z = c, c + G, ... for the first (Z / G) * G heads, plus an even share of the
      // (Z % G) * m_tiles left-over items.  Order: the widest and the narrowest tile of a head alternate (their score
      // blocks fit side by side in TMEM), the middle tiles follow.
      const int G = gridDim.x, c = blockIdx.x, m = p.m_tiles;
      const int nz = p.Z / G;
      int n = 0;
      if (m == 1) {
        for (int k = 0; k < nz; ++k) items[n++] = (uint32_t)(c + k * G) << 3;
      } else {
        for (int k = 0; k < nz; ++k) {
          const uint32_t z = (uint32_t)(c + k * G);
          items[n++] = (z << 3) | (uint32_t)(m - 1);
          items[n++] = (z << 3) | 0u;
        }
        for (int t = 1; t < m - 1; ++t)
          for (int k = 0; k < nz; ++k) items[n++] = ((uint32_t)(c + k * G) << 3) | (uint32_t)t;
      }
      const int left = (p.Z - nz * G) * m;
      const int r_lo = (int)(((long long)c * left + G - 1) / G), r_hi = (int)(((long long)(c + 1) * left + G - 1) / G);
      for (int r = r_lo; r < r_hi; ++r)
        items[n++] = ((uint32_t)(nz * G + r / m) << 3) | (uint32_t)(m - 1 - r % m);
      n_items_s = n;
    }
    __syncwarp();
    tmem_alloc(&tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int n_items = n_items_s;
  uint8_t* sA = smem;                              // kblocks x [128 rows][128 B]

  if (warp == 0) {
    if (lane == 0) {
      int prev_cols = 0;
      for (int it = 0; it < n_items; ++it) {
        const uint32_t item = items[it];
        const int tile = item & 7, z = item >> 3;
        const int b = z / p.H, h = z - b * p.H, kvh = h / p.G, m0 = tile * 128;
        const int n_eff = visible_keys(p, tile);
        const int n_mma = (n_eff + 15) & ~15;      // MMA N granularity
        const int n_box = (n_eff + 63) & ~63;      // TMA loads K/V in 64-row boxes
        const int cols = ((n_mma + 31) >> 5) << 5; // TMEM columns of this item's score block
        const int buf = it & 1, use = it >> 1;
        uint8_t* sB = smem + p.kblocks * 128 * 128;   // kblocks x [n_box rows][128 B]
        // the single operand buffer is free once the previous item's MMAs have retired
        if (it > 0) mbar_wait(&bar_mma[(it - 1) & 1], ((it - 1) >> 1) & 1);
        const uint32_t bytes = p.kblocks * (128 * 128 + n_box * 128);
        mbar_expect_tx(&bar_load, bytes);
        for (int kb = 0; kb < p.kblocks; ++kb) {
          tma_load_4d(sA + kb * 128 * 128, &p.tmA, &bar_load, kb * 64, m0, h, b);
          for (int n = 0; n < n_box; n += 64)
            tma_load_4d(sB + kb * n_box * 128 + n * 128, &p.tmB, &bar_load, kb * 64, n, kvh, b);
        }
        // this TMEM region was last used by item it-2; a region that does not fit next to item it-1's also waits for
        // that item's epilogue
        mbar_wait(&bar_free[buf], (use & 1) ^ 1);
        if (it > 0 && cols + prev_cols > 512) mbar_wait(&bar_free[buf ^ 1], ((it - 1) >> 1) & 1);
        mbar_wait(&bar_load, it & 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem + (buf ? (uint32_t)(512 - cols) : 0u);
        const uint32_t a_hi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO = 1024, version 1, SWIZZLE_128B
        const uint32_t lo0 = (16u >> 4) << 16;                         // LBO unused for K-major
        const uint32_t sa16 = smem_u32(sA) >> 4, sb16 = smem_u32(sB) >> 4;
        for (int kb = 0; kb < p.kblocks; ++kb) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t a_lo = lo0 | (sa16 + kb * (128 * 128 >> 4) + k * 2);
            for (int n0 = 0; n0 < n_mma; n0 += 256) {
              const int ncur = min(256, n_mma - n0);
              const uint32_t b_lo = lo0 | (sb16 + kb * (n_box * 128 >> 4) + n0 * (128 >> 4) + k * 2);
              const uint32_t idesc = umma_idesc(1u, 0, 0, 128, ncur);
              umma_issue<1, false>(d_tmem + n0, a_lo, a_hi, b_lo, a_hi, idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
        }
        umma_commit(&bar_mma[buf]);
        prev_cols = cols;
      }
    }
  } else {
    // ------------------------------------------------------------- epilogue
    const int quarter = warp & 3;
    const int cg = (warp - 1) >> 2;                 // column group: chunks cg, cg + 4, cg + 8, ...
    const int rt = quarter * 32 + lane;             // row inside the tile == TMEM lane
    const float sl2 = p.scale * 1.4426950408889634f;   // exp(x * scale - m) = exp2(x * sl2 - m * log2e)
    constexpr int kMaxOwn = (kAttnMaxSk / 32 + kColGroups - 1) / kColGroups;   // chunks one thread can own (4)
    for (int it = 0; it < n_items; ++it) {
      const uint32_t item = items[it];
      const int tile = item & 7, z = item >> 3;
      const int b = z / p.H, m0 = tile * 128;
      const int n_eff = visible_keys(p, tile);
      const int n_mma = (n_eff + 15) & ~15;
      const int n_chunks = (n_mma + 31) >> 5;
      const int buf = it & 1;
      const int q = m0 + rt;
      const bool row_ok = q < p.Sq;
      const uint32_t trow = tmem + (buf ? (uint32_t)(512 - n_chunks * 32) : 0u) +
                            (static_cast<uint32_t>(quarter * 32) << 16);
      const uint8_t* km = p.keymask != nullptr ? p.keymask + (size_t)b * p.Sk : nullptr;
      const int* bk = (!p.causal && p.bid_k != nullptr) ? p.bid_k + (size_t)b * p.Sk : nullptr;
      const int bq = (bk != nullptr && row_ok) ? p.bid_q[(size_t)b * p.Sq + q] : 0;
      const int limit = p.causal ? min(n_eff, q + 1) : n_eff;
      const long long row_off = ((long long)z * p.Sq + q) * p.ld;
      bf16* orow = p.out + row_off;
      const bf16* prow = p.mode == 1 ? p.p_in + row_off : nullptr;
      float (*rd)[128] = red[it & 1];               // alternate: a fast warp may already write the next item's partials
      float (*rs)[128] = red2[it & 1];
      // validity bits of this thread's chunks: the key-padding byte of key c*32+j is loaded once by lane j and
      // shared with a ballot; the causal / length limit is a per-row bit count
      uint32_t okbits[kMaxOwn];
#pragma unroll
      for (int i = 0; i < kMaxOwn; ++i) {
        const int c = cg + i * kColGroups;
        uint32_t w = 0;
        if (c < n_chunks) {
          const int k = c * 32 + lane;
          const bool kv = k < p.Sk && (km == nullptr || km[k] != 0);
          w = __ballot_sync(0xffffffffu, kv);
          const int rem = limit - c * 32;
          w &= rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
          if (bk != nullptr) {                      // generic block-id rule (pi0-style masks): per-element compare
            uint32_t w2 = 0;
            for (int j = 0; j < 32; ++j) {
              const int kk = c * 32 + j;
              if (kk < p.Sk && bk[kk] <= bq) w2 |= 1u << j;
            }
            w &= w2;
          }
        }
        okbits[i] = w;
      }
      mbar_wait(&bar_mma[buf], (it >> 1) & 1);
      tc_fence_after();
      uint32_t r[32];
      if (p.mode == 0) {
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < kMaxOwn; ++i) {
          const int c = cg + i * kColGroups;
          if (c >= n_chunks) break;
          tmem_ld_32x32(trow + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if ((okbits[i] >> j) & 1u) mx = fmaxf(mx, __uint_as_float(r[j]) * sl2);
        }
        rd[cg][rt] = mx;
        asm volatile("bar.sync 1, 512;" ::: "memory");
        mx = fmaxf(fmaxf(rd[0][rt], rd[1][rt]), fmaxf(rd[2][rt], rd[3][rt]));
        float sum = 0.0f;
#pragma unroll
        for (int i = 0; i < kMaxOwn; ++i) {
          const int c = cg + i * kColGroups;
          if (c >= n_chunks) break;
          tmem_ld_32x32(trow + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if ((okbits[i] >> j) & 1u) sum += fast_exp2(__uint_as_float(r[j]) * sl2 - mx);
        }
        rs[cg][rt] = sum;
        asm volatile("bar.sync 1, 512;" ::: "memory");
        sum = (rs[0][rt] + rs[1][rt]) + (rs[2][rt] + rs[3][rt]);
        const float inv = sum > 0.0f ? 1.0f / sum : 0.0f;
#pragma unroll
        for (int i = 0; i < kMaxOwn; ++i) {
          const int c = cg + i * kColGroups;
          if (c >= n_chunks) break;
          tmem_ld_32x32(trow + c * 32, r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j)
            v[j] = ((okbits[i] >> j) & 1u) ? fast_exp2(__uint_as_float(r[j]) * sl2 - mx) * inv : 0.0f;
          if (row_ok) {
#pragma unroll
            for (int g = 0; g < 4; ++g)
              if (c * 32 + g * 8 < p.ld) store8_bf16(orow + c * 32 + g * 8, v + g * 8);
          }
        }
      } else {
        float delta = 0.0f;
#pragma unroll
        for (int i = 0; i < kMaxOwn; ++i) {
          const int c = cg + i * kColGroups;
          if (c >= n_chunks) break;
          tmem_ld_32x32(trow + c * 32, r);
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int k0 = c * 32 + g * 8;
              if (k0 < limit) {
                const uint4 u = *reinterpret_cast<const uint4*>(prow + k0);
                const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 f = __bfloat1622float2(hh[j]);
                  delta += f.x * __uint_as_float(r[g * 8 + 2 * j]) + f.y * __uint_as_float(r[g * 8 + 2 * j + 1]);
                }
              }
            }
          }
        }
        rd[cg][rt] = delta;
        asm volatile("bar.sync 1, 512;" ::: "memory");
        delta = (rd[0][rt] + rd[1][rt]) + (rd[2][rt] + rd[3][rt]);
#pragma unroll
        for (int i = 0; i < kMaxOwn; ++i) {
          const int c = cg + i * kColGroups;
          if (c >= n_chunks) break;
          tmem_ld_32x32(trow + c * 32, r);
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int k0 = c * 32 + g * 8;
              if (k0 >= p.ld) continue;
              float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
              if (k0 < limit) {
                const uint4 u = *reinterpret_cast<const uint4*>(prow + k0);
                const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 f = __bfloat1622float2(hh[j]);   // P is exactly 0 at masked keys
                  v[2 * j] = p.scale * f.x * (__uint_as_float(r[g * 8 + 2 * j]) - delta);
                  v[2 * j + 1] = p.scale * f.y * (__uint_as_float(r[g * 8 + 2 * j + 1]) - delta);
                }
              }
              store8_bf16(orow + k0, v);
            }
          }
        }
      }
      // this item's TMEM region may be overwritten
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_free[buf]);
      // columns beyond this tile's visible keys are zero (the P V / dS K GEMMs read whole rows)
      if (row_ok) {
        const float zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int k0 = n_chunks * 32 + cg * 8; k0 < p.ld; k0 += 8 * kColGroups) store8_bf16(orow + k0, zero);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode_qk(CUtensorMap* m, const void* ptr, uint64_t hd, uint64_t S, uint64_t heads, uint64_t B, int64_t ld,
                     int64_t s_head, int64_t s_batch, uint32_t box_rows, const char* what) {
  static EncodeTiledFn2 fn = nullptr;
  if (fn == nullptr) {
    void* pfn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &pfn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn2>(pfn);
  }
  B200_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[4] = {hd, S, heads, B};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)s_head * 2, (cuuint64_t)s_batch * 2};
  cuuint32_t box[4] = {64, box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < 3; ++i) B200_CHECK(strides[i] % 16 == 0, "%s: stride %d not a multiple of 16 bytes", what, i + 1);
  B200_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "%s: base pointer not 16-byte aligned", what);
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_CHECK(r == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled failed (%d)", what, (int)r);
  return 0;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_attn_scores(const void* a, const void* bmat, const void* p_in, void* out, int64_t B, int64_t H,
                                int64_t KVH, int64_t Sq, int64_t Sk, int64_t head_dim, int64_t a_ld, int64_t a_s_head,
                                int64_t a_s_batch, int64_t b_ld, int64_t b_s_head, int64_t b_s_batch, int64_t p_ld,
                                float scale, int causal, const uint8_t* keymask, const int32_t* bid_q,
                                const int32_t* bid_k, int mode, void* stream) {
  B200_CHECK(head_dim <= 128 && head_dim % 8 == 0, "attn_scores: head_dim=%lld unsupported (<= 128)", (long long)head_dim);
  B200_CHECK(Sk <= kAttnMaxSk && Sk > 0 && Sq > 0, "attn_scores: Sk=%lld unsupported (<= 512)", (long long)Sk);
  B200_CHECK(H % KVH == 0 && p_ld % 8 == 0 && p_ld >= Sk, "attn_scores: bad head / ld geometry");
  B200_CHECK((bid_q == nullptr) == (bid_k == nullptr), "attn_scores: bid_q and bid_k go together");
  B200_CHECK(mode == 0 || p_in != nullptr, "attn_scores: backward mode needs the probabilities");
  AttnKParams kp;
  memset(&kp, 0, sizeof(kp));
  if (encode_qk(&kp.tmA, a, head_dim, Sq, H, B, a_ld, a_s_head, a_s_batch, 128, "attn A")) return 1;
  if (encode_qk(&kp.tmB, bmat, head_dim, Sk, KVH, B, b_ld, b_s_head, b_s_batch, 64, "attn B")) return 1;
  kp.Sq = (int)Sq;
  kp.Sk = (int)Sk;
  kp.H = (int)H;
  kp.G = (int)(H / KVH);
  kp.kblocks = (int)ceil_div(head_dim, 64);
  kp.scale = scale;
  kp.causal = causal;
  kp.mode = mode;
  kp.keymask = keymask;
  kp.bid_q = bid_q;
  kp.bid_k = bid_k;
  kp.p_in = (const bf16*)p_in;
  kp.out = (bf16*)out;
  kp.ld = p_ld;
  kp.m_tiles = (int)ceil_div(Sq, 128);
  const int sk_pad = (int)((Sk + 63) / 64 * 64);
  const int smem = 1024 + kp.kblocks * (128 * 128 + sk_pad * 128);
  static unsigned long long configured = 0;      // one bit per device ordinal: the attribute is per device
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || !((configured >> dev) & 1ull)) {
    B200_CUDA(cudaFuncSetAttribute(attn_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (dev < 64) configured |= 1ull << dev;
  }
  kp.Z = (int)(B * H);
  B200_CHECK(kp.m_tiles <= 4 && B * H < (1ll << 28), "attn_scores: geometry out of range");
  const long long total = (long long)kp.Z * kp.m_tiles;
  const unsigned grid = (unsigned)(total < num_sms() ? total : num_sms());
  B200_CHECK((kp.Z / (int)grid + 1) * kp.m_tiles + kp.m_tiles <= kMaxItems,
             "attn_scores: %lld items exceed the per-CTA list (use the unfused path)", total);
  attn_scores_kernel<<<grid, kAttnThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(kp);
  B200_LAUNCH_OK();
  return 0;
}
