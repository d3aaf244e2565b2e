// HBM-bound kernels of the VLA hot path: norms, RoPE, gated activations, masked softmax,
// reductions, AdamW.  All global traffic is 128-bit vectorised and coalesced; grids are sized
// in multiples of the SM count with grid-stride loops.  Reference call sites are cited in
// include/dexbotic_b200_ops.h next to each entry point.
#include "../../include/dexbotic_b200_ops.h"
#include "common.h"
#include "ptx.cuh"
#include "vec.cuh"

namespace b200 {

using bf16 = __nv_bfloat16;
constexpr int kMaxPacks = 4;  // packs of 8 columns per thread kept in registers (D <= 8 * 256 * 4)

// threads per row-block for the norm kernels: ~2 packs of 8 columns per thread, 64..256 threads
static inline int norm_threads(int64_t D) {
  int64_t t = ((D / 8 + 1) / 2 + 31) / 32 * 32;
  return (int)(t < 64 ? 64 : (t > 256 ? 256 : t));
}

static inline int grid_for_rows(int64_t rows, int per_sm = 8) {
  int64_t g = (int64_t)num_sms() * per_sm;
  return (int)(rows < g ? rows : g);
}

// ------------------------------------------------------------------ RMSNorm
// Eight elements held as loaded (prefetch registers: the NEXT row's loads are in flight while this row is reduced).
template <typename T>
struct Raw8;
template <>
struct Raw8<bf16> {
  uint4 u;
  __device__ __forceinline__ void load(const bf16* p) { u = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void unpack(float (&v)[8]) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    }
  }
};
template <>
struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = reinterpret_cast<const float4*>(p)[0];
    b = reinterpret_cast<const float4*>(p)[1];
  }
  __device__ __forceinline__ void unpack(float (&v)[8]) const {
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};

// One block per row, kPacks packs of 8 columns per thread held in registers: x is read ONCE, and the loads of the
// block's next row are issued before this row's reduction (a row is ~7 KB: with one row in flight per block the kernel
// was latency-bound at 0.59 of the copy bandwidth).
template <typename T, int kPacks>
__global__ void __launch_bounds__(256) rmsnorm_fwd_kernel(const T* __restrict__ x, const T* __restrict__ w,
                                                          T* __restrict__ y, float* __restrict__ rstd, int M, int D,
                                                          float eps, int unit_offset) {
  __shared__ float red[33];
  Raw8<T> nx[kPacks];
  int row = blockIdx.x;
  if (row < M) {
#pragma unroll
    for (int k = 0; k < kPacks; ++k) {
      const int i = (threadIdx.x + k * blockDim.x) * 8;
      if (i < D) nx[k].load(x + (size_t)row * D + i);
    }
  }
  for (; row < M; row += gridDim.x) {
    float xv[kPacks][8];
    float ss = 0.0f;
#pragma unroll
    for (int k = 0; k < kPacks; ++k) {
      const int i = (threadIdx.x + k * blockDim.x) * 8;
      if (i < D) {
        nx[k].unpack(xv[k]);
#pragma unroll
        for (int j = 0; j < 8; ++j) ss += xv[k][j] * xv[k][j];
      }
    }
    const int nrow = row + gridDim.x;
    if (nrow < M) {
#pragma unroll
      for (int k = 0; k < kPacks; ++k) {
        const int i = (threadIdx.x + k * blockDim.x) * 8;
        if (i < D) nx[k].load(x + (size_t)nrow * D + i);
      }
    }
    ss = block_sum(ss, red);
    const float r = rsqrtf(ss / (float)D + eps);
    if (threadIdx.x == 0 && rstd != nullptr) rstd[row] = r;
    T* yr = y + (size_t)row * D;
#pragma unroll
    for (int k = 0; k < kPacks; ++k) {
      const int i = (threadIdx.x + k * blockDim.x) * 8;
      if (i < D) {
        float g[8], o[8];
        Pack8<T>::load(w + i, g);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          o[j] = unit_offset ? (xv[k][j] * r) * (1.0f + g[j]) : g[j] * round_to<T>(xv[k][j] * r);
        Pack8<T>::store(yr + i, o);
      }
    }
  }
}

// Forward with the rows staged in shared memory by the bulk-copy engine (see rmsnorm_bwd_staged_kernel): a ring of
// kStages rows of x per block, so kStages - 1 rows per block are in flight instead of the one the prefetch registers hold.
template <typename T, int kPacks, int kStages>
__global__ void __launch_bounds__(256) rmsnorm_fwd_staged_kernel(const T* __restrict__ x, const T* __restrict__ w,
                                                                 T* __restrict__ y, float* __restrict__ rstd, int M,
                                                                 int D, float eps, int unit_offset) {
  extern __shared__ __align__(128) unsigned char stage_mem[];
  __shared__ float red[33];
  __shared__ __align__(8) uint64_t full[kStages];
  const uint32_t row_bytes = (uint32_t)D * (uint32_t)sizeof(T);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      const long long row = (long long)blockIdx.x + (long long)s * gridDim.x;
      if (row < M) {
        mbar_expect_tx(&full[s], row_bytes);
        bulk_copy_g2s(stage_mem + (size_t)s * row_bytes, x + (size_t)row * D, row_bytes, &full[s]);
      }
    }
  }
  int it = 0;
  for (int row = blockIdx.x; row < M; row += gridDim.x, ++it) {
    const int s = it % kStages;
    mbar_wait(&full[s], (uint32_t)(it / kStages) & 1u);
    const T* sx = reinterpret_cast<const T*>(stage_mem + (size_t)s * row_bytes);
    float xv[kPacks][8];
    float ss = 0.0f;
#pragma unroll
    for (int k = 0; k < kPacks; ++k) {
      const int i = (threadIdx.x + k * blockDim.x) * 8;
      if (i < D) {
        Pack8<T>::load(sx + i, xv[k]);
#pragma unroll
        for (int j = 0; j < 8; ++j) ss += xv[k][j] * xv[k][j];
      }
    }
    ss = block_sum(ss, red);                   // its barriers order every thread's reads of stage s before the refill
    if (threadIdx.x == 0) {
      const long long nrow = (long long)row + (long long)kStages * gridDim.x;
      if (nrow < M) {
        mbar_expect_tx(&full[s], row_bytes);
        bulk_copy_g2s(stage_mem + (size_t)s * row_bytes, x + (size_t)nrow * D, row_bytes, &full[s]);
      }
    }
    const float r = rsqrtf(ss / (float)D + eps);
    if (threadIdx.x == 0 && rstd != nullptr) rstd[row] = r;
    T* yr = y + (size_t)row * D;
#pragma unroll
    for (int k = 0; k < kPacks; ++k) {
      const int i = (threadIdx.x + k * blockDim.x) * 8;
      if (i < D) {
        float g[8], o[8];
        Pack8<T>::load(w + i, g);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          o[j] = unit_offset ? (xv[k][j] * r) * (1.0f + g[j]) : g[j] * round_to<T>(xv[k][j] * r);
        Pack8<T>::store(yr + i, o);
      }
    }
  }
}

// dx = r*g - x*r^3*mean(g.x),  g = dy*w_eff ;  dw[col] += sum_rows dy * x * r.
// x / dy are read once per row (kept in registers between the two phases).  dw partials go to a per-block row of
// `ws` (no atomics; reduced by colsum afterwards) when a workspace is given, else fp32 atomics.
template <typename T, int kPacks>
__global__ void __launch_bounds__(256, kPacks <= 2 ? 3 : 2) rmsnorm_bwd_kernel(
    const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ w, const float* __restrict__ rstd, T* dx,
    float* __restrict__ dw, float* __restrict__ ws, int M, int D, int unit_offset, int accumulate_dx) {
  // kPacks = packs of 8 columns per thread (2 covers D <= 16 * blockDim: every decoder / tower width); the row's
  // x / dy / dw-partials live in registers, the weight vector is re-read per row (7 KB, L1-resident) — with it in
  // registers too the kernel spilled ~0.7 KB per thread.
  __shared__ float red[33];
  const float uo = unit_offset ? 1.0f : 0.0f;
  float wacc[kPacks][8];
#pragma unroll
  for (int k = 0; k < kPacks; ++k) {
#pragma unroll
    for (int j = 0; j < 8; ++j) wacc[k][j] = 0.0f;
  }
  Raw8<T> nx[kPacks], ndy[kPacks];      // the next row's x / dy: in flight during this row's reduction
  float nr = 0.0f;
  if ((int)blockIdx.x < M) {
    nr = rstd[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kPacks; ++k) {
      const int i = (threadIdx.x + k * blockDim.x) * 8;
      if (i < D) {
        nx[k].load(x + (size_t)blockIdx.x * D + i);
        ndy[k].load(dy + (size_t)blockIdx.x * D + i);
      }
    }
  }
  for (int row = blockIdx.x; row < M; row += gridDim.x) {
    T* dxr = dx + (size_t)row * D;
    const float r = nr;
    float xv[kPacks][8], gv[kPacks][8];      // gv = dy * w_eff
    float c = 0.0f;
#pragma unroll
    for (int k = 0; k < kPacks; ++k) {
      const int i = (threadIdx.x + k * blockDim.x) * 8;
      if (i < D) {
        nx[k].unpack(xv[k]);
        ndy[k].unpack(gv[k]);
      }
    }
    const int nrow = row + gridDim.x;
    if (nrow < M) {
      nr = rstd[nrow];
#pragma unroll
      for (int k = 0; k < kPacks; ++k) {
        const int i = (threadIdx.x + k * blockDim.x) * 8;
        if (i < D) {
          nx[k].load(x + (size_t)nrow * D + i);
          ndy[k].load(dy + (size_t)nrow * D + i);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < kPacks; ++k) {
      const int i = (threadIdx.x + k * blockDim.x) * 8;
      if (i < D) {
        float we[8];
        Pack8<T>::load(w + i, we);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          wacc[k][j] += gv[k][j] * xv[k][j] * r;
          gv[k][j] *= we[j] + uo;
          c += gv[k][j] * xv[k][j];
        }
      }
    }
    c = block_sum(c, red);
    const float coef = c * r * r * r / (float)D;
#pragma unroll
    for (int k = 0; k < kPacks; ++k) {
      const int i = (threadIdx.x + k * blockDim.x) * 8;
      if (i < D) {
        float o[8];
        if (accumulate_dx) Pack8<T>::load(dxr + i, o);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float t = r * gv[k][j] - xv[k][j] * coef;
          o[j] = accumulate_dx ? o[j] + t : t;
        }
        Pack8<T>::store(dxr + i, o);
      }
    }
  }
  if (dw != nullptr) {
#pragma unroll
    for (int k = 0; k < kPacks; ++k) {
      const int i = (threadIdx.x + k * blockDim.x) * 8;
      if (i < D) {
        if (ws != nullptr) {
          Pack8<float>::store(ws + (size_t)blockIdx.x * D + i, wacc[k]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) atomicAdd(dw + i + j, wacc[k][j]);
        }
      }
    }
  }
}

// The same backward with the rows STAGED IN SHARED MEMORY by the bulk-copy engine (cp.async.bulk, mbarrier
// complete_tx): a ring of kStages rows of (x | dy | dx-to-accumulate) per block, refilled by one thread as soon as the
// block has consumed a stage.  The register version above can only keep ONE next row in flight per block (the prefetch
// registers), i.e. ~42 KB per SM at 3 blocks of a 7 KB row x 2 arrays: by Little's law (6.5 TB/s x ~1-2 us) that is half
// of what the HBM needs — measured 0.51 of the copy peak, and the accumulate variant issued its dx load AFTER the
// reduction.  Here kStages - 1 rows of all arrays are in flight per block without costing a register.
template <typename T, int kPacks, int kStages>
__global__ void __launch_bounds__(256, 3) rmsnorm_bwd_staged_kernel(
    const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ w, const float* __restrict__ rstd, T* dx,
    float* __restrict__ dw, float* __restrict__ ws, int M, int D, int unit_offset, int accumulate_dx) {
  extern __shared__ __align__(128) unsigned char stage_mem[];
  __shared__ float red[33];
  __shared__ __align__(8) uint64_t full[kStages];
  const float uo = unit_offset ? 1.0f : 0.0f;
  const uint32_t row_bytes = (uint32_t)D * (uint32_t)sizeof(T);
  const uint32_t stage_bytes = (accumulate_dx ? 3u : 2u) * row_bytes;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  auto issue = [&](int row, int s) {          // one thread: arm the stage's barrier, then the 2-3 row copies
    unsigned char* st = stage_mem + (size_t)s * stage_bytes;
    mbar_expect_tx(&full[s], stage_bytes);
    bulk_copy_g2s(st, x + (size_t)row * D, row_bytes, &full[s]);
    bulk_copy_g2s(st + row_bytes, dy + (size_t)row * D, row_bytes, &full[s]);
    if (accumulate_dx) bulk_copy_g2s(st + 2 * row_bytes, dx + (size_t)row * D, row_bytes, &full[s]);
  };
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      const long long row = (long long)blockIdx.x + (long long)s * gridDim.x;
      if (row < M) issue((int)row, s);
    }
  }
  float wacc[kPacks][8];
#pragma unroll
  for (int k = 0; k < kPacks; ++k) {
#pragma unroll
    for (int j = 0; j < 8; ++j) wacc[k][j] = 0.0f;
  }
  int it = 0;
  for (int row = blockIdx.x; row < M; row += gridDim.x, ++it) {
    const int s = it % kStages;
    const float r = rstd[row];
    mbar_wait(&full[s], (uint32_t)(it / kStages) & 1u);
    const T* sx = reinterpret_cast<const T*>(stage_mem + (size_t)s * stage_bytes);
    const T* sdy = sx + D;
    const T* sdx = sdy + D;
    float xv[kPacks][8], gv[kPacks][8];      // gv = dy * w_eff
    Raw8<T> od[kPacks];
    float c = 0.0f;
#pragma unroll
    for (int k = 0; k < kPacks; ++k) {
      const int i = (threadIdx.x + k * blockDim.x) * 8;
      if (i < D) {
        float we[8];
        Pack8<T>::load(sx + i, xv[k]);
        Pack8<T>::load(sdy + i, gv[k]);
        if (accumulate_dx) od[k].load(sdx + i);
        Pack8<T>::load(w + i, we);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          wacc[k][j] += gv[k][j] * xv[k][j] * r;
          gv[k][j] *= we[j] + uo;
          c += gv[k][j] * xv[k][j];
        }
      }
    }
    c = block_sum(c, red);                    // its barriers also order every thread's reads of stage s before the refill
    if (threadIdx.x == 0) {
      const long long nrow = (long long)row + (long long)kStages * gridDim.x;
      if (nrow < M) issue((int)nrow, s);
    }
    const float coef = c * r * r * r / (float)D;
    T* dxr = dx + (size_t)row * D;
#pragma unroll
    for (int k = 0; k < kPacks; ++k) {
      const int i = (threadIdx.x + k * blockDim.x) * 8;
      if (i < D) {
        float o[8];
        if (accumulate_dx) od[k].unpack(o);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float t = r * gv[k][j] - xv[k][j] * coef;
          o[j] = accumulate_dx ? o[j] + t : t;
        }
        Pack8<T>::store(dxr + i, o);
      }
    }
  }
  if (dw != nullptr) {
#pragma unroll
    for (int k = 0; k < kPacks; ++k) {
      const int i = (threadIdx.x + k * blockDim.x) * 8;
      if (i < D) {
        if (ws != nullptr) {
          Pack8<float>::store(ws + (size_t)blockIdx.x * D + i, wacc[k]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) atomicAdd(dw + i + j, wacc[k][j]);
        }
      }
    }
  }
}

// ---------------------------------------------------------------- LayerNorm
template <typename T>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const T* __restrict__ x, const T* __restrict__ w,
                                                            const T* __restrict__ b, T* __restrict__ y,
                                                            float* __restrict__ mean, float* __restrict__ rstd, int M,
                                                            int D, float eps) {
  __shared__ float red[33];
  for (int row = blockIdx.x; row < M; row += gridDim.x) {
    const T* xr = x + (size_t)row * D;
    T* yr = y + (size_t)row * D;
    float s = 0.0f;
    for (int i = threadIdx.x * 8; i < D; i += blockDim.x * 8) {
      float v[8];
      Pack8<T>::load(xr + i, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[j];
    }
    const float mu = block_sum(s, red) / (float)D;
    float ss = 0.0f;
    for (int i = threadIdx.x * 8; i < D; i += blockDim.x * 8) {
      float v[8];
      Pack8<T>::load(xr + i, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += (v[j] - mu) * (v[j] - mu);
    }
    const float r = rsqrtf(block_sum(ss, red) / (float)D + eps);
    if (threadIdx.x == 0) {
      if (mean != nullptr) mean[row] = mu;
      if (rstd != nullptr) rstd[row] = r;
    }
    for (int i = threadIdx.x * 8; i < D; i += blockDim.x * 8) {
      float v[8], o[8];
      Pack8<T>::load(xr + i, v);
      float wv[8], bv[8];
      if (w != nullptr) Pack8<T>::load(w + i, wv);
      if (b != nullptr) Pack8<T>::load(b + i, bv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = (v[j] - mu) * r;
        if (w != nullptr) t *= wv[j];
        if (b != nullptr) t += bv[j];
        o[j] = t;
      }
      Pack8<T>::store(yr + i, o);
    }
  }
}

template <typename T, int kPacks>
__global__ void __launch_bounds__(256, kPacks <= 2 ? 2 : 1) layernorm_bwd_kernel(
    const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ w, const float* __restrict__ mean,
    const float* __restrict__ rstd, T* dx, float* __restrict__ dw, float* __restrict__ db, float* __restrict__ ws, int M,
    int D, int accumulate_dx) {
  // same register budget as rmsnorm_bwd: x-hat / g / dw / db partials in registers, the weight re-read per row
  __shared__ float red[33];
  float wacc[kPacks][8], bacc[kPacks][8];
#pragma unroll
  for (int k = 0; k < kPacks; ++k) {
#pragma unroll
    for (int j = 0; j < 8; ++j) wacc[k][j] = bacc[k][j] = 0.0f;
  }
  for (int row = blockIdx.x; row < M; row += gridDim.x) {
    const T* xr = x + (size_t)row * D;
    const T* dyr = dy + (size_t)row * D;
    T* dxr = dx + (size_t)row * D;
    const float mu = mean[row], r = rstd[row];
    float xh[kPacks][8], gv[kPacks][8];
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int k = 0; k < kPacks; ++k) {
      const int i = (threadIdx.x + k * blockDim.x) * 8;
      if (i < D) {
        float xv[8], wv[8];
        Pack8<T>::load(xr + i, xv);
        Pack8<T>::load(dyr + i, gv[k]);
        if (w != nullptr) Pack8<T>::load(w + i, wv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[k][j] = (xv[j] - mu) * r;
          wacc[k][j] += gv[k][j] * xh[k][j];
          bacc[k][j] += gv[k][j];
          if (w != nullptr) gv[k][j] *= wv[j];
          s1 += gv[k][j];
          s2 += gv[k][j] * xh[k][j];
        }
      }
    }
    s1 = block_sum(s1, red) / (float)D;
    s2 = block_sum(s2, red) / (float)D;
#pragma unroll
    for (int k = 0; k < kPacks; ++k) {
      const int i = (threadIdx.x + k * blockDim.x) * 8;
      if (i < D) {
        float o[8];
        if (accumulate_dx) Pack8<T>::load(dxr + i, o);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float t = r * (gv[k][j] - s1 - xh[k][j] * s2);
          o[j] = accumulate_dx ? o[j] + t : t;
        }
        Pack8<T>::store(dxr + i, o);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kPacks; ++k) {
    const int i = (threadIdx.x + k * blockDim.x) * 8;
    if (i < D) {
      if (ws != nullptr) {
        Pack8<float>::store(ws + (size_t)blockIdx.x * 2 * D + i, wacc[k]);
        Pack8<float>::store(ws + (size_t)blockIdx.x * 2 * D + D + i, bacc[k]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (dw != nullptr) atomicAdd(dw + i + j, wacc[k][j]);
          if (db != nullptr) atomicAdd(db + i + j, bacc[k][j]);
        }
      }
    }
  }
}

// ---- LayerNorm, one WARP per row (D <= 1280: CLIP-L 1024, SigLIP 1152, DiT 384 / 768 / 1024) --------------------------
// The block-per-row kernels above put 64 threads on a 2 KB row and 8 such blocks on an SM: 8 rows in flight, two block
// barriers per row, 19 us forward / 34 us backward for [8224, 1024] = 0.26 of the copy peak.  Here a warp owns a row
// (lane l holds packs l, l + 32, ...): shuffles only, 64 rows in flight per SM.  The backward is split: this kernel
// writes dx; dw / db are a column reduction over rows (layernorm_dwdb_kernel) that re-reads x / dy out of L2.
template <typename T, int kP>
__global__ void __launch_bounds__(256, 3) layernorm_fwd_warp_kernel(const T* __restrict__ x, const T* __restrict__ w,
                                                                 const T* __restrict__ b, T* __restrict__ y,
                                                                 float* __restrict__ mean, float* __restrict__ rstd,
                                                                 int M, int D, float eps) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int n_pack = D >> 3;
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < M; row += gridDim.x * wpb) {
    const T* xr = x + (size_t)row * D;
    T* yr = y + (size_t)row * D;
    float v[kP][8];
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < kP; ++k) {
      const int pk = lane + 32 * k;
      if (pk < n_pack) {
        Pack8<T>::load(xr + pk * 8, v[k]);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[k][j];
      }
    }
    const float mu = warp_sum(s) / (float)D;
    float ss = 0.0f;
#pragma unroll
    for (int k = 0; k < kP; ++k) {
      if (lane + 32 * k < n_pack) {
#pragma unroll
        for (int j = 0; j < 8; ++j) ss += (v[k][j] - mu) * (v[k][j] - mu);
      }
    }
    const float r = rsqrtf(warp_sum(ss) / (float)D + eps);
    if (lane == 0) {
      if (mean != nullptr) mean[row] = mu;
      if (rstd != nullptr) rstd[row] = r;
    }
#pragma unroll
    for (int k = 0; k < kP; ++k) {
      const int pk = lane + 32 * k;
      if (pk < n_pack) {
        float wv[8], bv[8], o[8];
        if (w != nullptr) Pack8<T>::load(w + pk * 8, wv);
        if (b != nullptr) Pack8<T>::load(b + pk * 8, bv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float t = (v[k][j] - mu) * r;
          if (w != nullptr) t *= wv[j];
          if (b != nullptr) t += bv[j];
          o[j] = t;
        }
        Pack8<T>::store(yr + pk * 8, o);
      }
    }
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * w.  x / dy stay in registers as loaded (Raw8) between the
// statistics pass and the output pass.
template <typename T, int kP>
__global__ void __launch_bounds__(256, 2) layernorm_bwd_dx_warp_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                                    const T* __restrict__ w,
                                                                    const float* __restrict__ mean,
                                                                    const float* __restrict__ rstd, T* dx, int M, int D,
                                                                    int accumulate_dx) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int n_pack = D >> 3;
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < M; row += gridDim.x * wpb) {
    const T* xr = x + (size_t)row * D;
    const T* dyr = dy + (size_t)row * D;
    T* dxr = dx + (size_t)row * D;
    const float mu = mean[row], r = rstd[row];
    Raw8<T> rx[kP], rg[kP];
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int k = 0; k < kP; ++k) {
      const int pk = lane + 32 * k;
      if (pk < n_pack) {
        rx[k].load(xr + pk * 8);
        rg[k].load(dyr + pk * 8);
      }
    }
#pragma unroll
    for (int k = 0; k < kP; ++k) {
      const int pk = lane + 32 * k;
      if (pk < n_pack) {
        float xv[8], gv[8], wv[8];
        rx[k].unpack(xv);
        rg[k].unpack(gv);
        if (w != nullptr) Pack8<T>::load(w + pk * 8, wv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xv[j] - mu) * r;
          const float g = w != nullptr ? gv[j] * wv[j] : gv[j];
          s1 += g;
          s2 += g * xh;
        }
      }
    }
    s1 = warp_sum(s1) / (float)D;
    s2 = warp_sum(s2) / (float)D;
#pragma unroll
    for (int k = 0; k < kP; ++k) {
      const int pk = lane + 32 * k;
      if (pk < n_pack) {
        float xv[8], gv[8], wv[8], o[8];
        rx[k].unpack(xv);
        rg[k].unpack(gv);
        if (w != nullptr) Pack8<T>::load(w + pk * 8, wv);
        if (accumulate_dx) Pack8<T>::load(dxr + pk * 8, o);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xv[j] - mu) * r;
          const float g = w != nullptr ? gv[j] * wv[j] : gv[j];
          const float t = r * (g - s1 - xh * s2);
          o[j] = accumulate_dx ? o[j] + t : t;
        }
        Pack8<T>::store(dxr + pk * 8, o);
      }
    }
  }
}

// dw[col] += sum_rows dy * xhat,  db[col] += sum_rows dy.  Block = 64 column-threads (8 columns each) x 4 row groups, two
// rows of both operands in flight per thread; the row groups are combined in shared memory and one thread per column
// group issues the block's atomics (the same fp32-atomic finish as colsum: its order is not fixed).
template <typename T>
__global__ void __launch_bounds__(256) layernorm_dwdb_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                             const float* __restrict__ mean,
                                                             const float* __restrict__ rstd, float* __restrict__ dw,
                                                             float* __restrict__ db, int M, int D, int rows_per_block) {
  __shared__ float part[4][64][16];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int col8 = (blockIdx.x * 64 + tx) * 8;
  const bool live = col8 < D;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float wacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, bacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (live) {
    int r = r0 + ty;
    for (; r + 4 < r1; r += 8) {
      float x0[8], g0[8], x1[8], g1[8];
      Pack8<T>::load(x + (size_t)r * D + col8, x0);
      Pack8<T>::load(dy + (size_t)r * D + col8, g0);
      Pack8<T>::load(x + (size_t)(r + 4) * D + col8, x1);
      Pack8<T>::load(dy + (size_t)(r + 4) * D + col8, g1);
      const float m0 = mean[r], s0 = rstd[r], m1 = mean[r + 4], s1 = rstd[r + 4];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        wacc[j] += g0[j] * ((x0[j] - m0) * s0) + g1[j] * ((x1[j] - m1) * s1);
        bacc[j] += g0[j] + g1[j];
      }
    }
    for (; r < r1; r += 4) {
      float x0[8], g0[8];
      Pack8<T>::load(x + (size_t)r * D + col8, x0);
      Pack8<T>::load(dy + (size_t)r * D + col8, g0);
      const float m0 = mean[r], s0 = rstd[r];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        wacc[j] += g0[j] * ((x0[j] - m0) * s0);
        bacc[j] += g0[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    part[ty][tx][j] = wacc[j];
    part[ty][tx][8 + j] = bacc[j];
  }
  __syncthreads();
  if (ty == 0 && live) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (dw != nullptr)
        atomicAdd(dw + col8 + j, (part[0][tx][j] + part[1][tx][j]) + (part[2][tx][j] + part[3][tx][j]));
      if (db != nullptr)
        atomicAdd(db + col8 + j, (part[0][tx][8 + j] + part[1][tx][8 + j]) + (part[2][tx][8 + j] + part[3][tx][8 + j]));
    }
  }
}

// Rows wider than the register-cached kernel holds (D > 8*256*kMaxPacks; OFT's MLPResNet input LayerNorm has
// D = action_dim * hidden = 25088, oft/action_model/model.py:108,146): one block per row, two streaming passes,
// dw / db by fp32 atomics (such inputs have a few hundred rows).
template <typename T>
__global__ void __launch_bounds__(256) layernorm_bwd_wide_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                                 const T* __restrict__ w, const float* __restrict__ mean,
                                                                 const float* __restrict__ rstd, T* dx,
                                                                 float* __restrict__ dw, float* __restrict__ db, int M,
                                                                 int D, int accumulate_dx) {
  __shared__ float red[33];
  for (int row = blockIdx.x; row < M; row += gridDim.x) {
    const T* xr = x + (size_t)row * D;
    const T* dyr = dy + (size_t)row * D;
    T* dxr = dx + (size_t)row * D;
    const float mu = mean[row], r = rstd[row];
    float s1 = 0.0f, s2 = 0.0f;
    for (int i = threadIdx.x * 8; i < D; i += blockDim.x * 8) {
      float xv[8], dv[8], wv[8];
      Pack8<T>::load(xr + i, xv);
      Pack8<T>::load(dyr + i, dv);
      if (w != nullptr) Pack8<T>::load(w + i, wv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float g = dv[j] * (w != nullptr ? wv[j] : 1.0f);
        s1 += g;
        s2 += g * (xv[j] - mu) * r;
      }
    }
    s1 = block_sum(s1, red) / (float)D;
    s2 = block_sum(s2, red) / (float)D;
    for (int i = threadIdx.x * 8; i < D; i += blockDim.x * 8) {
      float xv[8], dv[8], wv[8], o[8];
      Pack8<T>::load(xr + i, xv);
      Pack8<T>::load(dyr + i, dv);
      if (w != nullptr) Pack8<T>::load(w + i, wv);
      if (accumulate_dx) Pack8<T>::load(dxr + i, o);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (xv[j] - mu) * r;
        const float g = dv[j] * (w != nullptr ? wv[j] : 1.0f);
        const float t = r * (g - s1 - xh * s2);
        o[j] = accumulate_dx ? o[j] + t : t;
        if (dw != nullptr) atomicAdd(dw + i + j, dv[j] * xh);
        if (db != nullptr) atomicAdd(db + i + j, dv[j]);
      }
      Pack8<T>::store(dxr + i, o);
    }
  }
}

// --------------------------------------------------------------------- RoPE
// In-place rotate_half RoPE on the first `n_rot_heads` heads of every row of a packed
// [M, row_stride] buffer (q heads then k heads).  cos/sin: fp32 tables [n_pos, hd/2].
template <typename T>
__global__ void rope_kernel(T* __restrict__ qkv, const int* __restrict__ pos, const float* __restrict__ cos_t,
                            const float* __restrict__ sin_t, int M, int n_rot_heads, int hd, int64_t row_stride,
                            int inverse) {
  const int half = hd >> 1;
  const int packs_per_head = half >> 3;
  const int work = n_rot_heads * packs_per_head;
  for (int row = blockIdx.x; row < M; row += gridDim.x) {
    const int p = pos[row];
    const float* cr = cos_t + (size_t)p * half;
    const float* sr = sin_t + (size_t)p * half;
    T* base = qkv + (size_t)row * row_stride;
    for (int t = threadIdx.x; t < work; t += blockDim.x) {
      const int head = t / packs_per_head;
      const int i = (t - head * packs_per_head) * 8;
      T* lo = base + head * hd + i;
      T* hi = lo + half;
      float a[8], b[8], oa[8], ob[8];
      Pack8<T>::load(lo, a);
      Pack8<T>::load(hi, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float c = round_to<T>(cr[i + j]);
        float s = round_to<T>(sr[i + j]);
        if (inverse) s = -s;
        oa[j] = a[j] * c - b[j] * s;
        ob[j] = b[j] * c + a[j] * s;
      }
      Pack8<T>::store(lo, oa);
      Pack8<T>::store(hi, ob);
    }
  }
}

// ------------------------------------------------------- activations / GLU
template <typename T>
__global__ void act_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t n8, int act) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float v[8];
    Pack8<T>::load(x + i * 8, v);
    act_fwd_n(v, act);
    Pack8<T>::store(y + i * 8, v);
  }
}
template <typename T>
__global__ void act_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, T* __restrict__ dx, int64_t n8,
                               int act) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float v[8], d[8];
    Pack8<T>::load(x + i * 8, v);
    Pack8<T>::load(dy + i * 8, d);
    act_grad_mul_n(d, v, act);
    Pack8<T>::store(dx + i * 8, d);
  }
}
template <typename T, int kAct>
__global__ void glu_fwd_kernel(const T* __restrict__ g, const T* __restrict__ u, T* __restrict__ h, int64_t n8,
                               int act) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  for (; i + stride < n8; i += 2 * stride) {      // two packs in flight per thread
    float a0[8], b0[8], a1[8], b1[8];
    Pack8<T>::load(g + i * 8, a0);
    Pack8<T>::load(u + i * 8, b0);
    Pack8<T>::load(g + (i + stride) * 8, a1);
    Pack8<T>::load(u + (i + stride) * 8, b1);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a0[j] = act_only<kAct>(a0[j], act) * b0[j];
      a1[j] = act_only<kAct>(a1[j], act) * b1[j];
    }
    Pack8<T>::store(h + i * 8, a0);
    Pack8<T>::store(h + (i + stride) * 8, a1);
  }
  for (; i < n8; i += stride) {
    float a[8], b[8];
    Pack8<T>::load(g + i * 8, a);
    Pack8<T>::load(u + i * 8, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = act_only<kAct>(a[j], act) * b[j];
    Pack8<T>::store(h + i * 8, a);
  }
}
// dg = dh*u*act'(g), du = dh*act(g); dg/du may alias g/u; optional h_out = act(g)*u
template <typename T, int kAct>
__global__ void glu_bwd_kernel(const T* dh, const T* g, const T* u, T* dg, T* du, T* h_out,
                               int64_t n8, int act) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float a[8], b[8], d[8], og[8], ou[8], oh[8];
    Pack8<T>::load(g + i * 8, a);
    Pack8<T>::load(u + i * 8, b);
    Pack8<T>::load(dh + i * 8, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float f, df;
      act_pair<kAct>(a[j], act, f, df);
      og[j] = d[j] * b[j] * df;
      ou[j] = d[j] * f;
      oh[j] = f * b[j];
    }
    Pack8<T>::store(dg + i * 8, og);
    Pack8<T>::store(du + i * 8, ou);
    if (h_out != nullptr) Pack8<T>::store(h_out + i * 8, oh);
  }
}

// ------------------------------------------------------------------ softmax
// One warp per (z, q) row.  allowed(q,k) = (!keymask || keymask[b,k]) && (!bid || bid_k[b,k] <= bid_q[b,q]).
// b = z / heads.  Scores are fp32 (already scaled); P is written as T.  Fully masked rows -> zeros.
// One warp per (z, q) row; the row is read ONCE with 128-bit loads (lane owns 4 consecutive keys per step, up to
// 8 steps = 1024 keys cached in registers; longer rows fall back to re-reading).  causal != 0: bid_* are not read,
// allowed(q,k) = k <= q, and keys beyond q are never loaded.
template <typename T>
__device__ __forceinline__ void store4(T* p, float a, float b, float c, float d);
template <>
__device__ __forceinline__ void store4<float>(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
template <>
__device__ __forceinline__ void store4<bf16>(bf16* p, float a, float b, float c, float d) {
  uint2 u;
  __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  u.x = *reinterpret_cast<uint32_t*>(&lo);
  u.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = u;
}

template <typename T, int kSoftmaxSteps>
__global__ void __launch_bounds__(256, kSoftmaxSteps <= 4 ? 6 : 3) softmax_fwd_kernel(const float* __restrict__ s, T* __restrict__ p, int64_t rows,
                                                          int Sq, int Sk, int64_t s_ld, int64_t p_ld, int heads,
                                                          const uint8_t* __restrict__ keymask,
                                                          const int* __restrict__ bid_q, const int* __restrict__ bid_k,
                                                          int causal) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const bool vec_ok = (s_ld % 4 == 0) && (p_ld % 4 == 0) && Sk <= 128 * kSoftmaxSteps;
  for (int64_t row = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (int64_t)gridDim.x * wpb) {
    const int64_t z = row / Sq;
    const int q = (int)(row - z * Sq);
    const int b = (int)(z / heads);
    const float* sr = s + row * s_ld;
    T* pr = p + row * p_ld;
    const uint8_t* km = keymask != nullptr ? keymask + (size_t)b * Sk : nullptr;
    const int* bk = (!causal && bid_k != nullptr) ? bid_k + (size_t)b * Sk : nullptr;
    const int bq = (!causal && bid_q != nullptr) ? bid_q[(size_t)b * Sq + q] : 0;
    const int limit = causal ? min(Sk, q + 1) : Sk;      // keys >= limit are masked
    if (vec_ok) {
      float v[kSoftmaxSteps][4];
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < kSoftmaxSteps; ++i) {
        const int k0 = (lane + 32 * i) * 4;
        v[i][0] = v[i][1] = v[i][2] = v[i][3] = -INFINITY;
        if (k0 < limit) {
          const float4 f = *reinterpret_cast<const float4*>(sr + k0);   // row padding up to s_ld is readable
          const float t[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = k0 + j;
            bool ok = k < limit;
            if (ok && km != nullptr) ok = km[k] != 0;
            if (ok && bk != nullptr) ok = bk[k] <= bq;
            if (ok) v[i][j] = t[j];
            mx = fmaxf(mx, v[i][j]);
          }
        }
      }
      mx = warp_max(mx);
      float sum = 0.0f;
#pragma unroll
      for (int i = 0; i < kSoftmaxSteps; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[i][j] = v[i][j] == -INFINITY ? 0.0f : __expf(v[i][j] - mx);
          sum += v[i][j];
        }
      sum = warp_sum(sum);
      const float inv = sum > 0.0f ? 1.0f / sum : 0.0f;
#pragma unroll
      for (int i = 0; i < kSoftmaxSteps; ++i) {
        const int k0 = (lane + 32 * i) * 4;
        if (k0 < Sk) store4<T>(pr + k0, v[i][0] * inv, v[i][1] * inv, v[i][2] * inv, v[i][3] * inv);
      }
      continue;
    }
    float mx = -INFINITY;
    for (int k = lane; k < limit; k += 32) {
      const bool ok = (km == nullptr || km[k]) && (bk == nullptr || bk[k] <= bq);
      if (ok) mx = fmaxf(mx, sr[k]);
    }
    mx = warp_max(mx);
    float sum = 0.0f;
    for (int k = lane; k < limit; k += 32) {
      const bool ok = (km == nullptr || km[k]) && (bk == nullptr || bk[k] <= bq);
      if (ok) sum += __expf(sr[k] - mx);
    }
    sum = warp_sum(sum);
    const float inv = sum > 0.0f ? 1.0f / sum : 0.0f;
    for (int k = lane; k < Sk; k += 32) {
      const bool ok = k < limit && (km == nullptr || km[k]) && (bk == nullptr || bk[k] <= bq);
      pr[k] = from_f<T>(ok ? __expf(sr[k] - mx) * inv : 0.0f);
    }
  }
}
// dS = scale * P * (dP - sum_k P*dP)
template <typename T>
__global__ void softmax_bwd_kernel(const T* __restrict__ p, const float* __restrict__ dp, T* __restrict__ ds,
                                   int64_t rows, int Sk, int64_t p_ld, int64_t dp_ld, int64_t ds_ld, float scale) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int64_t row = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (int64_t)gridDim.x * wpb) {
    const T* pr = p + row * p_ld;
    const float* dr = dp + row * dp_ld;
    T* or_ = ds + row * ds_ld;
    float acc = 0.0f;
    for (int k = lane; k < Sk; k += 32) acc += to_f(pr[k]) * dr[k];
    acc = warp_sum(acc);
    for (int k = lane; k < Sk; k += 32) {
      const float pv = to_f(pr[k]);
      or_[k] = from_f<T>(scale * pv * (dr[k] - acc));
    }
  }
}

// --------------------------------------------------------------- reductions
// out[col] += sum_rows x[row, col]   (bias gradients).  Block = 64 column-threads (8 columns each) x 4 row groups; every
// thread keeps 4 independent 16-byte loads in flight, the row groups are combined in shared memory and ONE thread per
// column group issues the 8 atomics of the block (a 256-row block: 4x fewer atomics per column than round 1's 64-row
// blocks, which is what the kernel was bound by).
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, float* __restrict__ out, int M, int N,
                                                     int rows_per_block) {
  __shared__ float part[4][64][8];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int col8 = (blockIdx.x * 64 + tx) * 8;
  const bool live = col8 < N;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (live) {
    int r = r0 + ty;
    for (; r + 12 < r1; r += 16) {
      float v0[8], v1[8], v2[8], v3[8];
      Pack8<T>::load(x + (size_t)r * N + col8, v0);
      Pack8<T>::load(x + (size_t)(r + 4) * N + col8, v1);
      Pack8<T>::load(x + (size_t)(r + 8) * N + col8, v2);
      Pack8<T>::load(x + (size_t)(r + 12) * N + col8, v3);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += (v0[j] + v1[j]) + (v2[j] + v3[j]);
    }
    for (; r < r1; r += 4) {
      float v[8];
      Pack8<T>::load(x + (size_t)r * N + col8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) part[ty][tx][j] = acc[j];
  __syncthreads();
  if (ty == 0 && live) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      atomicAdd(out + col8 + j, (part[0][tx][j] + part[1][tx][j]) + (part[2][tx][j] + part[3][tx][j]));
  }
}

// out[col] += sum_rows x[row*ld + col]  (fp32 partials with a leading dimension)
__global__ void colsum_strided_kernel(const float* __restrict__ x, float* __restrict__ out, int M, int N, int64_t ld,
                                      int rows_per_block) {
  const int col8 = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (col8 >= N) return;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int r = r0;
  for (; r + 3 < r1; r += 4) {          // four independent 32-byte loads in flight per thread
    float v0[8], v1[8], v2[8], v3[8];
    Pack8<float>::load(x + (size_t)r * ld + col8, v0);
    Pack8<float>::load(x + (size_t)(r + 1) * ld + col8, v1);
    Pack8<float>::load(x + (size_t)(r + 2) * ld + col8, v2);
    Pack8<float>::load(x + (size_t)(r + 3) * ld + col8, v3);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += (v0[j] + v1[j]) + (v2[j] + v3[j]);
  }
  for (; r < r1; ++r) {
    float v[8];
    Pack8<float>::load(x + (size_t)r * ld + col8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += v[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) atomicAdd(out + col8 + j, acc[j]);
}

// scratch of the grid-wide deterministic reductions (launches of one stream are ordered; grid <= 8 x SMs)
__device__ float g_partials[4096];
__device__ unsigned g_ticket;

template <typename T>
__global__ void __launch_bounds__(256) sumsq_kernel(const T* __restrict__ x, int64_t n, float* __restrict__ out) {
  __shared__ float red[33];
  float acc = 0.0f;
  const int64_t n8 = n >> 3;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float v[8];
    Pack8<T>::load(x + i * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += v[j] * v[j];
  }
  if (blockIdx.x == 0)
    for (int64_t i = (n8 << 3) + threadIdx.x; i < n; i += blockDim.x) {
      float v = to_f(x[i]);
      acc += v * v;
    }
  acc = block_sum(acc, red);
  // deterministic grid reduction: per-block partials, the last block to finish adds them up in block order
  __shared__ bool last;
  if (threadIdx.x == 0) {
    g_partials[blockIdx.x] = acc;
    __threadfence();
    last = atomicInc(&g_ticket, gridDim.x - 1) == gridDim.x - 1;   // wraps to 0: ready for the next launch
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  float tot = 0.0f;
  for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) tot += __ldcg(&g_partials[i]);
  tot = block_sum(tot, red);
  if (threadIdx.x == 0) *out += tot;
}

// mean((a-b)^2) forward: out += sum/(n) ; backward: da = 2*(a-b)/n * gscale  (db = -da)
template <typename T>
__global__ void __launch_bounds__(256) mse_fwd_kernel(const T* __restrict__ a, const T* __restrict__ b, int64_t n,
                                                      float inv_n, float* __restrict__ out) {
  __shared__ float red[33];
  float acc = 0.0f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = to_f(a[i]) - to_f(b[i]);
    acc += d * d;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(out, acc * inv_n);
}
template <typename T>
__global__ void mse_bwd_kernel(const T* __restrict__ a, const T* __restrict__ b, int64_t n, float coef,
                               const float* __restrict__ gscale, T* __restrict__ da) {
  const float g = gscale != nullptr ? *gscale : 1.0f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    da[i] = from_f<T>(coef * g * (to_f(a[i]) - to_f(b[i])));
}

// -------------------------------------------------------------------- AdamW
// fp32 master weights + fp32 moments; gradient in G (bf16 or fp32); optional bf16 shadow written.
// torch.optim.AdamW semantics (decoupled decay; bias-corrected).  *clip is the global-norm clip
// coefficient computed on device (no host sync).
template <typename G>
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const G* __restrict__ g,
                                                    float* __restrict__ m, float* __restrict__ v,
                                                    bf16* __restrict__ shadow, int64_t n, float b2, float omb1,
                                                    float omb2, float eps, float decay, float step_size,
                                                    float bc2_sqrt, const float* __restrict__ clip,
                                                    const float* __restrict__ hyper) {
  // hyper != NULL: the seven scalars come from device memory (b200_adamw_dev): a captured CUDA graph of the step is
  // replayed with new learning rates / bias corrections without re-capturing
  if (hyper != nullptr) {
    b2 = hyper[0];
    omb1 = hyper[1];
    omb2 = hyper[2];
    eps = hyper[3];
    decay = hyper[4];
    step_size = hyper[5];
    bc2_sqrt = hyper[6];
  }
  // torch.optim.AdamW arithmetic (torch/optim/adamw.py, single-tensor path) with its host scalars evaluated in
  // double exactly as Python does and rounded to fp32 once: exp_avg.lerp_(g, 1-b1); exp_avg_sq.mul_(b2).addcmul_(g, g,
  // value=1-b2); p.mul_(1 - lr*wd); denom = sqrt(exp_avg_sq) / sqrt(bc2) + eps; p.addcdiv_(exp_avg, denom, -lr/bc1)
  const float cs = clip != nullptr ? *clip : 1.0f;
  const int64_t n8 = n >> 3;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float pv[8], gv[8], mv[8], vv[8];
    Pack8<float>::load(p + i * 8, pv);
    Pack8<G>::load(g + i * 8, gv);
    Pack8<float>::load(m + i * 8, mv);
    Pack8<float>::load(v + i * 8, vv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float gg = gv[j] * cs;
      mv[j] = mv[j] + omb1 * (gg - mv[j]);
      vv[j] = b2 * vv[j] + omb2 * gg * gg;
      const float denom = sqrtf(vv[j]) / bc2_sqrt + eps;
      pv[j] = pv[j] * decay - step_size * (mv[j] / denom);
    }
    Pack8<float>::store(p + i * 8, pv);
    Pack8<float>::store(m + i * 8, mv);
    Pack8<float>::store(v + i * 8, vv);
    if (shadow != nullptr) Pack8<bf16>::store(shadow + i * 8, pv);
  }
  if (blockIdx.x == 0) {
    for (int64_t i = (n8 << 3) + threadIdx.x; i < n; i += blockDim.x) {
      const float gg = to_f(g[i]) * cs;
      const float mm = m[i] + omb1 * (gg - m[i]);
      const float vv = b2 * v[i] + omb2 * gg * gg;
      const float pp = p[i] * decay - step_size * (mm / (sqrtf(vv) / bc2_sqrt + eps));
      m[i] = mm;
      v[i] = vv;
      p[i] = pp;
      if (shadow != nullptr) shadow[i] = __float2bfloat16(pp);
    }
  }
}

// clip = min(1, max_norm / (sqrt(sumsq) + 1e-6))   (torch.nn.utils.clip_grad_norm_)
__global__ void clip_coef_kernel(const float* __restrict__ sumsq, float max_norm, float* __restrict__ clip,
                                 float* __restrict__ norm_out) {
  const float nrm = sqrtf(*sumsq);
  if (norm_out != nullptr) *norm_out = nrm;
  *clip = fminf(1.0f, max_norm / (nrm + 1e-6f));
}

template <typename S, typename D>
__global__ void cast_kernel(const S* __restrict__ src, D* __restrict__ dst, int64_t n) {
  const int64_t n8 = n >> 3;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float v[8];
    Pack8<S>::load(src + i * 8, v);
    Pack8<D>::store(dst + i * 8, v);
  }
  if (blockIdx.x == 0)
    for (int64_t i = (n8 << 3) + threadIdx.x; i < n; i += blockDim.x) dst[i] = from_f<D>(to_f(src[i]));
}

// y = a + b (optionally y may alias a)
template <typename T>
__global__ void add_kernel(const T* a, const T* __restrict__ b, T* y, int64_t n8) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float u[8], v[8];
    Pack8<T>::load(a + i * 8, u);
    Pack8<T>::load(b + i * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) u[j] += v[j];
    Pack8<T>::store(y + i * 8, u);
  }
}

static inline int grid_1d(int64_t work_items, int block) {
  int64_t g = ceil_div(work_items, block);
  int64_t cap = (int64_t)num_sms() * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace b200

using namespace b200;
#define STREAM reinterpret_cast<cudaStream_t>(stream)
#define DISPATCH_T(dtype, ...)                   \
  if ((dtype) == B200_F32) {                     \
    using T = float;                             \
    __VA_ARGS__;                                 \
  } else {                                       \
    using T = bf16;                              \
    __VA_ARGS__;                                 \
  }

// Persistent grid of the backward norm kernels: exactly the number of blocks that are resident at once (a grid sized
// for more than fit would run a second, partial wave of the grid-stride loop).  The workspace has one row per block; the
// bound below (16 blocks per SM) is what callers allocate.
// 1-D grid-stride launch sized to what is resident at once (blocks beyond that would start only after whole first-wave
// blocks finish their strided share, leaving the SMs at partial occupancy for the tail).
template <typename K>
static int grid_fit(K kernel, int64_t work_items, int block) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  int64_t g = ceil_div(work_items, block);
  const int64_t cap = (int64_t)num_sms() * per_sm;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}

template <typename K>
static int resident_grid(K kernel, int nt, int64_t rows) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, nt, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
  if (per_sm > 16) per_sm = 16;
  return grid_for_rows(rows, per_sm);
}

// SM-resident grid of the staged rmsnorm backward (3 blocks of <= 72 KB of stage ring per SM)
// bit 0: rmsnorm backward staged (default: 54 / 63 us against 65 / 87 us for the register-prefetch kernel at [9888, 3584],
// plain / accumulate); bit 1: rmsnorm forward staged (measured SLOWER, 35.8 vs 30.7 us — a 7 KB row per block and one
// barrier round trip per row leave the ring's depth unused — so it is off by default and kept for the A/B test)
// bit 2: LayerNorm forward / backward with one warp per row + a column kernel for dw / db (D <= 1280).  Parity-tested on
// B200 against fp32 torch and the block-per-row kernels (tests/test_gpu_kernels.py::test_layernorm_warp_and_block_kernels)
// but not yet timed or run under the model-level suite: opt-in until it is.
static int g_norm_staged = 1;

template <typename T>
static int launch_rmsnorm_bwd_staged(const T* dy, const T* x, const T* w, const float* rstd, T* dx, float* dw, float* ws,
                                     int64_t M, int64_t D, int unit_offset, int accumulate_dx, int nt, size_t smem,
                                     int* grid_out, cudaStream_t stream) {
  auto kernel = rmsnorm_bwd_staged_kernel<T, 2, 3>;
  B200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, nt, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  const int grid = grid_for_rows(M, per_sm > 16 ? 16 : per_sm);
  kernel<<<grid, nt, smem, stream>>>(dy, x, w, rstd, dx, dw, ws, (int)M, (int)D, unit_offset, accumulate_dx);
  *grid_out = grid;
  return 0;
}

template <typename K>
static int warp_row_grid(K kernel, int64_t M) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
  const int64_t want = ceil_div(M, 8), cap = (int64_t)num_sms() * per_sm;
  return (int)(want < cap ? want : cap);
}

extern "C" {

int b200_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd, int64_t M, int64_t D, float eps,
                     int unit_offset, int dtype, void* stream) {
  B200_CHECK(D % 8 == 0, "rmsnorm_fwd: D=%lld must be a multiple of 8", (long long)D);
  if (M == 0) return 0;
  B200_CHECK(D <= 8 * 256 * kMaxPacks, "rmsnorm_fwd: unsupported D=%lld", (long long)D);
  const int nt = norm_threads(D);
  int grid;
  const int64_t row_bytes = D * (dtype == B200_BF16 ? 2 : 4);
  const bool aligned16 = (reinterpret_cast<uintptr_t>(x) & 15) == 0 && row_bytes % 16 == 0;
  if ((g_norm_staged & 2) && D <= (int64_t)16 * nt && aligned16 && 4 * row_bytes <= 48 * 1024) {
    const size_t smem = (size_t)(4 * row_bytes);           // 4-deep ring of x rows per block
    DISPATCH_T(dtype, {
      int per_sm = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rmsnorm_fwd_staged_kernel<T, 2, 4>, nt, smem) !=
              cudaSuccess || per_sm < 1)
        per_sm = 1;
      grid = grid_for_rows(M, per_sm > 16 ? 16 : per_sm);
      rmsnorm_fwd_staged_kernel<T, 2, 4><<<grid, nt, smem, STREAM>>>((const T*)x, (const T*)w, (T*)y, rstd, (int)M,
                                                                     (int)D, eps, unit_offset);
    });
  } else if (D <= (int64_t)16 * nt) {
    DISPATCH_T(dtype, grid = resident_grid(rmsnorm_fwd_kernel<T, 2>, nt, M));
    DISPATCH_T(dtype, (rmsnorm_fwd_kernel<T, 2><<<grid, nt, 0, STREAM>>>((const T*)x, (const T*)w, (T*)y, rstd, (int)M,
                                                                          (int)D, eps, unit_offset)));
  } else {
    DISPATCH_T(dtype, grid = resident_grid(rmsnorm_fwd_kernel<T, kMaxPacks>, nt, M));
    DISPATCH_T(dtype, (rmsnorm_fwd_kernel<T, kMaxPacks><<<grid, nt, 0, STREAM>>>((const T*)x, (const T*)w, (T*)y, rstd,
                                                                                  (int)M, (int)D, eps, unit_offset)));
  }
  B200_LAUNCH_OK();
  return 0;
}

int b200_set_norm_staged(int on) {
  g_norm_staged = on;
  return 0;
}

// rows of fp32 workspace the backward norm kernels want (upper bound on the launched blocks); 0 -> atomics path
int64_t b200_norm_bwd_workspace_rows(int64_t M, int64_t D) {
  (void)D;
  return grid_for_rows(M, 16);
}

int b200_rmsnorm_bwd(const void* dy, const void* x, const void* w, const float* rstd, void* dx, float* dw,
                     float* workspace, int64_t M, int64_t D, int unit_offset, int accumulate_dx, int dtype,
                     void* stream) {
  B200_CHECK(D % 8 == 0 && D <= 8 * 256 * kMaxPacks, "rmsnorm_bwd: unsupported D=%lld", (long long)D);
  if (M == 0) return 0;
  const int nt = norm_threads(D);
  int grid;
  float* ws = dw != nullptr ? workspace : nullptr;
  // rows staged in shared memory by the bulk-copy engine (3-deep ring per block): every decoder / tower width in bf16
  const int64_t row_bytes = D * (dtype == B200_BF16 ? 2 : 4);
  const int64_t stage_smem = 3 * (accumulate_dx ? 3 : 2) * row_bytes;
  const bool aligned16 = ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) |
                           reinterpret_cast<uintptr_t>(dx)) & 15) == 0 && row_bytes % 16 == 0;
  if ((g_norm_staged & 1) && D <= (int64_t)16 * nt && aligned16 && stage_smem <= 72 * 1024) {
    int rc;
    DISPATCH_T(dtype, rc = launch_rmsnorm_bwd_staged<T>((const T*)dy, (const T*)x, (const T*)w, rstd, (T*)dx, dw, ws, M, D,
                                                        unit_offset, accumulate_dx, nt, (size_t)stage_smem, &grid,
                                                        STREAM));
    if (rc) return rc;
  } else if (D <= (int64_t)16 * nt) {
    DISPATCH_T(dtype, grid = resident_grid(rmsnorm_bwd_kernel<T, 2>, nt, M));
    DISPATCH_T(dtype, (rmsnorm_bwd_kernel<T, 2><<<grid, nt, 0, STREAM>>>(
                          (const T*)dy, (const T*)x, (const T*)w, rstd, (T*)dx, dw, ws, (int)M, (int)D, unit_offset,
                          accumulate_dx)));
  } else {
    DISPATCH_T(dtype, grid = resident_grid(rmsnorm_bwd_kernel<T, kMaxPacks>, nt, M));
    DISPATCH_T(dtype, (rmsnorm_bwd_kernel<T, kMaxPacks><<<grid, nt, 0, STREAM>>>(
                          (const T*)dy, (const T*)x, (const T*)w, rstd, (T*)dx, dw, ws, (int)M, (int)D, unit_offset,
                          accumulate_dx)));
  }
  B200_LAUNCH_OK();
  if (ws != nullptr) {   // dw[D] += column sums of the [grid, D] partials
    dim3 g2((unsigned)ceil_div(D / 8, 64), (unsigned)ceil_div(grid, 64));
    colsum_kernel<float><<<g2, 256, 0, STREAM>>>(ws, dw, grid, (int)D, 64);
    B200_LAUNCH_OK();
  }
  return 0;
}

int b200_layernorm_fwd(const void* x, const void* w, const void* b, void* y, float* mean, float* rstd, int64_t M,
                       int64_t D, float eps, int dtype, void* stream) {
  B200_CHECK(D % 8 == 0, "layernorm_fwd: D=%lld must be a multiple of 8", (long long)D);
  if (M == 0) return 0;
  if ((g_norm_staged & 4) && D <= 1280) {          // one warp per row
    const int kp = D <= 512 ? 2 : (D <= 1024 ? 4 : 5);
#define B200_LN_FWD(KP)                                                                                          \
  DISPATCH_T(dtype, (layernorm_fwd_warp_kernel<T, KP><<<warp_row_grid(layernorm_fwd_warp_kernel<T, KP>, M), 256, 0, \
                                                        STREAM>>>((const T*)x, (const T*)w, (const T*)b, (T*)y, mean, \
                                                                  rstd, (int)M, (int)D, eps)))
    if (kp == 2) { B200_LN_FWD(2); } else if (kp == 4) { B200_LN_FWD(4); } else { B200_LN_FWD(5); }
#undef B200_LN_FWD
    B200_LAUNCH_OK();
    return 0;
  }
  const int nt = norm_threads(D);
  int grid;
  DISPATCH_T(dtype, grid = resident_grid(layernorm_fwd_kernel<T>, nt, M));
  DISPATCH_T(dtype, (layernorm_fwd_kernel<T><<<grid, nt, 0, STREAM>>>(
                        (const T*)x, (const T*)w, (const T*)b, (T*)y, mean, rstd, (int)M, (int)D, eps)));
  B200_LAUNCH_OK();
  return 0;
}

int b200_layernorm_bwd(const void* dy, const void* x, const void* w, const float* mean, const float* rstd, void* dx,
                       float* dw, float* db, float* workspace, int64_t M, int64_t D, int accumulate_dx, int dtype,
                       void* stream) {
  B200_CHECK(D % 8 == 0, "layernorm_bwd: D=%lld must be a multiple of 8", (long long)D);
  if (M == 0) return 0;
  if (D > 8 * 256 * kMaxPacks) {
    DISPATCH_T(dtype, (layernorm_bwd_wide_kernel<T><<<grid_for_rows(M, 4), 256, 0, STREAM>>>(
                          (const T*)dy, (const T*)x, (const T*)w, mean, rstd, (T*)dx, dw, db, (int)M, (int)D,
                          accumulate_dx)));
    B200_LAUNCH_OK();
    return 0;
  }
  if ((g_norm_staged & 4) && D <= 1280) {          // one warp per row for dx, a column kernel for dw / db
    const int kp = D <= 512 ? 2 : (D <= 1024 ? 4 : 5);
#define B200_LN_BWD(KP)                                                                                              \
  DISPATCH_T(dtype, (layernorm_bwd_dx_warp_kernel<T, KP><<<warp_row_grid(layernorm_bwd_dx_warp_kernel<T, KP>, M), 256, \
                                                           0, STREAM>>>((const T*)dy, (const T*)x, (const T*)w, mean,   \
                                                                        rstd, (T*)dx, (int)M, (int)D, accumulate_dx)))
    if (kp == 2) { B200_LN_BWD(2); } else if (kp == 4) { B200_LN_BWD(4); } else { B200_LN_BWD(5); }
#undef B200_LN_BWD
    B200_LAUNCH_OK();
    if (dw != nullptr || db != nullptr) {
      const int rows_per_block = 64;
      dim3 g((unsigned)ceil_div(D / 8, 64), (unsigned)ceil_div(M, rows_per_block));
      DISPATCH_T(dtype, (layernorm_dwdb_kernel<T><<<g, 256, 0, STREAM>>>((const T*)dy, (const T*)x, mean, rstd, dw, db,
                                                                         (int)M, (int)D, rows_per_block)));
      B200_LAUNCH_OK();
    }
    return 0;
  }
  const int nt = norm_threads(D);
  int grid;
  // workspace rows hold [dw partial | db partial]; it is only usable when both gradients are wanted
  float* ws = (dw != nullptr && db != nullptr) ? workspace : nullptr;
  if (D <= (int64_t)16 * nt) {
    DISPATCH_T(dtype, grid = resident_grid(layernorm_bwd_kernel<T, 2>, nt, M));
    DISPATCH_T(dtype, (layernorm_bwd_kernel<T, 2><<<grid, nt, 0, STREAM>>>(
                          (const T*)dy, (const T*)x, (const T*)w, mean, rstd, (T*)dx, dw, db, ws, (int)M, (int)D,
                          accumulate_dx)));
  } else {
    DISPATCH_T(dtype, grid = resident_grid(layernorm_bwd_kernel<T, kMaxPacks>, nt, M));
    DISPATCH_T(dtype, (layernorm_bwd_kernel<T, kMaxPacks><<<grid, nt, 0, STREAM>>>(
                          (const T*)dy, (const T*)x, (const T*)w, mean, rstd, (T*)dx, dw, db, ws, (int)M, (int)D,
                          accumulate_dx)));
  }
  B200_LAUNCH_OK();
  if (ws != nullptr) {
    dim3 g2((unsigned)ceil_div(2 * D / 8, 64), (unsigned)ceil_div(grid, 64));
    // dw and db are reduced in one pass over the [grid, 2D] partials when they are adjacent; else two passes
    if (db == dw + D) {
      colsum_kernel<float><<<g2, 256, 0, STREAM>>>(ws, dw, grid, (int)(2 * D), 64);
      B200_LAUNCH_OK();
    } else {
      // 8 partial rows per block: a [~1200, 1024] workspace gives ~300 blocks (64 rows per block left 38 blocks of 64
      // threads for 10 MB: 22 us per launch, 94 launches per CogACT step)
      dim3 g1((unsigned)ceil_div(D / 8, 64), (unsigned)ceil_div(grid, 8));
      colsum_strided_kernel<<<g1, 64, 0, STREAM>>>(ws, dw, grid, (int)D, 2 * D, 8);
      B200_LAUNCH_OK();
      colsum_strided_kernel<<<g1, 64, 0, STREAM>>>(ws + D, db, grid, (int)D, 2 * D, 8);
      B200_LAUNCH_OK();
    }
  }
  return 0;
}

int b200_rope(void* qkv, const int32_t* pos, const float* cos_t, const float* sin_t, int64_t M, int n_rot_heads,
              int head_dim, int64_t row_stride, int inverse, int dtype, void* stream) {
  B200_CHECK(head_dim % 16 == 0, "rope: head_dim=%d must be a multiple of 16", head_dim);
  if (M == 0) return 0;
  int work = n_rot_heads * (head_dim / 16);
  int block = work < 32 ? 32 : (work > 512 ? 512 : ((work + 31) / 32) * 32);
  int grid;
  DISPATCH_T(dtype, grid = resident_grid(rope_kernel<T>, block, M));
  DISPATCH_T(dtype, (rope_kernel<T><<<grid, block, 0, STREAM>>>((T*)qkv, pos, cos_t, sin_t, (int)M,
                                                                            n_rot_heads, head_dim, row_stride,
                                                                            inverse)));
  B200_LAUNCH_OK();
  return 0;
}

int b200_act_fwd(const void* x, void* y, int64_t n, int act, int dtype, void* stream) {
  B200_CHECK(n % 8 == 0, "act_fwd: n must be a multiple of 8");
  if (n == 0) return 0;
  DISPATCH_T(dtype, (act_fwd_kernel<T><<<grid_fit(act_fwd_kernel<T>, n / 8, 256), 256, 0, STREAM>>>((const T*)x, (T*)y, n / 8, act)));
  B200_LAUNCH_OK();
  return 0;
}
int b200_act_bwd(const void* dy, const void* x, void* dx, int64_t n, int act, int dtype, void* stream) {
  B200_CHECK(n % 8 == 0, "act_bwd: n must be a multiple of 8");
  if (n == 0) return 0;
  DISPATCH_T(dtype, (act_bwd_kernel<T><<<grid_fit(act_bwd_kernel<T>, n / 8, 256), 256, 0, STREAM>>>((const T*)dy, (const T*)x, (T*)dx,
                                                                                n / 8, act)));
  B200_LAUNCH_OK();
  return 0;
}
int b200_glu_fwd(const void* g, const void* u, void* h, int64_t n, int act, int dtype, void* stream) {
  B200_CHECK(n % 8 == 0, "glu_fwd: n must be a multiple of 8");
  if (n == 0) return 0;
  if (dtype == B200_BF16 && act == B200_ACT_SILU)
    glu_fwd_kernel<bf16, 4><<<grid_fit(glu_fwd_kernel<bf16, 4>, n / 8, 256), 256, 0, STREAM>>>(
        (const bf16*)g, (const bf16*)u, (bf16*)h, n / 8, act);
  else if (dtype == B200_BF16 && act == B200_ACT_GELU_TANH)
    glu_fwd_kernel<bf16, 2><<<grid_fit(glu_fwd_kernel<bf16, 2>, n / 8, 256), 256, 0, STREAM>>>(
        (const bf16*)g, (const bf16*)u, (bf16*)h, n / 8, act);
  else
    DISPATCH_T(dtype, (glu_fwd_kernel<T, -1><<<grid_fit(glu_fwd_kernel<T, -1>, n / 8, 256), 256, 0, STREAM>>>(
                          (const T*)g, (const T*)u, (T*)h, n / 8, act)));
  B200_LAUNCH_OK();
  return 0;
}
int b200_glu_bwd(const void* dh, const void* g, const void* u, void* dg, void* du, void* h_out, int64_t n, int act,
                 int dtype, void* stream) {
  B200_CHECK(n % 8 == 0, "glu_bwd: n must be a multiple of 8");
  if (n == 0) return 0;
  if (dtype == B200_BF16 && act == B200_ACT_SILU)
    glu_bwd_kernel<bf16, 4><<<grid_fit(glu_bwd_kernel<bf16, 4>, n / 8, 256), 256, 0, STREAM>>>(
        (const bf16*)dh, (const bf16*)g, (const bf16*)u, (bf16*)dg, (bf16*)du, (bf16*)h_out, n / 8, act);
  else if (dtype == B200_BF16 && act == B200_ACT_GELU_TANH)
    glu_bwd_kernel<bf16, 2><<<grid_fit(glu_bwd_kernel<bf16, 2>, n / 8, 256), 256, 0, STREAM>>>(
        (const bf16*)dh, (const bf16*)g, (const bf16*)u, (bf16*)dg, (bf16*)du, (bf16*)h_out, n / 8, act);
  else
    DISPATCH_T(dtype, (glu_bwd_kernel<T, -1><<<grid_fit(glu_bwd_kernel<T, -1>, n / 8, 256), 256, 0, STREAM>>>(
                          (const T*)dh, (const T*)g, (const T*)u, (T*)dg, (T*)du, (T*)h_out, n / 8, act)));
  B200_LAUNCH_OK();
  return 0;
}

int b200_softmax_fwd(const float* scores, void* p, int64_t Z, int64_t Sq, int64_t Sk, int64_t s_ld, int64_t p_ld,
                     int heads, const uint8_t* keymask, const int32_t* bid_q, const int32_t* bid_k, int causal,
                     int p_dtype, void* stream) {
  const int64_t rows = Z * Sq;
  if (rows == 0) return 0;
  B200_CHECK((bid_q == nullptr) == (bid_k == nullptr), "softmax_fwd: bid_q and bid_k go together");
  const int dtype = p_dtype;
  const int steps = (int)ceil_div(Sk, 128);
#define SMX(T_, N_)                                                                                              \
  softmax_fwd_kernel<T_, N_><<<grid_1d(ceil_div(rows, 8), 1), 256, 0, STREAM>>>(                                  \
      scores, (T_*)p, rows, (int)Sq, (int)Sk, s_ld, p_ld, heads > 0 ? heads : 1, keymask, bid_q, bid_k, causal)
  if (dtype == B200_F32) {
    if (steps <= 1) SMX(float, 1); else if (steps <= 2) SMX(float, 2); else if (steps <= 4) SMX(float, 4); else SMX(float, 8);
  } else {
    if (steps <= 1) SMX(bf16, 1); else if (steps <= 2) SMX(bf16, 2); else if (steps <= 3) SMX(bf16, 3);
    else if (steps <= 4) SMX(bf16, 4); else SMX(bf16, 8);
  }
#undef SMX
  B200_LAUNCH_OK();
  return 0;
}
int b200_softmax_bwd(const void* p, const float* dp, void* ds, int64_t rows, int64_t Sk, int64_t p_ld, int64_t dp_ld,
                     int64_t ds_ld, float scale, int p_dtype, void* stream) {
  if (rows == 0) return 0;
  const int dtype = p_dtype;
  DISPATCH_T(dtype, (softmax_bwd_kernel<T><<<grid_1d(rows, 8), 256, 0, STREAM>>>((const T*)p, dp, (T*)ds, rows,
                                                                                 (int)Sk, p_ld, dp_ld, ds_ld, scale)));
  B200_LAUNCH_OK();
  return 0;
}

int b200_colsum(const void* x, float* out, int64_t M, int64_t N, int dtype, void* stream) {
  B200_CHECK(N % 8 == 0, "colsum: N must be a multiple of 8");
  if (M == 0) return 0;
  // enough row blocks to fill the SMs a few times over, as few as possible beyond that (atomics per column)
  const int64_t col_blocks = ceil_div(N / 8, 64);
  int rows_per_block = 256;
  // (128-row blocks on [9888, 4608] — 702 blocks instead of 351 — measured the same 28.7 us and double the atomics whose
  // order is the step's only non-determinism: not adopted)
  while (rows_per_block > 64 && col_blocks * ceil_div(M, rows_per_block) < 2 * num_sms()) rows_per_block >>= 1;
  dim3 grid((unsigned)col_blocks, (unsigned)ceil_div(M, rows_per_block));
  DISPATCH_T(dtype, (colsum_kernel<T><<<grid, 256, 0, STREAM>>>((const T*)x, out, (int)M, (int)N, rows_per_block)));
  B200_LAUNCH_OK();
  return 0;
}

int b200_sumsq(const void* x, int64_t n, float* out, int dtype, void* stream) {
  if (n == 0) return 0;
  DISPATCH_T(dtype, (sumsq_kernel<T><<<grid_1d(n / 8 + 1, 256), 256, 0, STREAM>>>((const T*)x, n, out)));
  B200_LAUNCH_OK();
  return 0;
}

int b200_mse_fwd(const void* a, const void* b, int64_t n, float* out, int dtype, void* stream) {
  if (n == 0) return 0;
  // action-chunk losses are a few 10^4 elements: one block = a fixed summation order (run-to-run identical losses)
  DISPATCH_T(dtype, (mse_fwd_kernel<T><<<n <= (1 << 22) ? 1 : grid_1d(n, 256), 256, 0, STREAM>>>((const T*)a, (const T*)b, n,
                                                                            1.0f / (float)n, out)));
  B200_LAUNCH_OK();
  return 0;
}
int b200_mse_bwd(const void* a, const void* b, int64_t n, const float* gscale, void* da, int dtype, void* stream) {
  if (n == 0) return 0;
  DISPATCH_T(dtype, (mse_bwd_kernel<T><<<grid_1d(n, 256), 256, 0, STREAM>>>((const T*)a, (const T*)b, n,
                                                                            2.0f / (float)n, gscale, (T*)da)));
  B200_LAUNCH_OK();
  return 0;
}

int b200_adamw(float* p, const void* g, float* m, float* v, void* shadow_bf16, int64_t n, double lr, double beta1,
               double beta2, double eps, double weight_decay, int64_t step, const float* clip, int g_dtype,
               void* stream) {
  if (n == 0) return 0;
  B200_CHECK(step >= 1, "adamw: step must be >= 1");
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  const float b2 = (float)beta2, omb1 = (float)(1.0 - beta1), omb2 = (float)(1.0 - beta2), e = (float)eps;
  const float decay = (float)(1.0 - lr * weight_decay), step_size = (float)(lr / bc1), bc2_sqrt = (float)sqrt(bc2);
  if (g_dtype == B200_F32)
    adamw_kernel<float><<<grid_1d(n / 8 + 1, 256), 256, 0, STREAM>>>(p, (const float*)g, m, v, (bf16*)shadow_bf16, n, b2,
                                                                     omb1, omb2, e, decay, step_size, bc2_sqrt, clip,
                                                                     nullptr);
  else
    adamw_kernel<bf16><<<grid_1d(n / 8 + 1, 256), 256, 0, STREAM>>>(p, (const bf16*)g, m, v, (bf16*)shadow_bf16, n, b2,
                                                                    omb1, omb2, e, decay, step_size, bc2_sqrt, clip,
                                                                    nullptr);
  B200_LAUNCH_OK();
  return 0;
}

int b200_adamw_hyper(double lr, double beta1, double beta2, double eps, double weight_decay, int64_t step, float* out8) {
  // the seven fp32 scalars of one AdamW update exactly as b200_adamw derives them (double arithmetic, rounded once);
  // step == 0 yields the identity update (weights, moments unchanged)
  if (step <= 0) {
    const float id[8] = {1.0f, 0.0f, 0.0f, 1.0f, 1.0f, 0.0f, 1.0f, 0.0f};
    for (int i = 0; i < 8; ++i) out8[i] = id[i];
    return 0;
  }
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  out8[0] = (float)beta2;
  out8[1] = (float)(1.0 - beta1);
  out8[2] = (float)(1.0 - beta2);
  out8[3] = (float)eps;
  out8[4] = (float)(1.0 - lr * weight_decay);
  out8[5] = (float)(lr / bc1);
  out8[6] = (float)sqrt(bc2);
  out8[7] = 0.0f;
  return 0;
}

int b200_adamw_dev(float* p, const void* g, float* m, float* v, void* shadow_bf16, int64_t n, const float* hyper8,
                   const float* clip, int g_dtype, void* stream) {
  if (n == 0) return 0;
  B200_CHECK(hyper8 != nullptr, "adamw_dev: hyper8 (device pointer to the 8-float scalar block) is required");
  if (g_dtype == B200_F32)
    adamw_kernel<float><<<grid_1d(n / 8 + 1, 256), 256, 0, STREAM>>>(p, (const float*)g, m, v, (bf16*)shadow_bf16, n, 0.f,
                                                                     0.f, 0.f, 0.f, 0.f, 0.f, 1.f, clip, hyper8);
  else
    adamw_kernel<bf16><<<grid_1d(n / 8 + 1, 256), 256, 0, STREAM>>>(p, (const bf16*)g, m, v, (bf16*)shadow_bf16, n, 0.f,
                                                                    0.f, 0.f, 0.f, 0.f, 0.f, 1.f, clip, hyper8);
  B200_LAUNCH_OK();
  return 0;
}

int b200_clip_coef(const float* sumsq, float max_norm, float* clip, float* norm_out, void* stream) {
  clip_coef_kernel<<<1, 1, 0, STREAM>>>(sumsq, max_norm, clip, norm_out);
  B200_LAUNCH_OK();
  return 0;
}

int b200_cast(const void* src, void* dst, int64_t n, int src_dtype, int dst_dtype, void* stream) {
  if (n == 0) return 0;
  const int64_t w8 = n / 8 + 1;
  if (src_dtype == B200_F32 && dst_dtype == B200_BF16)
    cast_kernel<float, bf16><<<grid_fit(cast_kernel<float, bf16>, w8, 256), 256, 0, STREAM>>>((const float*)src,
                                                                                             (bf16*)dst, n);
  else if (src_dtype == B200_BF16 && dst_dtype == B200_F32)
    cast_kernel<bf16, float><<<grid_fit(cast_kernel<bf16, float>, w8, 256), 256, 0, STREAM>>>((const bf16*)src,
                                                                                             (float*)dst, n);
  else if (src_dtype == B200_F32)
    cast_kernel<float, float><<<grid_fit(cast_kernel<float, float>, w8, 256), 256, 0, STREAM>>>((const float*)src,
                                                                                               (float*)dst, n);
  else
    cast_kernel<bf16, bf16><<<grid_fit(cast_kernel<bf16, bf16>, w8, 256), 256, 0, STREAM>>>((const bf16*)src,
                                                                                           (bf16*)dst, n);
  B200_LAUNCH_OK();
  return 0;
}

int b200_add(const void* a, const void* b, void* y, int64_t n, int dtype, void* stream) {
  B200_CHECK(n % 8 == 0, "add: n must be a multiple of 8");
  if (n == 0) return 0;
  DISPATCH_T(dtype, (add_kernel<T><<<grid_fit(add_kernel<T>, n / 8, 256), 256, 0, STREAM>>>((const T*)a, (const T*)b, (T*)y, n / 8)));
  B200_LAUNCH_OK();
  return 0;
}

}  // extern "C"
