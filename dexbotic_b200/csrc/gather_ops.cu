// Index-driven kernels of the VLA hot path: image-token splice (plan / gather / scatter), row
// gathers, diffusion noising, timestep embedding, the OFT discrete action tokenizer (integer,
// bit-exact), cross entropy, and a SIMT GEMM for shapes TMA cannot describe.
// Reference call sites: include/dexbotic_b200_ops.h.
#include <limits.h>

#include "../../include/dexbotic_b200_ops.h"
#include "common.h"
#include "vec.cuh"

namespace b200 {

using bf16 = __nv_bfloat16;
constexpr int kImageTokenIndex = -200;  // dexbotic/constants.py IMAGE_TOKEN_INDEX
constexpr int64_t kIgnoreIndex = -100;  // dexbotic/constants.py IGNORE_INDEX
constexpr int kPadSrc = INT_MIN;

static inline int grid_cap(int64_t want, int per_sm = 8) {
  int64_t cap = (int64_t)num_sms() * per_sm;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}

// ------------------------------------------------------------------- splice
// new length of sample b after mask-compaction, image expansion and truncation
__global__ void splice_lengths_kernel(const int64_t* __restrict__ ids, const uint8_t* __restrict__ mask, int B, int L,
                                      int P, int64_t max_len, int32_t* __restrict__ lengths) {
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
    int64_t n = 0;
    for (int i = 0; i < L; ++i) {
      if (mask != nullptr && !mask[(size_t)b * L + i]) continue;
      n += ids[(size_t)b * L + i] == kImageTokenIndex ? P : 1;
    }
    if (max_len > 0 && n > max_len) n = max_len;
    lengths[b] = (int32_t)n;
  }
}

// single block; thread b walks sample b (L is a few hundred tokens, B a few hundred samples)
__global__ void splice_plan_kernel(const int64_t* __restrict__ ids, const uint8_t* __restrict__ mask,
                                   const int64_t* __restrict__ labels, int B, int L, int P, int64_t max_len, int S,
                                   int left_pad, int32_t* __restrict__ src, int64_t* __restrict__ new_labels,
                                   uint8_t* __restrict__ new_mask, int32_t* __restrict__ pos) {
  extern __shared__ int sh[];  // [B] image entries consumed per sample, then exclusive prefix
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    int n_img = 0;
    for (int i = 0; i < L; ++i) {
      if (mask != nullptr && !mask[(size_t)b * L + i]) continue;
      n_img += ids[(size_t)b * L + i] == kImageTokenIndex;
    }
    sh[b] = n_img > 0 ? n_img : 1;  // an image-less sample still consumes one (unused) entry: dexbotic_arch.py:264-272
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int b = 0; b < B; ++b) {
      int c = sh[b];
      sh[b] = run;
      run += c;
    }
  }
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    // pass 1: length
    int64_t n = 0;
    for (int i = 0; i < L; ++i) {
      if (mask != nullptr && !mask[(size_t)b * L + i]) continue;
      n += ids[(size_t)b * L + i] == kImageTokenIndex ? P : 1;
    }
    if (max_len > 0 && n > max_len) n = max_len;
    const int len = (int)(n < S ? n : S);
    const int off = left_pad ? S - len : 0;
    int32_t* srow = src + (size_t)b * S;
    int64_t* lrow = new_labels + (size_t)b * S;
    uint8_t* mrow = new_mask + (size_t)b * S;
    int32_t* prow = pos + (size_t)b * S;
    for (int s = 0; s < S; ++s) {
      srow[s] = kPadSrc;
      lrow[s] = kIgnoreIndex;
      mrow[s] = 0;
      prow[s] = 0;
    }
    int img_entry = sh[b];
    int j = 0;
    for (int i = 0; i < L && j < len; ++i) {
      if (mask != nullptr && !mask[(size_t)b * L + i]) continue;
      const int64_t id = ids[(size_t)b * L + i];
      if (id == kImageTokenIndex) {
        for (int t = 0; t < P && j < len; ++t, ++j) {
          srow[off + j] = -1 - (img_entry * P + t);
          mrow[off + j] = 1;
          prow[off + j] = j;
        }
        ++img_entry;
      } else {
        srow[off + j] = (int32_t)id;
        lrow[off + j] = labels != nullptr ? labels[(size_t)b * L + i] : kIgnoreIndex;
        mrow[off + j] = 1;
        prow[off + j] = j;
        ++j;
      }
    }
  }
}

template <typename T>
__global__ void splice_gather_kernel(const int32_t* __restrict__ src, const T* __restrict__ table,
                                     const T* __restrict__ feats, T* __restrict__ out, int64_t rows, int D8) {
  const int64_t total = rows * D8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / D8;
    const int c = (int)(i - r * D8) * 8;
    const int s = src[r];
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (s >= 0)
      Pack8<T>::load(table + (size_t)s * D8 * 8 + c, v);
    else if (s != kPadSrc)
      Pack8<T>::load(feats + (size_t)(-1 - s) * D8 * 8 + c, v);
    Pack8<T>::store(out + r * D8 * 8 + c, v);
  }
}

// Image rows: plain copy.  Token rows: one block per row; the FIRST occurrence of a token id sums the
// gradient rows of all its occurrences in fp32 (ascending row order) and adds the result to d_table once
// (deterministic).  The later occurrences are found by the whole block in one strided pass over `src` and sorted in
// shared memory; the first version let every thread walk all `rows` entries per 8-column chunk: 130 us per token row
// that is a first occurrence, 2.2 ms per CogACT-7B step for a 71 MB scatter.
constexpr int kMaxDupRows = 1024;
template <typename T>
__global__ void __launch_bounds__(256) splice_scatter_kernel(const int32_t* __restrict__ src,
                                                             const T* __restrict__ dout, T* __restrict__ d_table,
                                                             T* __restrict__ d_feats, int rows, int D) {
  __shared__ int found[kMaxDupRows];
  __shared__ int sorted[kMaxDupRows];
  __shared__ int n_found;
  const int r = blockIdx.x;
  const int s = src[r];
  if (s == kPadSrc) return;
  const T* drow = dout + (size_t)r * D;
  if (s < 0) {
    if (d_feats == nullptr) return;
    T* o = d_feats + (size_t)(-1 - s) * D;
    for (int c = threadIdx.x * 8; c < D; c += blockDim.x * 8) {
      float v[8];
      Pack8<T>::load(drow + c, v);
      Pack8<T>::store(o + c, v);
    }
    return;
  }
  if (d_table == nullptr) return;
  int dup = 0;
  for (int i = threadIdx.x; i < r; i += blockDim.x) dup |= (src[i] == s);
  if (threadIdx.x == 0) n_found = 0;
  if (__syncthreads_or(dup)) return;
  for (int i = r + 1 + threadIdx.x; i < rows; i += blockDim.x) {
    if (src[i] == s) {
      const int k = atomicAdd(&n_found, 1);
      if (k < kMaxDupRows) found[k] = i;
    }
  }
  __syncthreads();
  const int n = n_found;
  if (n <= kMaxDupRows) {
    // rank sort (row indices are distinct): ascending order = the summation order of the serial walk
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
      const int v = found[t];
      int rank = 0;
      for (int j = 0; j < n; ++j) rank += found[j] < v;
      sorted[rank] = v;
    }
    __syncthreads();
  }
  for (int c = threadIdx.x * 8; c < D; c += blockDim.x * 8) {
    float acc[8];
    Pack8<T>::load(drow + c, acc);
    if (n <= kMaxDupRows) {
      for (int k = 0; k < n; ++k) {
        float v[8];
        Pack8<T>::load(dout + (size_t)sorted[k] * D + c, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[j];
      }
    } else {          // more repeats of one token than the list holds: serial walk
      for (int i = r + 1; i < rows; ++i) {
        if (src[i] == s) {
          float v[8];
          Pack8<T>::load(dout + (size_t)i * D + c, v);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += v[j];
        }
      }
    }
    T* o = d_table + (size_t)s * D + c;
    float old[8];
    Pack8<T>::load(o, old);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += old[j];
    Pack8<T>::store(o, acc);
  }
}

// ------------------------------------------------------------- row gathers
template <typename T>
__global__ void gather_rows_kernel(const T* __restrict__ x, const int32_t* __restrict__ idx, T* __restrict__ out,
                                   int64_t n_idx, int D8) {
  const int64_t total = n_idx * D8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / D8;
    const int c = (int)(i - r * D8) * 8;
    float v[8];
    Pack8<T>::load(x + (size_t)idx[r] * D8 * 8 + c, v);
    Pack8<T>::store(out + r * D8 * 8 + c, v);
  }
}
// indices may repeat (CogACT repeats the cognition token 4x): serialise over i inside one thread per column pack
template <typename T>
__global__ void scatter_rows_add_kernel(const T* __restrict__ dout, const int32_t* __restrict__ idx, T* __restrict__ dx,
                                        int64_t n_idx, int D8) {
  for (int c8 = blockIdx.x * blockDim.x + threadIdx.x; c8 < D8; c8 += gridDim.x * blockDim.x) {
    for (int64_t r = 0; r < n_idx; ++r) {
      float a[8], b[8];
      T* o = dx + (size_t)idx[r] * D8 * 8 + c8 * 8;
      Pack8<T>::load(o, a);
      Pack8<T>::load(dout + r * D8 * 8 + c8 * 8, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += b[j];
      Pack8<T>::store(o, a);
    }
  }
}
// reference: argmax(cumsum(mask)==max) = first position where the running count reaches its maximum
// = index of the last non-zero mask entry (cogact_arch.py:112-117); all-zero row -> 0.
__global__ void last_valid_index_kernel(const uint8_t* __restrict__ mask, int B, int S, int32_t* __restrict__ idx) {
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
    int last = 0;
    for (int s = 0; s < S; ++s)
      if (mask[(size_t)b * S + s]) last = s;
    idx[b] = b * S + last;
  }
}

// ------------------------------------------------------- diffusion helpers
template <typename T>
__global__ void q_sample_kernel(const T* __restrict__ x, const T* __restrict__ noise, const int32_t* __restrict__ t,
                                const float* __restrict__ sa, const float* __restrict__ sb, T* __restrict__ xt,
                                int64_t total, int per) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int tt = t[i / per];
    xt[i] = from_f<T>(sa[tt] * to_f(x[i]) + sb[tt] * to_f(noise[i]));
  }
}
template <typename T>
__global__ void timestep_embedding_kernel(const float* __restrict__ t, T* __restrict__ out, int B, int dim,
                                          float max_period) {
  const int half = dim / 2;
  const int64_t total = (int64_t)B * dim;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / dim), c = (int)(i - (int64_t)b * dim);
    float v = 0.0f;
    if (c < 2 * half) {
      const int k = c < half ? c : c - half;
      const float f = expf(-logf(max_period) * (float)k / (float)half);
      const float a = t[b] * f;
      v = c < half ? cosf(a) : sinf(a);
    }
    out[i] = from_f<T>(v);
  }
}

// --------------------------------------------- OFT discrete action tokenizer
__global__ void discretize_kernel(const float* __restrict__ a, int64_t n, float scale, int64_t* __restrict__ bins) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float x = fminf(fmaxf(a[i], -1.0f), 1.0f);
    // ((x + 1) / 2 * (n_bins-1)).round(): three separately rounded fp32 ops, then round-half-even
    float y = __fmul_rn(__fmul_rn(__fadd_rn(x, 1.0f), 0.5f), scale);
    bins[i] = (int64_t)rintf(y);
  }
}
__global__ void bins_to_cont_kernel(const int64_t* __restrict__ bins, int64_t n, float scale, float* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __fsub_rn(__fmul_rn(__fdiv_rn((float)bins[i], scale), 2.0f), 1.0f);
}
// first maximum wins (torch.argmax)
template <typename T>
__global__ void __launch_bounds__(256) argmax_last_kernel(const T* __restrict__ logits, int64_t V, int n_last,
                                                          int64_t* __restrict__ idx) {
  __shared__ float sv[256];
  __shared__ int si[256];
  const T* row = logits + (size_t)blockIdx.x * V + (V - n_last);
  float best = -INFINITY;
  int bi = INT_MAX;
  for (int j = threadIdx.x; j < n_last; j += blockDim.x) {
    const float v = to_f(row[j]);
    if (v > best || (v == best && j < bi) || bi == INT_MAX) {
      best = v;
      bi = j;
    }
  }
  sv[threadIdx.x] = best;
  si[threadIdx.x] = bi;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const float v = sv[threadIdx.x + o];
      const int j = si[threadIdx.x + o];
      if (j != INT_MAX && (si[threadIdx.x] == INT_MAX || v > sv[threadIdx.x] ||
                           (v == sv[threadIdx.x] && j < si[threadIdx.x]))) {
        sv[threadIdx.x] = v;
        si[threadIdx.x] = j;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) idx[blockIdx.x] = si[0] == INT_MAX ? 0 : si[0];
}

// Temperature sampling over the last n_last logits of a row (OFTDiscreteForCausalLM.generate_action,
// oft_discrete_arch.py:264-270: softmax(logits[..., -n_last:] / T) then one multinomial draw).  The draw is the inverse
// CDF at the caller's uniform u[row] in [0, 1): the smallest j with sum_{i<=j} p_i > u * sum_i p_i, in fp32 with a
// fixed summation order (one warp per row, lane-strided partial sums + shuffle scan) — reproducible given u.
template <typename T>
__global__ void __launch_bounds__(32) sample_last_kernel(const T* __restrict__ logits, int64_t V, int n_last,
                                                         float inv_temp, const float* __restrict__ u,
                                                         int64_t* __restrict__ idx) {
  const int lane = threadIdx.x;
  const T* row = logits + (size_t)blockIdx.x * V + (V - n_last);
  const int per = (n_last + 31) / 32;                  // consecutive entries per lane: lane l owns [l*per, (l+1)*per)
  float mx = -INFINITY;
  for (int j = lane * per; j < min(n_last, (lane + 1) * per); ++j) mx = fmaxf(mx, to_f(row[j]) * inv_temp);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float part = 0.0f;
  for (int j = lane * per; j < min(n_last, (lane + 1) * per); ++j) part += expf(to_f(row[j]) * inv_temp - mx);
  float incl = part;                                   // inclusive scan over lanes
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  const float total = __shfl_sync(0xffffffffu, incl, 31);
  const float target = u[blockIdx.x] * total;
  const float before = incl - part;
  int found = INT_MAX;
  if (target < incl && target >= before) {             // the draw falls into this lane's run
    float c = before;
    for (int j = lane * per; j < min(n_last, (lane + 1) * per); ++j) {
      c += expf(to_f(row[j]) * inv_temp - mx);
      if (target < c) {
        found = j;
        break;
      }
    }
    if (found == INT_MAX) found = min(n_last, (lane + 1) * per) - 1;   // rounding at the run's end
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) found = min(found, __shfl_xor_sync(0xffffffffu, found, o));
  if (lane == 0) idx[blockIdx.x] = found == INT_MAX ? n_last - 1 : found;
}

// ------------------------------------------------------------ cross entropy
template <typename T>
__global__ void __launch_bounds__(256) ce_fwd_kernel(const T* __restrict__ logits, const int64_t* __restrict__ labels,
                                                     int64_t V, float* __restrict__ lse, float* __restrict__ loss_sum,
                                                     int32_t* __restrict__ n_valid) {
  __shared__ float red[33];
  const int64_t r = blockIdx.x;
  const T* row = logits + (size_t)r * V;
  float mx = -INFINITY;
  for (int64_t j = threadIdx.x; j < V; j += blockDim.x) mx = fmaxf(mx, to_f(row[j]));
  // block max via sum trick is not available; reduce with shuffles + smem
  mx = warp_max(mx);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -INFINITY;
    t = warp_max(t);
    if (threadIdx.x == 0) red[32] = t;
  }
  __syncthreads();
  mx = red[32];
  float s = 0.0f;
  for (int64_t j = threadIdx.x; j < V; j += blockDim.x) s += expf(to_f(row[j]) - mx);
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    const float l = mx + logf(s);
    if (lse != nullptr) lse[r] = l;
    const int64_t lab = labels[r];
    if (lab != kIgnoreIndex && lab >= 0 && lab < V) {
      atomicAdd(loss_sum, l - to_f(row[lab]));
      atomicAdd(n_valid, 1);
    }
  }
}
template <typename T>
__global__ void __launch_bounds__(256) ce_bwd_kernel(const T* __restrict__ logits, const int64_t* __restrict__ labels,
                                                     const float* __restrict__ lse, const int32_t* __restrict__ n_valid,
                                                     const float* __restrict__ gscale, T* __restrict__ dlogits,
                                                     int64_t V) {
  const int64_t r = blockIdx.x;
  const T* row = logits + (size_t)r * V;
  T* drow = dlogits + (size_t)r * V;
  const int64_t lab = labels[r];
  const bool valid = lab != kIgnoreIndex && lab >= 0 && lab < V;
  const int nv = *n_valid;
  const float g = valid ? (gscale != nullptr ? *gscale : 1.0f) / (float)(nv > 0 ? nv : 1) : 0.0f;
  const float l = lse[r];
  for (int64_t j = threadIdx.x; j < V; j += blockDim.x) {
    float p = valid ? expf(to_f(row[j]) - l) : 0.0f;
    if (j == lab) p -= 1.0f;
    drow[j] = from_f<T>(p * g);
  }
}

// ---------------------------------------------------------------- SIMT GEMM
template <typename TA, typename TD>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const TA* __restrict__ A, const TA* __restrict__ B,
                                                        TD* __restrict__ D, int M, int N, int K, int64_t lda,
                                                        int64_t ldb, int64_t ldd, int a_mn, int b_mn, float alpha,
                                                        const void* __restrict__ bias, int bias_fp32,
                                                        const void* res, int res_fp32, int64_t ldr, int act) {
  __shared__ float sa[16][17], sb[16][17];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m = blockIdx.y * 16 + ty, n = blockIdx.x * 16 + tx;
  float acc = 0.0f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    {  // A tile: sa[ty][tx] = A[m0+ty, k0+tx]
      const int mm = blockIdx.y * 16 + ty, kk = k0 + tx;
      sa[ty][tx] = (mm < M && kk < K) ? to_f(a_mn ? A[(size_t)kk * lda + mm] : A[(size_t)mm * lda + kk]) : 0.0f;
      const int nn = blockIdx.x * 16 + ty;  // sb[ty][tx] = B[n0+ty, k0+tx]
      sb[ty][tx] = (nn < N && kk < K) ? to_f(b_mn ? B[(size_t)kk * ldb + nn] : B[(size_t)nn * ldb + kk]) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc += sa[ty][k] * sb[tx][k];
    __syncthreads();
  }
  if (m < M && n < N) {
    float v = acc * alpha;
    if (bias != nullptr) v += bias_fp32 ? ((const float*)bias)[n] : __bfloat162float(((const bf16*)bias)[n]);
    v = act_fwd(v, act);
    if (res != nullptr)
      v += res_fp32 ? ((const float*)res)[(size_t)m * ldr + n] : __bfloat162float(((const bf16*)res)[(size_t)m * ldr + n]);
    D[(size_t)m * ldd + n] = from_f<TD>(v);
  }
}

}  // namespace b200

using namespace b200;
#define STREAM reinterpret_cast<cudaStream_t>(stream)
#define DISPATCH_T(dtype, ...) \
  if ((dtype) == B200_F32) {   \
    using T = float;           \
    __VA_ARGS__;               \
  } else {                     \
    using T = bf16;            \
    __VA_ARGS__;               \
  }

extern "C" {

int b200_splice_lengths(const int64_t* input_ids, const uint8_t* attention_mask, int64_t B, int64_t L,
                        int n_img_tokens, int64_t max_len, int32_t* lengths, void* stream) {
  if (B == 0) return 0;
  splice_lengths_kernel<<<(unsigned)ceil_div(B, 128), 128, 0, STREAM>>>(input_ids, attention_mask, (int)B, (int)L,
                                                                        n_img_tokens, max_len, lengths);
  B200_LAUNCH_OK();
  return 0;
}

int b200_splice_plan(const int64_t* input_ids, const uint8_t* attention_mask, const int64_t* labels, int64_t B,
                     int64_t L, int n_img_tokens, int64_t max_len, int64_t S, int left_pad, int32_t* src,
                     int64_t* new_labels, uint8_t* new_mask, int32_t* pos, void* stream) {
  if (B == 0) return 0;
  B200_CHECK(B <= 8192, "splice_plan: B=%lld too large", (long long)B);
  const int block = B < 32 ? 32 : (B > 256 ? 256 : (int)((B + 31) / 32 * 32));
  splice_plan_kernel<<<1, block, B * sizeof(int), STREAM>>>(input_ids, attention_mask, labels, (int)B, (int)L,
                                                            n_img_tokens, max_len, (int)S, left_pad, src, new_labels,
                                                            new_mask, pos);
  B200_LAUNCH_OK();
  return 0;
}

int b200_splice_gather(const int32_t* src, const void* table, const void* feats, void* out, int64_t rows, int64_t D,
                       int dtype, void* stream) {
  B200_CHECK(D % 8 == 0, "splice_gather: D must be a multiple of 8");
  if (rows == 0) return 0;
  DISPATCH_T(dtype, (splice_gather_kernel<T><<<grid_cap(ceil_div(rows * (D / 8), 256)), 256, 0, STREAM>>>(
                        src, (const T*)table, (const T*)feats, (T*)out, rows, (int)(D / 8))));
  B200_LAUNCH_OK();
  return 0;
}

int b200_splice_scatter(const int32_t* src, const void* dout, void* d_table, void* d_feats, int64_t rows, int64_t D,
                        int dtype, void* stream) {
  B200_CHECK(D % 8 == 0, "splice_scatter: D must be a multiple of 8");
  if (rows == 0) return 0;
  DISPATCH_T(dtype, (splice_scatter_kernel<T><<<(unsigned)rows, 256, 0, STREAM>>>(src, (const T*)dout, (T*)d_table,
                                                                                  (T*)d_feats, (int)rows, (int)D)));
  B200_LAUNCH_OK();
  return 0;
}

int b200_gather_rows(const void* x, const int32_t* idx, void* out, int64_t n_idx, int64_t D, int dtype, void* stream) {
  B200_CHECK(D % 8 == 0, "gather_rows: D must be a multiple of 8");
  if (n_idx == 0) return 0;
  DISPATCH_T(dtype, (gather_rows_kernel<T><<<grid_cap(ceil_div(n_idx * (D / 8), 256)), 256, 0, STREAM>>>(
                        (const T*)x, idx, (T*)out, n_idx, (int)(D / 8))));
  B200_LAUNCH_OK();
  return 0;
}
int b200_scatter_rows_add(const void* dout, const int32_t* idx, void* dx, int64_t n_idx, int64_t D, int dtype,
                          void* stream) {
  B200_CHECK(D % 8 == 0, "scatter_rows_add: D must be a multiple of 8");
  if (n_idx == 0) return 0;
  DISPATCH_T(dtype, (scatter_rows_add_kernel<T><<<(unsigned)ceil_div(D / 8, 64), 64, 0, STREAM>>>(
                        (const T*)dout, idx, (T*)dx, n_idx, (int)(D / 8))));
  B200_LAUNCH_OK();
  return 0;
}
int b200_last_valid_index(const uint8_t* mask, int64_t B, int64_t S, int32_t* idx, void* stream) {
  if (B == 0) return 0;
  last_valid_index_kernel<<<(unsigned)ceil_div(B, 128), 128, 0, STREAM>>>(mask, (int)B, (int)S, idx);
  B200_LAUNCH_OK();
  return 0;
}

int b200_q_sample(const void* x, const void* noise, const int32_t* t, const float* sqrt_ac, const float* sqrt_1mac,
                  void* x_t, int64_t B, int64_t per_sample, int dtype, void* stream) {
  const int64_t total = B * per_sample;
  if (total == 0) return 0;
  DISPATCH_T(dtype, (q_sample_kernel<T><<<grid_cap(ceil_div(total, 256)), 256, 0, STREAM>>>(
                        (const T*)x, (const T*)noise, t, sqrt_ac, sqrt_1mac, (T*)x_t, total, (int)per_sample)));
  B200_LAUNCH_OK();
  return 0;
}
int b200_timestep_embedding(const float* t, void* out, int64_t B, int dim, float max_period, int dtype, void* stream) {
  if (B == 0) return 0;
  DISPATCH_T(dtype, (timestep_embedding_kernel<T><<<grid_cap(ceil_div(B * dim, 256)), 256, 0, STREAM>>>(
                        t, (T*)out, (int)B, dim, max_period)));
  B200_LAUNCH_OK();
  return 0;
}

int b200_discretize_actions(const float* actions, int64_t n, int n_bins, int64_t* bins, void* stream) {
  if (n == 0) return 0;
  discretize_kernel<<<grid_cap(ceil_div(n, 256)), 256, 0, STREAM>>>(actions, n, (float)(n_bins - 1), bins);
  B200_LAUNCH_OK();
  return 0;
}
int b200_bins_to_continuous(const int64_t* bins, int64_t n, int n_bins, float* out, void* stream) {
  if (n == 0) return 0;
  bins_to_cont_kernel<<<grid_cap(ceil_div(n, 256)), 256, 0, STREAM>>>(bins, n, (float)(n_bins - 1), out);
  B200_LAUNCH_OK();
  return 0;
}
int b200_argmax_last(const void* logits, int64_t rows, int64_t V, int n_last, int64_t* idx, int dtype, void* stream) {
  if (rows == 0) return 0;
  B200_CHECK(n_last > 0 && n_last <= V, "argmax_last: bad n_last");
  DISPATCH_T(dtype, (argmax_last_kernel<T><<<(unsigned)rows, 256, 0, STREAM>>>((const T*)logits, V, n_last, idx)));
  B200_LAUNCH_OK();
  return 0;
}

int b200_sample_last(const void* logits, int64_t rows, int64_t V, int n_last, float temperature, const float* u,
                     int64_t* idx, int dtype, void* stream) {
  if (rows == 0) return 0;
  B200_CHECK(n_last > 0 && n_last <= V && temperature > 0.0f && u != nullptr, "sample_last: bad arguments");
  DISPATCH_T(dtype, (sample_last_kernel<T><<<(unsigned)rows, 32, 0, STREAM>>>((const T*)logits, V, n_last,
                                                                              1.0f / temperature, u, idx)));
  B200_LAUNCH_OK();
  return 0;
}

int b200_cross_entropy_fwd(const void* logits, const int64_t* labels, int64_t rows, int64_t V, float* lse,
                           float* loss_sum, int32_t* n_valid, int dtype, void* stream) {
  if (rows == 0) return 0;
  DISPATCH_T(dtype, (ce_fwd_kernel<T><<<(unsigned)rows, 256, 0, STREAM>>>((const T*)logits, labels, V, lse, loss_sum,
                                                                          n_valid)));
  B200_LAUNCH_OK();
  return 0;
}
int b200_cross_entropy_bwd(const void* logits, const int64_t* labels, const float* lse, const int32_t* n_valid,
                           const float* gscale, void* dlogits, int64_t rows, int64_t V, int dtype, void* stream) {
  if (rows == 0) return 0;
  DISPATCH_T(dtype, (ce_bwd_kernel<T><<<(unsigned)rows, 256, 0, STREAM>>>((const T*)logits, labels, lse, n_valid,
                                                                          gscale, (T*)dlogits, V)));
  B200_LAUNCH_OK();
  return 0;
}

int b200_gemm_simt(const b200_gemm_args* a, void* stream) {
  B200_CHECK(a != nullptr, "gemm_simt: null args");
  B200_CHECK((a->z_lo <= 1) && (a->z_hi <= 1) && (a->k_segs <= 1) && !a->dual_b, "gemm_simt: no batching / dual");
  if (a->m == 0 || a->n == 0) return 0;
  dim3 grid((unsigned)ceil_div(a->n, 16), (unsigned)ceil_div(a->m, 16));
  const float alpha = a->alpha == 0.0f ? 1.0f : a->alpha;
  const int bf = a->bias_dtype == B200_F32, rf = a->res_dtype == B200_F32;
#define SIMT(TA, TD)                                                                                               \
  gemm_simt_kernel<TA, TD><<<grid, 256, 0, STREAM>>>((const TA*)a->a, (const TA*)a->b, (TD*)a->d, (int)a->m,        \
                                                     (int)a->n, (int)a->k, a->a_ld, a->b_ld, a->d_ld, a->a_mn_major, \
                                                     a->b_mn_major, alpha, a->bias, bf, a->residual, rf, a->res_ld,  \
                                                     a->act)
  if (a->ab_dtype == B200_F32 && a->d_dtype == B200_F32)
    SIMT(float, float);
  else if (a->ab_dtype == B200_F32)
    SIMT(float, bf16);
  else if (a->d_dtype == B200_F32)
    SIMT(bf16, float);
  else
    SIMT(bf16, bf16);
#undef SIMT
  B200_LAUNCH_OK();
  return 0;
}

}  // extern "C"
