// MemVLA memory-path kernels (memvla_arch.py): stateless counter-based dropout, squeeze-excite pooling / channel
// rescale of BottleneckSE and their backward passes, and the sigmoid gate of GateFusion.  All HBM-bound, 128-bit packs.
#include "../../include/dexbotic_b200_ops.h"
#include "common.h"
#include "vec.cuh"

namespace b200 {
using bf16 = __nv_bfloat16;

static inline int mem_grid_cap(int64_t want) {
  const int64_t cap = (int64_t)num_sms() * 8;
  return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

// splitmix64 finaliser over (seed, element index): one independent 24-bit uniform per element, reproducible from
// (seed, row, col) alone so backward regenerates the forward mask instead of storing it.
__device__ __forceinline__ float uniform01(uint64_t seed, uint64_t idx) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (float)(uint32_t)(z >> 40) * (1.0f / 16777216.0f);
}

template <typename T>
__global__ void dropout_kernel(const T* __restrict__ x, T* __restrict__ out, int64_t rows, int cols, int64_t x_ld,
                               int64_t out_ld, float p, float inv_keep, uint64_t seed) {
  const int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols;
    const int c = (int)(i - r * cols);
    const float v = to_f(x[r * x_ld + c]);
    out[r * out_ld + c] = from_f<T>(uniform01(seed, (uint64_t)i) >= p ? v * inv_keep : 0.0f);
  }
}

// sum_f32[b, c] += scale * sum_{p in chunk} x[b,p,c] * (y ? y[b,p,c] : 1)     (grid: x = C/8 packs, y = P chunks, z = B)
template <typename T>
__global__ void se_reduce_kernel(const T* __restrict__ x, const T* __restrict__ y, float* __restrict__ out, int P, int C8,
                                 int chunk, float scale) {
  const int c8 = blockIdx.x * blockDim.x + threadIdx.x;
  if (c8 >= C8) return;
  const int b = blockIdx.z;
  const int p0 = blockIdx.y * chunk;
  const int p1 = min(P, p0 + chunk);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const size_t base = (size_t)b * P * C8 * 8 + (size_t)c8 * 8;
  for (int p = p0; p < p1; ++p) {
    float v[8];
    Pack8<T>::load(x + base + (size_t)p * C8 * 8, v);
    if (y != nullptr) {
      float w[8];
      Pack8<T>::load(y + base + (size_t)p * C8 * 8, w);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j] * w[j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
  float* o = out + (size_t)b * C8 * 8 + (size_t)c8 * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) atomicAdd(o + j, acc[j] * scale);
}

// out[b,p,c] = x[b,p,c] * w[b,c] + (add ? add[b,c] * add_scale : 0)
template <typename T>
__global__ void se_scale_kernel(const T* __restrict__ x, const T* __restrict__ w, const T* __restrict__ add,
                                T* __restrict__ out, int B, int P, int C8, float add_scale) {
  const int64_t total = (int64_t)B * P * C8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    const int b = (int)(i / ((int64_t)P * C8));
    float v[8], s[8];
    Pack8<T>::load(x + i * 8, v);
    Pack8<T>::load(w + ((size_t)b * C8 + c8) * 8, s);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= s[j];
    if (add != nullptr) {
      Pack8<T>::load(add + ((size_t)b * C8 + c8) * 8, s);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += s[j] * add_scale;
    }
    Pack8<T>::store(out + i * 8, v);
  }
}

// GateFusion (memvla_arch.py:176-192): s = sigmoid(z); out = s*x1 + (1-s)*x2
template <typename T>
__global__ void gate_fuse_fwd_kernel(const T* __restrict__ z, const T* __restrict__ x1, const T* __restrict__ x2,
                                     T* __restrict__ out, int64_t n8) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float a[8], b[8], g[8];
    Pack8<T>::load(z + i * 8, g);
    Pack8<T>::load(x1 + i * 8, a);
    Pack8<T>::load(x2 + i * 8, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float s = 1.0f / (1.0f + __expf(-g[j]));
      a[j] = s * a[j] + (1.0f - s) * b[j];
    }
    Pack8<T>::store(out + i * 8, a);
  }
}
// dz = dout*(x1-x2)*s*(1-s); dx1 = dout*s; dx2 = dout*(1-s)
template <typename T>
__global__ void gate_fuse_bwd_kernel(const T* __restrict__ dout, const T* __restrict__ z, const T* __restrict__ x1,
                                     const T* __restrict__ x2, T* __restrict__ dz, T* __restrict__ dx1,
                                     T* __restrict__ dx2, int64_t n8) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float d[8], a[8], b[8], g[8], o1[8], o2[8], oz[8];
    Pack8<T>::load(dout + i * 8, d);
    Pack8<T>::load(z + i * 8, g);
    Pack8<T>::load(x1 + i * 8, a);
    Pack8<T>::load(x2 + i * 8, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float s = 1.0f / (1.0f + __expf(-g[j]));
      o1[j] = d[j] * s;
      o2[j] = d[j] * (1.0f - s);
      oz[j] = d[j] * (a[j] - b[j]) * s * (1.0f - s);
    }
    Pack8<T>::store(dz + i * 8, oz);
    Pack8<T>::store(dx1 + i * 8, o1);
    Pack8<T>::store(dx2 + i * 8, o2);
  }
}

}  // namespace b200

using namespace b200;
#define STREAM reinterpret_cast<cudaStream_t>(stream)

extern "C" {

int b200_dropout(const void* x, void* out, int64_t rows, int64_t cols, int64_t x_ld, int64_t out_ld, float p,
                 uint64_t seed, int dtype, void* stream) {
  B200_CHECK(p >= 0.0f && p < 1.0f, "dropout: p must be in [0, 1)");
  if (rows * cols == 0) return 0;
  const int g = mem_grid_cap(ceil_div(rows * cols, 256));
  const float inv = 1.0f / (1.0f - p);
  if (dtype == B200_F32)
    dropout_kernel<float><<<g, 256, 0, STREAM>>>((const float*)x, (float*)out, rows, (int)cols, x_ld, out_ld, p, inv, seed);
  else
    dropout_kernel<bf16><<<g, 256, 0, STREAM>>>((const bf16*)x, (bf16*)out, rows, (int)cols, x_ld, out_ld, p, inv, seed);
  B200_LAUNCH_OK();
  return 0;
}

int b200_se_reduce(const void* x, const void* y, float* out_f32, int64_t B, int64_t P, int64_t C, float scale, int dtype,
                   void* stream) {
  B200_CHECK(C % 8 == 0, "se_reduce: C must be a multiple of 8");
  if (B * P * C == 0) return 0;
  const int C8 = (int)(C / 8);
  const int chunk = 16;
  dim3 grid((unsigned)ceil_div(C8, 128), (unsigned)ceil_div(P, chunk), (unsigned)B);
  if (dtype == B200_F32)
    se_reduce_kernel<float><<<grid, 128, 0, STREAM>>>((const float*)x, (const float*)y, out_f32, (int)P, C8, chunk, scale);
  else
    se_reduce_kernel<bf16><<<grid, 128, 0, STREAM>>>((const bf16*)x, (const bf16*)y, out_f32, (int)P, C8, chunk, scale);
  B200_LAUNCH_OK();
  return 0;
}

int b200_se_scale(const void* x, const void* w, const void* add, void* out, int64_t B, int64_t P, int64_t C,
                  float add_scale, int dtype, void* stream) {
  B200_CHECK(C % 8 == 0, "se_scale: C must be a multiple of 8");
  const int64_t total = B * P * (C / 8);
  if (total == 0) return 0;
  const int g = mem_grid_cap(ceil_div(total, 256));
  if (dtype == B200_F32)
    se_scale_kernel<float><<<g, 256, 0, STREAM>>>((const float*)x, (const float*)w, (const float*)add, (float*)out, (int)B,
                                                   (int)P, (int)(C / 8), add_scale);
  else
    se_scale_kernel<bf16><<<g, 256, 0, STREAM>>>((const bf16*)x, (const bf16*)w, (const bf16*)add, (bf16*)out, (int)B,
                                                  (int)P, (int)(C / 8), add_scale);
  B200_LAUNCH_OK();
  return 0;
}

int b200_gate_fuse_fwd(const void* z, const void* x1, const void* x2, void* out, int64_t n, int dtype, void* stream) {
  B200_CHECK(n % 8 == 0, "gate_fuse_fwd: n must be a multiple of 8");
  if (n == 0) return 0;
  const int g = mem_grid_cap(ceil_div(n / 8, 256));
  if (dtype == B200_F32)
    gate_fuse_fwd_kernel<float><<<g, 256, 0, STREAM>>>((const float*)z, (const float*)x1, (const float*)x2, (float*)out, n / 8);
  else
    gate_fuse_fwd_kernel<bf16><<<g, 256, 0, STREAM>>>((const bf16*)z, (const bf16*)x1, (const bf16*)x2, (bf16*)out, n / 8);
  B200_LAUNCH_OK();
  return 0;
}

int b200_gate_fuse_bwd(const void* dout, const void* z, const void* x1, const void* x2, void* dz, void* dx1, void* dx2,
                       int64_t n, int dtype, void* stream) {
  B200_CHECK(n % 8 == 0, "gate_fuse_bwd: n must be a multiple of 8");
  if (n == 0) return 0;
  const int g = mem_grid_cap(ceil_div(n / 8, 256));
  if (dtype == B200_F32)
    gate_fuse_bwd_kernel<float><<<g, 256, 0, STREAM>>>((const float*)dout, (const float*)z, (const float*)x1,
                                                        (const float*)x2, (float*)dz, (float*)dx1, (float*)dx2, n / 8);
  else
    gate_fuse_bwd_kernel<bf16><<<g, 256, 0, STREAM>>>((const bf16*)dout, (const bf16*)z, (const bf16*)x1, (const bf16*)x2,
                                                       (bf16*)dz, (bf16*)dx1, (bf16*)dx2, n / 8);
  B200_LAUNCH_OK();
  return 0;
}

}  // extern "C"
