// 8-element packs (16 B of bf16 / 32 B of fp32) for the HBM-bound kernels: every global access
// is a 128-bit vector, consecutive lanes touch consecutive packs (fully coalesced).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

template <typename T>
struct Pack8;

template <>
struct Pack8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};

template <>
struct Pack8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    float4 a = reinterpret_cast<const float4*>(p)[0];
    float4 b = reinterpret_cast<const float4*>(p)[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
};

__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T>
__device__ __forceinline__ T from_f(float x);
template <>
__device__ __forceinline__ float from_f<float>(float x) { return x; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float x) { return __float2bfloat16(x); }

// round-trip through T (bf16 rounding where the reference rounds intermediate results)
template <typename T>
__device__ __forceinline__ float round_to(float x) { return to_f(from_f<T>(x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum for blockDim.x <= 1024 (result valid in every thread)
__device__ __forceinline__ float block_sum(float v, float* red /* >= 33 floats smem */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float t = lane < nw ? red[lane] : 0.0f;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

// Activations (codes: dexbotic_b200.h B200_ACT_*).  The templated forms are the arithmetic; the runtime forms dispatch
// once per call.  Kernels must dispatch once per PACK / register array (act_fwd_n, act_grad_mul_n), never per element:
// a switch inside an unrolled element loop compiles to one indirect branch (BRX) per element.
template <int kAct>
__device__ __forceinline__ float act_fwd_t(float x) {
  if constexpr (kAct == 1) {          // gelu erf
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
  } else if constexpr (kAct == 2) {   // gelu tanh
    const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
    return 0.5f * x * (1.0f + tanhf(u));
  } else if constexpr (kAct == 3) {   // quick gelu
    return x / (1.0f + __expf(-1.702f * x));
  } else if constexpr (kAct == 4) {   // silu
    return x / (1.0f + __expf(-x));
  } else if constexpr (kAct == 5) {
    return fmaxf(x, 0.0f);
  } else {
    return x;
  }
}
// d act(x) / dx
template <int kAct>
__device__ __forceinline__ float act_grad_t(float x) {
  if constexpr (kAct == 1) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
    return cdf + x * pdf;
  } else if constexpr (kAct == 2) {
    const float x2 = x * x;
    const float u = 0.7978845608028654f * (x + 0.044715f * x * x2);
    const float t = tanhf(u);
    const float du = 0.7978845608028654f * (1.0f + 3.0f * 0.044715f * x2);
    return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * du;
  } else if constexpr (kAct == 3) {
    const float sg = 1.0f / (1.0f + __expf(-1.702f * x));
    return sg + 1.702f * x * sg * (1.0f - sg);
  } else if constexpr (kAct == 4) {
    const float sg = 1.0f / (1.0f + __expf(-x));
    return sg + x * sg * (1.0f - sg);
  } else if constexpr (kAct == 5) {
    return x > 0.0f ? 1.0f : 0.0f;
  } else {
    return 1.0f;
  }
}
__device__ __forceinline__ float act_fwd(float x, int act) {
  switch (act) {
    case 1: return act_fwd_t<1>(x);
    case 2: return act_fwd_t<2>(x);
    case 3: return act_fwd_t<3>(x);
    case 4: return act_fwd_t<4>(x);
    case 5: return act_fwd_t<5>(x);
    default: return x;
  }
}
__device__ __forceinline__ float act_grad(float x, int act) {
  switch (act) {
    case 1: return act_grad_t<1>(x);
    case 2: return act_grad_t<2>(x);
    case 3: return act_grad_t<3>(x);
    case 4: return act_grad_t<4>(x);
    case 5: return act_grad_t<5>(x);
    default: return 1.0f;
  }
}
// v[j] = act(v[j]) / d[j] *= act'(x[j]) over a register array, one dispatch for the whole array
#define B200_ACT_CASE_N(K, BODY)   \
  case K: {                        \
    _Pragma("unroll") for (int j = 0; j < N; ++j) { BODY; } \
    break;                         \
  }
template <int N>
__device__ __forceinline__ void act_fwd_n(float (&v)[N], int act) {
  switch (act) {
    B200_ACT_CASE_N(1, v[j] = act_fwd_t<1>(v[j]))
    B200_ACT_CASE_N(2, v[j] = act_fwd_t<2>(v[j]))
    B200_ACT_CASE_N(3, v[j] = act_fwd_t<3>(v[j]))
    B200_ACT_CASE_N(4, v[j] = act_fwd_t<4>(v[j]))
    B200_ACT_CASE_N(5, v[j] = act_fwd_t<5>(v[j]))
    default: break;
  }
}
template <int N>
__device__ __forceinline__ void act_grad_mul_n(float (&d)[N], const float (&x)[N], int act) {
  switch (act) {
    B200_ACT_CASE_N(1, d[j] *= act_grad_t<1>(x[j]))
    B200_ACT_CASE_N(2, d[j] *= act_grad_t<2>(x[j]))
    B200_ACT_CASE_N(3, d[j] *= act_grad_t<3>(x[j]))
    B200_ACT_CASE_N(4, d[j] *= act_grad_t<4>(x[j]))
    B200_ACT_CASE_N(5, d[j] *= act_grad_t<5>(x[j]))
    default: break;
  }
}
#undef B200_ACT_CASE_N

// act(x) and act'(x) together.  kAct = 4 (SiLU) / 2 (tanh-GELU) are the bf16 fast paths: one MUFU.EX2 + one MUFU.RCP
// per element instead of two exponentials and IEEE divisions — at 8 elements per 16-byte pack the
// generic path is MUFU-bound, not HBM-bound.  Their error (2 ulp fp32) vanishes in the bf16 rounding of the outputs.
// kAct = -1: the exact runtime-dispatched functions (fp32 tensors, other activations).
template <int kAct>
__device__ __forceinline__ void act_pair(float x, int act, float& f, float& df) {
  if constexpr (kAct == 4) {
    const float s = __fdividef(1.0f, 1.0f + __expf(-x));
    f = x * s;
    df = s + f * (1.0f - s);
  } else if constexpr (kAct == 2) {
    const float x2 = x * x;
    // tanh(u) = 1 - 2 / (1 + e^{2u}): one ex2 + one rcp, ~2 ulp (tanh.approx.f32 would be 2^-11)
    const float t = 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * 0.7978845608028654f * (x + 0.044715f * x * x2)));
    f = 0.5f * x * (1.0f + t);
    df = 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * 0.7978845608028654f * (1.0f + 0.134145f * x2);
  } else {
    f = act_fwd(x, act);
    df = act_grad(x, act);
  }
}
template <int kAct>
__device__ __forceinline__ float act_only(float x, int act) {
  if constexpr (kAct == 4) {
    return x * __fdividef(1.0f, 1.0f + __expf(-x));
  } else if constexpr (kAct == 2) {
    const float t = 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * 0.7978845608028654f * (x + 0.044715f * x * x * x)));
    return 0.5f * x * (1.0f + t);
  } else {
    return act_fwd(x, act);
  }
}

}  // namespace b200
