// ViT front-end kernels: patchify (im2col for the stride=patch conv), CLS/pos-emb assembly and its
// backward, plus fp32 -> param-grad accumulation.  Reference call sites: include/dexbotic_b200_ops.h.
#include "../../include/dexbotic_b200_ops.h"
#include "common.h"
#include "vec.cuh"

namespace b200 {
using bf16 = __nv_bfloat16;

// out[(b*P + py*gw + px), c*ps*ps + i*ps + j] = img[b, c, py*ps+i, px*ps+j]; columns >= C*ps*ps are zero.
template <typename TI, typename TO>
__global__ void im2col_kernel(const TI* __restrict__ img, TO* __restrict__ out, int B, int C, int H, int W, int ps,
                              int Kpad) {
  const int gh = H / ps, gw = W / ps;
  const int K = C * ps * ps;
  const int64_t total = (int64_t)B * gh * gw * Kpad;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int col = (int)(i % Kpad);
    const int64_t row = i / Kpad;
    float v = 0.0f;
    if (col < K) {
      const int c = col / (ps * ps), r = col - c * ps * ps;
      const int ii = r / ps, jj = r - ii * ps;
      const int px = (int)(row % gw);
      const int py = (int)((row / gw) % gh);
      const int b = (int)(row / ((int64_t)gw * gh));
      v = to_f(img[(((size_t)b * C + c) * H + py * ps + ii) * W + px * ps + jj]);
    }
    out[i] = from_f<TO>(v);
  }
}

// out[b,0,:] = cls + pos[0];  out[b,1+p,:] = patches[b,p,:] + pos[1+p]
template <typename T>
__global__ void vit_embed_fwd_kernel(const T* __restrict__ patches, const T* __restrict__ cls,
                                     const T* __restrict__ pos, T* __restrict__ out, int B, int P, int D8) {
  const int64_t total = (int64_t)B * (P + 1) * D8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % D8) * 8;
    const int64_t r = i / D8;
    const int t = (int)(r % (P + 1));
    const int b = (int)(r / (P + 1));
    float a[8], e[8];
    if (t == 0)
      Pack8<T>::load(cls + c, a);
    else
      Pack8<T>::load(patches + ((size_t)b * P + (t - 1)) * D8 * 8 + c, a);
    Pack8<T>::load(pos + (size_t)t * D8 * 8 + c, e);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += e[j];
    Pack8<T>::store(out + r * D8 * 8 + c, a);
  }
}
// d_patches[b,p] = dout[b,1+p];  d_pos[t] += sum_b dout[b,t];  d_cls += sum_b dout[b,0]   (fp32 accumulators)
template <typename T>
__global__ void vit_embed_bwd_kernel(const T* __restrict__ dout, T* __restrict__ d_patches, float* __restrict__ d_cls,
                                     float* __restrict__ d_pos, int B, int P, int D8) {
  const int64_t total = (int64_t)(P + 1) * D8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % D8) * 8;
    const int t = (int)(i / D8);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int b = 0; b < B; ++b) {
      float v[8];
      Pack8<T>::load(dout + ((size_t)b * (P + 1) + t) * D8 * 8 + c, v);
      if (t > 0 && d_patches != nullptr) Pack8<T>::store(d_patches + ((size_t)b * P + (t - 1)) * D8 * 8 + c, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (d_pos != nullptr) d_pos[(size_t)t * D8 * 8 + c + j] += acc[j];
      if (t == 0 && d_cls != nullptr) d_cls[c + j] += acc[j];
    }
  }
}

// dst (+)= src   (fp32 scratch -> parameter gradient of either dtype)
template <typename D>
__global__ void cast_add_kernel(const float* __restrict__ src, D* __restrict__ dst, int64_t n, int accumulate) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = src[i];
    if (accumulate) v += to_f(dst[i]);
    dst[i] = from_f<D>(v);
  }
}
// dst[r, 0:cols] (+)= src[r, 0:cols] with different leading dims (padded patch-embedding weights)
template <typename S, typename D>
__global__ void copy2d_kernel(const S* __restrict__ src, D* __restrict__ dst, int rows, int cols, int64_t lds,
                              int64_t ldd, int accumulate) {
  const int64_t total = (int64_t)rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (int64_t)r * cols);
    float v = to_f(src[(size_t)r * lds + c]);
    if (accumulate) v += to_f(dst[(size_t)r * ldd + c]);
    dst[(size_t)r * ldd + c] = from_f<D>(v);
  }
}

// dst[b, r, 0:cols] (+)= alpha * src[b, r, 0:cols]   (row-block moves between packed / joint sequence buffers)
template <typename T>
__global__ void copy3d_kernel(const T* __restrict__ src, T* __restrict__ dst, int B, int rows, int cols8, int64_t sbs,
                              int64_t sld, int64_t dbs, int64_t dld, float alpha, int accumulate) {
  const int64_t total = (int64_t)B * rows * cols8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % cols8) * 8;
    const int64_t t = i / cols8;
    const int r = (int)(t % rows);
    const int b = (int)(t / rows);
    float v[8];
    Pack8<T>::load(src + b * sbs + r * sld + c, v);
    T* d = dst + b * dbs + r * dld + c;
    if (accumulate) {
      float o[8];
      Pack8<T>::load(d, o);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = o[j] + alpha * v[j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= alpha;
    }
    Pack8<T>::store(d, v);
  }
}
// out[b, p, :] = x[b, p, :] + pos[p, :]    (SigLIP position embedding, no CLS token)
template <typename T>
__global__ void add_pos_fwd_kernel(const T* __restrict__ x, const T* __restrict__ pos, T* __restrict__ out, int B, int P,
                                   int D8) {
  const int64_t total = (int64_t)B * P * D8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % D8) * 8;
    const int64_t r = i / D8;
    const int p = (int)(r % P);
    float a[8], e[8];
    Pack8<T>::load(x + r * D8 * 8 + c, a);
    Pack8<T>::load(pos + (size_t)p * D8 * 8 + c, e);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += e[j];
    Pack8<T>::store(out + r * D8 * 8 + c, a);
  }
}
// d_pos[p, :] += sum_b dout[b, p, :]
template <typename T>
__global__ void add_pos_bwd_kernel(const T* __restrict__ dout, float* __restrict__ d_pos, int B, int P, int D8) {
  const int64_t total = (int64_t)P * D8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % D8) * 8;
    const int p = (int)(i / D8);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int b = 0; b < B; ++b) {
      float v[8];
      Pack8<T>::load(dout + ((size_t)b * P + p) * D8 * 8 + c, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) d_pos[(size_t)p * D8 * 8 + c + j] += acc[j];
  }
}

static inline int grid_cap(int64_t want) {
  int64_t cap = (int64_t)num_sms() * 8;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}
}  // namespace b200

using namespace b200;
#define STREAM reinterpret_cast<cudaStream_t>(stream)

extern "C" {

int b200_im2col_patches(const void* images, void* out, int64_t B, int C, int H, int W, int patch, int Kpad,
                        int in_dtype, int out_dtype, void* stream) {
  B200_CHECK(H % patch == 0 && W % patch == 0 && Kpad >= C * patch * patch, "im2col: bad geometry");
  const int64_t total = B * (H / patch) * (W / patch) * Kpad;
  if (total == 0) return 0;
  const int g = grid_cap(ceil_div(total, 256));
  if (in_dtype == B200_F32 && out_dtype == B200_BF16)
    im2col_kernel<float, bf16><<<g, 256, 0, STREAM>>>((const float*)images, (bf16*)out, (int)B, C, H, W, patch, Kpad);
  else if (in_dtype == B200_BF16 && out_dtype == B200_BF16)
    im2col_kernel<bf16, bf16><<<g, 256, 0, STREAM>>>((const bf16*)images, (bf16*)out, (int)B, C, H, W, patch, Kpad);
  else if (in_dtype == B200_F32)
    im2col_kernel<float, float><<<g, 256, 0, STREAM>>>((const float*)images, (float*)out, (int)B, C, H, W, patch, Kpad);
  else
    im2col_kernel<bf16, float><<<g, 256, 0, STREAM>>>((const bf16*)images, (float*)out, (int)B, C, H, W, patch, Kpad);
  B200_LAUNCH_OK();
  return 0;
}

int b200_vit_embed_fwd(const void* patches, const void* cls, const void* pos, void* out, int64_t B, int64_t P,
                       int64_t D, int dtype, void* stream) {
  B200_CHECK(D % 8 == 0, "vit_embed_fwd: D must be a multiple of 8");
  const int64_t total = B * (P + 1) * (D / 8);
  if (total == 0) return 0;
  if (dtype == B200_F32)
    vit_embed_fwd_kernel<float><<<grid_cap(ceil_div(total, 256)), 256, 0, STREAM>>>(
        (const float*)patches, (const float*)cls, (const float*)pos, (float*)out, (int)B, (int)P, (int)(D / 8));
  else
    vit_embed_fwd_kernel<bf16><<<grid_cap(ceil_div(total, 256)), 256, 0, STREAM>>>(
        (const bf16*)patches, (const bf16*)cls, (const bf16*)pos, (bf16*)out, (int)B, (int)P, (int)(D / 8));
  B200_LAUNCH_OK();
  return 0;
}

int b200_vit_embed_bwd(const void* dout, void* d_patches, float* d_cls, float* d_pos, int64_t B, int64_t P, int64_t D,
                       int dtype, void* stream) {
  B200_CHECK(D % 8 == 0, "vit_embed_bwd: D must be a multiple of 8");
  const int64_t total = (P + 1) * (D / 8);
  if (total == 0 || B == 0) return 0;
  if (dtype == B200_F32)
    vit_embed_bwd_kernel<float><<<grid_cap(ceil_div(total, 128)), 128, 0, STREAM>>>(
        (const float*)dout, (float*)d_patches, d_cls, d_pos, (int)B, (int)P, (int)(D / 8));
  else
    vit_embed_bwd_kernel<bf16><<<grid_cap(ceil_div(total, 128)), 128, 0, STREAM>>>(
        (const bf16*)dout, (bf16*)d_patches, d_cls, d_pos, (int)B, (int)P, (int)(D / 8));
  B200_LAUNCH_OK();
  return 0;
}

int b200_cast_add(const float* src, void* dst, int64_t n, int dst_dtype, int accumulate, void* stream) {
  if (n == 0) return 0;
  if (dst_dtype == B200_F32)
    cast_add_kernel<float><<<grid_cap(ceil_div(n, 256)), 256, 0, STREAM>>>(src, (float*)dst, n, accumulate);
  else
    cast_add_kernel<bf16><<<grid_cap(ceil_div(n, 256)), 256, 0, STREAM>>>(src, (bf16*)dst, n, accumulate);
  B200_LAUNCH_OK();
  return 0;
}

int b200_copy3d(const void* src, void* dst, int64_t B, int64_t rows, int64_t cols, int64_t src_bs, int64_t src_ld,
                int64_t dst_bs, int64_t dst_ld, float alpha, int accumulate, int dtype, void* stream) {
  B200_CHECK(cols % 8 == 0 && src_ld % 8 == 0 && dst_ld % 8 == 0 && src_bs % 8 == 0 && dst_bs % 8 == 0,
             "copy3d: cols / strides must be multiples of 8");
  const int64_t total = B * rows * (cols / 8);
  if (total == 0) return 0;
  const int g = grid_cap(ceil_div(total, 256));
  if (dtype == B200_F32)
    copy3d_kernel<float><<<g, 256, 0, STREAM>>>((const float*)src, (float*)dst, (int)B, (int)rows, (int)(cols / 8),
                                                 src_bs, src_ld, dst_bs, dst_ld, alpha, accumulate);
  else
    copy3d_kernel<bf16><<<g, 256, 0, STREAM>>>((const bf16*)src, (bf16*)dst, (int)B, (int)rows, (int)(cols / 8), src_bs,
                                                src_ld, dst_bs, dst_ld, alpha, accumulate);
  B200_LAUNCH_OK();
  return 0;
}

int b200_add_pos_fwd(const void* x, const void* pos, void* out, int64_t B, int64_t P, int64_t D, int dtype,
                     void* stream) {
  B200_CHECK(D % 8 == 0, "add_pos_fwd: D must be a multiple of 8");
  const int64_t total = B * P * (D / 8);
  if (total == 0) return 0;
  const int g = grid_cap(ceil_div(total, 256));
  if (dtype == B200_F32)
    add_pos_fwd_kernel<float><<<g, 256, 0, STREAM>>>((const float*)x, (const float*)pos, (float*)out, (int)B, (int)P,
                                                      (int)(D / 8));
  else
    add_pos_fwd_kernel<bf16><<<g, 256, 0, STREAM>>>((const bf16*)x, (const bf16*)pos, (bf16*)out, (int)B, (int)P,
                                                     (int)(D / 8));
  B200_LAUNCH_OK();
  return 0;
}

int b200_add_pos_bwd(const void* dout, float* d_pos, int64_t B, int64_t P, int64_t D, int dtype, void* stream) {
  B200_CHECK(D % 8 == 0, "add_pos_bwd: D must be a multiple of 8");
  const int64_t total = P * (D / 8);
  if (total == 0 || B == 0) return 0;
  const int g = grid_cap(ceil_div(total, 128));
  if (dtype == B200_F32)
    add_pos_bwd_kernel<float><<<g, 128, 0, STREAM>>>((const float*)dout, d_pos, (int)B, (int)P, (int)(D / 8));
  else
    add_pos_bwd_kernel<bf16><<<g, 128, 0, STREAM>>>((const bf16*)dout, d_pos, (int)B, (int)P, (int)(D / 8));
  B200_LAUNCH_OK();
  return 0;
}

int b200_copy2d(const void* src, void* dst, int64_t rows, int64_t cols, int64_t lds, int64_t ldd, int src_dtype,
                int dst_dtype, int accumulate, void* stream) {
  if (rows * cols == 0) return 0;
  const int g = grid_cap(ceil_div(rows * cols, 256));
#define C2D(S, D) \
  copy2d_kernel<S, D><<<g, 256, 0, STREAM>>>((const S*)src, (D*)dst, (int)rows, (int)cols, lds, ldd, accumulate)
  if (src_dtype == B200_F32 && dst_dtype == B200_F32)
    C2D(float, float);
  else if (src_dtype == B200_F32)
    C2D(float, bf16);
  else if (dst_dtype == B200_F32)
    C2D(bf16, float);
  else
    C2D(bf16, bf16);
#undef C2D
  B200_LAUNCH_OK();
  return 0;
}

}  // extern "C"
