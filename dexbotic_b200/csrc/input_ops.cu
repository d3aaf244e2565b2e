// GPU-side input pipeline: what the reference does per sample on CPU dataloader workers with PIL / numpy,
// as batched kernels on uint8 camera frames already resident on the device.
//
//   image_resize_h / image_resize_v_norm   PreprocessRGB.__call__ (dexbotic/data/dataset/rgb_preprocess.py:13-30) with
//       the default image_aspect_ratio='pad': expand2square (:32-44, background = the processor's mean colour or 0) ->
//       HF CLIPImageProcessor.preprocess = PIL bicubic resize to the crop size, rescale 1/255, (x - mean) / std.
//       PIL's resize is a separable convolution in 22-bit fixed point with a uint8 round trip between the horizontal
//       and the vertical pass (Pillow src/libImaging/Resample.c); both passes are restated here in the same integer
//       arithmetic, from coefficient tables the host computes exactly as Pillow's precompute_coeffs /
//       normalize_coeffs_8bpc do — the uint8 result is BIT-EXACT, and the float stage is a 256-entry table per channel
//       evaluated on the host in the processor's own arithmetic, so the tensor is bit-exact too.
//       The square padding is never materialised: out-of-frame taps read the background colour.
//   action_normalize   ActionNorm._normalize (data/dataset/transform/action.py:268-275): quantile mode
//       (x - min) / (max - min + 1e-6) * 2 - 1, or (x - mean) / (std + 1e-6); float64 arithmetic, rounded to fp32 once.
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/dexbotic_b200_ops.h"
#include "common.h"

namespace b200 {

constexpr int kPrecisionBits = 32 - 8 - 2;   // Pillow: PRECISION_BITS

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kPrecisionBits;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// horizontal pass over the (virtual) padded square [L x L]: tmp[b, y, xx, c], y in [0, L), xx in [0, out)
__global__ void __launch_bounds__(256) image_resize_h_kernel(const uint8_t* __restrict__ src, int B, int H, int W, int L,
                                                             int ox, int oy, int out, const int* __restrict__ kk,
                                                             const int* __restrict__ bounds, int ksize, int bg0, int bg1,
                                                             int bg2, uint8_t* __restrict__ tmp) {
  const long long n = (long long)B * L * out;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % out);
    const int y = (int)((i / out) % L);
    const int b = (int)(i / ((long long)out * L));
    const int xmin = bounds[2 * xx], xmax = bounds[2 * xx + 1];
    const int* k = kk + xx * ksize;
    int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
    const int sy = y - oy;
    const bool row_in = sy >= 0 && sy < H;
    const uint8_t* row = src + ((size_t)b * H + (row_in ? sy : 0)) * W * 3;
    for (int x = 0; x < xmax; ++x) {
      const int sx = x + xmin - ox;
      int p0 = bg0, p1 = bg1, p2 = bg2;
      if (row_in && sx >= 0 && sx < W) {
        p0 = row[sx * 3];
        p1 = row[sx * 3 + 1];
        p2 = row[sx * 3 + 2];
      }
      s0 += p0 * k[x];
      s1 += p1 * k[x];
      s2 += p2 * k[x];
    }
    uint8_t* o = tmp + (size_t)i * 3;
    o[0] = clip8(s0);
    o[1] = clip8(s1);
    o[2] = clip8(s2);
  }
}

// vertical pass + rescale / normalise through the per-channel table: out[b, c, yy, xx]
template <typename T>
__global__ void __launch_bounds__(256) image_resize_v_norm_kernel(const uint8_t* __restrict__ tmp, int B, int L, int out,
                                                                  const int* __restrict__ kk,
                                                                  const int* __restrict__ bounds, int ksize,
                                                                  const float* __restrict__ lut, T* __restrict__ dst,
                                                                  uint8_t* __restrict__ dst_u8) {
  const long long n = (long long)B * out * out;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % out);
    const int yy = (int)((i / out) % out);
    const int b = (int)(i / ((long long)out * out));
    const int ymin = bounds[2 * yy], ymax = bounds[2 * yy + 1];
    const int* k = kk + yy * ksize;
    int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
    for (int y = 0; y < ymax; ++y) {
      const uint8_t* px = tmp + (((size_t)b * L + (y + ymin)) * out + xx) * 3;
      s0 += px[0] * k[y];
      s1 += px[1] * k[y];
      s2 += px[2] * k[y];
    }
    const uint8_t v0 = clip8(s0), v1 = clip8(s1), v2 = clip8(s2);
    if (dst_u8 != nullptr) {
      uint8_t* o = dst_u8 + (size_t)i * 3;
      o[0] = v0;
      o[1] = v1;
      o[2] = v2;
    }
    const size_t plane = (size_t)out * out;
    const size_t base = (size_t)b * 3 * plane + (size_t)yy * out + xx;
    if constexpr (sizeof(T) == 4) {
      dst[base] = lut[v0];
      dst[base + plane] = lut[256 + v1];
      dst[base + 2 * plane] = lut[512 + v2];
    } else {
      dst[base] = __float2bfloat16(lut[v0]);
      dst[base + plane] = __float2bfloat16(lut[256 + v1]);
      dst[base + 2 * plane] = __float2bfloat16(lut[512 + v2]);
    }
  }
}

__global__ void action_normalize_kernel(const double* __restrict__ x, const double* __restrict__ a,
                                        const double* __restrict__ b, float* __restrict__ out, long long rows, int D,
                                        int quantile) {
  const long long n = rows * D;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i % D);
    // quantile: a = min, b = max ; else a = mean, b = std — the expression order of action.py:270-275
    const double v = quantile ? (x[i] - a[d]) / (b[d] - a[d] + 1e-6) * 2.0 - 1.0 : (x[i] - a[d]) / (b[d] + 1e-6);
    out[i] = (float)v;
  }
}

static inline unsigned grid_for(long long n) {
  long long g = (n + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace b200

using namespace b200;

extern "C" int b200_image_preprocess(const uint8_t* src, int64_t B, int64_t H, int64_t W, int64_t out_size,
                                     const int32_t* coeff_h, const int32_t* bounds_h, int ksize_h,
                                     const int32_t* coeff_v, const int32_t* bounds_v, int ksize_v, int bg_r, int bg_g,
                                     int bg_b, const float* lut, uint8_t* tmp, void* dst, uint8_t* dst_u8, int dtype,
                                     void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK(B > 0 && H > 0 && W > 0 && out_size > 0 && ksize_h > 0 && ksize_v > 0, "image_preprocess: bad geometry");
  B200_CHECK(src && coeff_h && bounds_h && coeff_v && bounds_v && lut && tmp && dst, "image_preprocess: null pointer");
  const int L = (int)(H > W ? H : W);
  const int ox = W >= H ? 0 : (int)((H - W) / 2), oy = W > H ? (int)((W - H) / 2) : 0;   // expand2square's paste offset
  image_resize_h_kernel<<<grid_for(B * L * out_size), 256, 0, stream>>>(src, (int)B, (int)H, (int)W, L, ox, oy,
                                                                        (int)out_size, coeff_h, bounds_h, ksize_h, bg_r,
                                                                        bg_g, bg_b, tmp);
  B200_LAUNCH_OK();
  if (dtype == B200_F32)
    image_resize_v_norm_kernel<float><<<grid_for(B * out_size * out_size), 256, 0, stream>>>(
        tmp, (int)B, L, (int)out_size, coeff_v, bounds_v, ksize_v, lut, (float*)dst, dst_u8);
  else
    image_resize_v_norm_kernel<__nv_bfloat16><<<grid_for(B * out_size * out_size), 256, 0, stream>>>(
        tmp, (int)B, L, (int)out_size, coeff_v, bounds_v, ksize_v, lut, (__nv_bfloat16*)dst, dst_u8);
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int b200_action_normalize(const double* x, const double* a, const double* b, float* out, int64_t rows,
                                     int64_t D, int quantile, void* stream) {
  if (rows * D == 0) return 0;
  action_normalize_kernel<<<grid_for(rows * D), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, a, b, out, rows,
                                                                                                  (int)D, quantile);
  B200_LAUNCH_OK();
  return 0;
}
