// Data-parallel exchange kernel over NVLink peer memory (reference: DeepSpeed ZeRO reduce-scatter behind
// script/deepspeed/zero2.json, base_exp.py:229).  The gradient buffers of all ranks live in symmetric memory
// (torch.distributed._symmetric_memory), so a rank's kernel can LOAD its piece of a chunk straight out of every peer's
// buffer: one pass, sum in fp32 in a fixed rank order (deterministic), average, store bf16 in place.  HBM traffic per
// element of the piece: one read + one write locally, one read on each peer — the copy-engine version staged the peers'
// pieces and ran N torch elementwise passes over an fp32 accumulator (30 B / element at N = 2, 90 B at N = 8).
#include "../../include/dexbotic_b200_ops.h"
#include "common.h"
#include "vec.cuh"

namespace b200 {
using bf16 = __nv_bfloat16;

constexpr int kMaxPeers = 7;       // one NVSwitch node: up to 8 ranks (28 raw registers per thread in flight)
struct PeerPtrs {
  const bf16* p[kMaxPeers];
};

// The loads cross NVLink (~2-3 us latency): every thread keeps kUnroll packs x all peers in flight before it adds.  The
// grid is deliberately small (the persistent GEMM owns the SMs), so the bytes in flight per thread set the bandwidth:
// with one pack per peer a 16-CTA grid moved ~30 GB/s at N = 2 (one 16-byte load in flight per thread) — a 233 MB piece
// took as long as the decoder block's backward it hides under, i.e. the kernel was resident all the time.  kUnroll = 4
// (N = 2), 2 (N <= 4), 1 (N = 8: seven loads per pack already) keeps ~4-8 loads in flight per thread; the register array
// is sized per instance (16 / 24 / 28 registers) so that a block still fits next to the GEMM's CTA on an SM.
template <int kUnroll, int kPeers>      // kPeers: capacity of the in-flight register array (n_peers <= kPeers)
__global__ void __launch_bounds__(256) reduce_scatter_p2p_kernel(bf16* __restrict__ own, PeerPtrs peers, int n_peers,
                                                                 int64_t n8, float scale) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < n8; i0 += stride * kUnroll) {
    uint4 raw[kUnroll][kPeers];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t i = i0 + u * stride;
#pragma unroll
      for (int r = 0; r < kPeers; ++r)
        if (r < n_peers && i < n8) raw[u][r] = __ldcv(reinterpret_cast<const uint4*>(peers.p[r]) + i);   // never cached stale
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t i = i0 + u * stride;
      if (i >= n8) break;
      float acc[8];
      Pack8<bf16>::load(own + i * 8, acc);
#pragma unroll
      for (int r = 0; r < kPeers; ++r) {
        if (r < n_peers) {
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw[u][r]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __bfloat1622float2(h[e]);
            acc[2 * e] += f.x;
            acc[2 * e + 1] += f.y;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] *= scale;
      Pack8<bf16>::store(own + i * 8, acc);
    }
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_reduce_scatter_p2p(void* own, const void* const* peer_ptrs, int n_peers, int64_t n, float scale,
                                       int ctas, void* stream) {
  if (n == 0) return 0;
  B200_CHECK(n_peers >= 1 && n_peers <= kMaxPeers, "reduce_scatter_p2p: %d peers (1..%d supported)", n_peers, kMaxPeers);
  B200_CHECK(n % 8 == 0 && (reinterpret_cast<uintptr_t>(own) & 15) == 0, "reduce_scatter_p2p: piece must be 16-byte packs");
  PeerPtrs pp;
  for (int r = 0; r < kMaxPeers; ++r) {
    pp.p[r] = r < n_peers ? reinterpret_cast<const bf16*>(peer_ptrs[r]) : nullptr;
    if (r < n_peers) B200_CHECK((reinterpret_cast<uintptr_t>(peer_ptrs[r]) & 15) == 0, "reduce_scatter_p2p: peer pointer alignment");
  }
  if (ctas <= 0) ctas = 16;
  const int64_t want = ceil_div(n / 8, 256);
  const unsigned grid = (unsigned)(want < ctas ? want : ctas);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (n_peers == 1)
    reduce_scatter_p2p_kernel<4, 1><<<grid, 256, 0, st>>>(reinterpret_cast<bf16*>(own), pp, n_peers, n / 8, scale);
  else if (n_peers <= 3)
    reduce_scatter_p2p_kernel<2, 3><<<grid, 256, 0, st>>>(reinterpret_cast<bf16*>(own), pp, n_peers, n / 8, scale);
  else
    reduce_scatter_p2p_kernel<1, kMaxPeers><<<grid, 256, 0, st>>>(reinterpret_cast<bf16*>(own), pp, n_peers, n / 8, scale);
  B200_LAUNCH_OK();
  return 0;
}
