// tcgen05 GEMM for sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma (TMEM accumulators,
// double buffered) -> fused epilogue -> swizzled smem staging -> TMA store.
//
// One persistent CTA per SM, 8 warps:
//   warp 0   TMA producer (one lane)            warp 2   TMEM allocator
//   warp 1   MMA issuer   (one lane)            warps 4-7 epilogue (TMEM lane quarter = warp % 4)
// Three pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue), static
// round-robin tile schedule with an M-grouped order for L2 reuse.
//
// Everything is sized in BYTES so bf16 (kind::f16) and fp32 (kind::tf32) share one kernel:
// a k-block is 128 bytes of K per row (64 bf16 / 32 fp32) = 4 UMMA instructions.
// Operands may be K-major or MN-major (both via SWIZZLE_128B canonical layouts), so the
// forward (x W^T), dgrad (dy W) and wgrad (dy^T x) forms of nn.Linear and the batched
// attention products QK^T, PV, dP, dQ, dK, dV all run here without any transposes.
//
// Reference arithmetic replaced: see include/dexbotic_b200.h (b200_gemm).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/dexbotic_b200.h"
#include "common.h"
#include "ptx.cuh"
#include "vec.cuh"

namespace b200 {

constexpr int kBlockM = 128;
constexpr int kStageABytes = kBlockM * 128;  // 16 KB
constexpr int kGemmThreads = 256;
constexpr int kGroupM = 16;
constexpr int kStagingPerWarp = 2 * 32 * 128;  // two 32-row x 128-byte buffers

struct GemmKParams {
  CUtensorMap tmA, tmB, tmB2, tmD;
  int M, N;
  int m_blocks, n_blocks, tiles_per_batch, total_tiles;
  int z_lo;
  int kbps;      // k-blocks per K segment
  int k_blocks;  // total k-blocks (= kbps * k_segs)
  int bk_elems;  // elements of K per k-block (64 bf16 / 32 fp32)
  int a_div, a_mul, a_seg, b_div, b_mul, b_seg;
  int a_mn, b_mn;
  int a_5d, b_5d;  // MN-major operand described by a 5-D map (all atoms of a tile in one TMA)
  int atom_elems;  // MN elements per 128-byte swizzle atom
  int atom_bytes;  // bytes of one MN-major atom (bk_elems rows x 128 B)
  int a_kadv, b_kadv, a_lbo, b_lbo, a_sbo, b_sbo, a_lt, b_lt;
  int ab_fp32, d_fp32;
  int dual;
  int glu_bwd;          // epilogue = SwiGLU / GeGLU backward: D = dg, d2 = du from the accumulator (dh) and g, u
  const void* glu_g;
  const void* glu_u;
  void* d2;
  long long glu_ld;
  int n_per_tile;  // output columns per tile (kBlockN, or 128 in dual mode)
  float alpha;
  const void* bias;
  int bias_fp32;
  const void* res;
  int res_fp32;
  long long res_ld, res_s2, res_s3;
  void* aux;
  void* aux2;
  long long aux_ld, aux_s2, aux_s3;
  int act;
  int debug;  // bit0: skip TMA loads (MMA-only pipeline), bit1: skip MMAs (load-only pipeline) — perf experiments
};

template <int kBlockN, int kCtas = 1>
struct GemmCfg {
  static constexpr int kStageBBytes = kBlockN / kCtas * 128;  // a CTA pair splits the B tile
  static constexpr int kStageBytes = kStageABytes + kStageBBytes;
  static constexpr int kStages = (kBlockN == 256 && kCtas == 1) ? 4 : (kBlockN / kCtas == 128 ? 6 : 8);
  static constexpr int kTmemCols = 2 * kBlockN;  // two accumulator stages
  static constexpr int kBarBytes = 256;
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + 4 * kStagingPerWarp + kBarBytes;
};

// Load up to 32 consecutive elements starting at p (16-byte aligned when valid >= 32).
template <typename T>
__device__ __forceinline__ void load_row32(const T* p, int valid, float (&out)[32]);

template <>
__device__ __forceinline__ void load_row32<__nv_bfloat16>(const __nv_bfloat16* p, int valid, float (&out)[32]) {
  if (valid >= 32) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 v = __ldg(q + i);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = __bfloat1622float2(h[j]);
        out[i * 8 + j * 2] = f.x;
        out[i * 8 + j * 2 + 1] = f.y;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) out[i] = i < valid ? __bfloat162float(p[i]) : 0.0f;
  }
}
template <>
__device__ __forceinline__ void load_row32<float>(const float* p, int valid, float (&out)[32]) {
  if (valid >= 32) {
    const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 v = __ldg(q + i);
      out[i * 4] = v.x;
      out[i * 4 + 1] = v.y;
      out[i * 4 + 2] = v.z;
      out[i * 4 + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) out[i] = i < valid ? p[i] : 0.0f;
  }
}
__device__ __forceinline__ void load_row32_any(const void* base, long long elem_off, int fp32, int valid,
                                               float (&out)[32]) {
  if (fp32)
    load_row32<float>(reinterpret_cast<const float*>(base) + elem_off, valid, out);
  else
    load_row32<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(base) + elem_off, valid, out);
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void store_row32_any(void* base, long long elem_off, int fp32, int valid,
                                                const float (&v)[32]) {
  if (fp32) {
    float* p = reinterpret_cast<float*>(base) + elem_off;
    if (valid >= 32) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        reinterpret_cast<float4*>(p)[i] = make_float4(v[i * 4], v[i * 4 + 1], v[i * 4 + 2], v[i * 4 + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < valid) p[i] = v[i];
    }
  } else {
    __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(base) + elem_off;
    if (valid >= 32) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        reinterpret_cast<uint4*>(p)[i] =
            make_uint4(pack_bf16(v[i * 8], v[i * 8 + 1]), pack_bf16(v[i * 8 + 2], v[i * 8 + 3]),
                       pack_bf16(v[i * 8 + 4], v[i * 8 + 5]), pack_bf16(v[i * 8 + 6], v[i * 8 + 7]));
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < valid) p[i] = __float2bfloat16(v[i]);
    }
  }
}

struct TileCoord {
  int zl, zh, m0, n0;
};
template <int kCtas>
__device__ __forceinline__ TileCoord decode_tile(const GemmKParams& p, int tile, int cta_rank) {
  int z = tile / p.tiles_per_batch;
  int t = tile - z * p.tiles_per_batch;
  TileCoord c;
  c.zh = z / p.z_lo;
  c.zl = z - c.zh * p.z_lo;
  int gsz = kGroupM * p.n_blocks;
  int g = t / gsz;
  int r = t - g * gsz;
  int first_m = g * kGroupM;
  int gm = min(p.m_blocks - first_m, kGroupM);
  int nb = r / gm;
  int mb = first_m + (r - nb * gm);
  c.m0 = mb * (kBlockM * kCtas) + cta_rank * kBlockM;
  c.n0 = nb * p.n_per_tile;
  return c;
}

// The single MMA-issuing thread.  Everything that does not change per instruction is hoisted: the descriptor
// high words, the LBO field and the per-UMMA address steps are computed once; per k-block the thread does one
// barrier wait, one fence, four (add, add, issue) and one commit.  (Measured: with descriptors rebuilt per MMA
// this thread took ~465 ns per k-block and bounded the whole GEMM, in 1-CTA and 2-CTA mode alike.)
template <int kBlockN, int kCtas, bool kTF32>
__device__ __forceinline__ void mma_issue_loop(const GemmKParams& p, uint8_t* smem, uint64_t* full_bar,
                                               uint64_t* empty_bar, uint64_t* tfull_bar, uint64_t* tempty_bar,
                                               uint32_t tmem_base, int tile0, int tile_step) {
  using Cfg = GemmCfg<kBlockN, kCtas>;
  const uint32_t idesc = umma_idesc(kTF32 ? 2u : 1u, p.a_mn, p.b_mn, kBlockM * kCtas, kBlockN);
  const uint32_t a_hi = ((static_cast<uint32_t>(p.a_sbo) >> 4) & 0x3FFFu) | (1u << 14) | (static_cast<uint32_t>(p.a_lt) << 29);
  const uint32_t b_hi = ((static_cast<uint32_t>(p.b_sbo) >> 4) & 0x3FFFu) | (1u << 14) | (static_cast<uint32_t>(p.b_lt) << 29);
  const uint32_t a_lo0 = ((static_cast<uint32_t>(p.a_lbo) >> 4) & 0x3FFFu) << 16;
  const uint32_t b_lo0 = ((static_cast<uint32_t>(p.b_lbo) >> 4) & 0x3FFFu) << 16;
  const uint32_t a_step = static_cast<uint32_t>(p.a_kadv) >> 4, b_step = static_cast<uint32_t>(p.b_kadv) >> 4;
  const uint32_t smem16 = smem_u32(smem) >> 4;
  const int k_blocks = p.k_blocks;
  int stage = 0;
  uint32_t phase = 0;
  int acc = 0;
  uint32_t acc_phase = 0;
  // The whole warp runs this loop (warp-uniform control flow, waits spin inside one asm block); one elected lane
  // issues.  Under `if (lane == 0)` ptxas wrapped every UTCHMMA / UTCBAR in an elect-and-loop sequence.
  for (int tile = tile0; tile < p.total_tiles; tile += tile_step) {
    mbar_wait_u(&tempty_bar[acc], acc_phase ^ 1);
    tc_fence_after();
    const uint32_t d_tmem = tmem_base + acc * kBlockN;
    uint32_t accum = 0;
    for (int kb = 0; kb < k_blocks; ++kb) {
      mbar_wait_u(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_lo = a_lo0 | (smem16 + stage * (Cfg::kStageBytes >> 4));
        const uint32_t b_lo = b_lo0 | (smem16 + stage * (Cfg::kStageBytes >> 4) + (kStageABytes >> 4));
        umma_issue<kCtas, kTF32>(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, accum);
        umma_issue<kCtas, kTF32>(d_tmem, a_lo + a_step, a_hi, b_lo + b_step, b_hi, idesc, 1u);
        umma_issue<kCtas, kTF32>(d_tmem, a_lo + 2 * a_step, a_hi, b_lo + 2 * b_step, b_hi, idesc, 1u);
        umma_issue<kCtas, kTF32>(d_tmem, a_lo + 3 * a_step, a_hi, b_lo + 3 * b_step, b_hi, idesc, 1u);
        umma_commit_n<kCtas>(&empty_bar[stage]);  // smem slot is free (in both CTAs of a pair) once these retire
        if (kb + 1 == k_blocks) umma_commit_n<kCtas>(&tfull_bar[acc]);  // accumulator complete -> epilogue (of both CTAs)
      }
      __syncwarp();
      accum = 1u;
      if (++stage == Cfg::kStages) {
        stage = 0;
        phase ^= 1;
      }
    }
    acc ^= 1;
    if (acc == 0) acc_phase ^= 1;
  }
}

// 168 registers (not the 255 a 256-thread CTA may take): one GEMM CTA per SM then leaves ~22 K registers, so the
// HBM-bound kernels that run beside it on other streams (the overlapped AdamW, the gradient exchange) can be co-resident
// instead of waiting for a whole SM.
template <int kBlockN, int kCtas>
__global__ void __maxnreg__(168) gemm_tcgen05_kernel(const __grid_constant__ GemmKParams p) {
  using Cfg = GemmCfg<kBlockN, kCtas>;
  const int cta_rank = kCtas == 2 ? (int)cluster_ctarank() : 0;
  const bool leader = cta_rank == 0;
  const int tile0 = blockIdx.x / kCtas, tile_step = gridDim.x / kCtas;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* staging = smem + Cfg::kStages * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + 4 * kStagingPerWarp);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::kStages;
  uint64_t* tfull_bar = bars + 2 * Cfg::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmD);
    if (p.dual) tma_prefetch_desc(&p.tmB2);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4 * kCtas);  // the leader collects the epilogue warps of both CTAs
    }
    mbar_fence_init();
    fence_proxy_async_smem();
  }
  if (warp == 2) {
    if (kCtas == 2)
      tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols);
    else
      tmem_alloc(tmem_slot, Cfg::kTmemCols);
  }
  tc_fence_before();
  if (kCtas == 2)
    cluster_sync_all();  // peer barriers are initialised before any remote arrive / multicast commit
  else
    __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (whole warp, one elected lane issues)
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < p.total_tiles; tile += tile_step) {
        const TileCoord tc = decode_tile<kCtas>(p, tile, cta_rank);
        const int a_zbase = (tc.zl / p.a_div) * p.a_mul;
        const int b_zbase = (tc.zl / p.b_div) * p.b_mul;
        int seg = 0, kin = 0;  // K segment and k-block inside it (no divisions in the loop)
        // a CTA pair splits the B tile: rank r stages columns [n0 + r*kBlockN/2, ...) (dual: rank 0 = gate, 1 = up)
        const int bn0 = tc.n0 + ((kCtas == 2 && !p.dual) ? cta_rank * (kBlockN / 2) : 0);
        const void* tmB = (kCtas == 2 && p.dual && cta_rank == 1) ? &p.tmB2 : &p.tmB;
        const int a_atom0 = tc.m0 / p.atom_elems, b_atom0 = bn0 / p.atom_elems;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          const int k0 = kin * p.bk_elems;
          mbar_wait_u(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + kStageABytes;
          const int a_c2 = a_zbase + seg * p.a_seg;
          const int b_c2 = b_zbase + seg * p.b_seg;
          if (p.debug & 1) {
            if (leader) mbar_arrive(&full_bar[stage]);
          } else if constexpr (kCtas == 2) {
            // both CTAs' bytes complete on the leader's barrier; only the leader arms it
            const uint32_t lbar = smem_u32(&full_bar[stage]) & kPeerBitMask;
            if (leader) mbar_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
            if (!p.a_mn)
              tma_load_4d_2sm(sa, &p.tmA, lbar, k0, tc.m0, a_c2, tc.zh);
            else
              tma_load_5d_2sm(sa, &p.tmA, lbar, 0, k0, a_atom0, a_c2, tc.zh);
            if (!p.b_mn)
              tma_load_4d_2sm(sb, tmB, lbar, k0, bn0, b_c2, tc.zh);
            else
              tma_load_5d_2sm(sb, tmB, lbar, 0, k0, b_atom0, b_c2, tc.zh);
          } else {
            mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            if (!p.a_mn) {
              tma_load_4d(sa, &p.tmA, &full_bar[stage], k0, tc.m0, a_c2, tc.zh);
            } else if (p.a_5d) {
              tma_load_5d(sa, &p.tmA, &full_bar[stage], 0, k0, a_atom0, a_c2, tc.zh);
            } else {
              const int atoms = kBlockM / p.atom_elems;
              for (int i = 0; i < atoms; ++i)
                tma_load_4d(sa + i * p.atom_bytes, &p.tmA, &full_bar[stage], tc.m0 + i * p.atom_elems, k0, a_c2,
                            tc.zh);
            }
            if (p.dual) {
              tma_load_4d(sb, &p.tmB, &full_bar[stage], k0, tc.n0, b_c2, tc.zh);
              tma_load_4d(sb + 128 * 128, &p.tmB2, &full_bar[stage], k0, tc.n0, b_c2, tc.zh);
            } else if (!p.b_mn) {
              tma_load_4d(sb, &p.tmB, &full_bar[stage], k0, tc.n0, b_c2, tc.zh);
            } else if (p.b_5d) {
              tma_load_5d(sb, &p.tmB, &full_bar[stage], 0, k0, b_atom0, b_c2, tc.zh);
            } else {
              const int atoms = kBlockN / p.atom_elems;
              for (int i = 0; i < atoms; ++i)
                tma_load_4d(sb + i * p.atom_bytes, &p.tmB, &full_bar[stage], tc.n0 + i * p.atom_elems, k0, b_c2,
                            tc.zh);
            }
          }
          }
          __syncwarp();
          if (++kin == p.kbps) {
            kin = 0;
            ++seg;
          }
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    if (leader) {
      if (p.ab_fp32)
        mma_issue_loop<kBlockN, kCtas, true>(p, smem, full_bar, empty_bar, tfull_bar, tempty_bar, tmem_base, tile0,
                                             tile_step);
      else
        mma_issue_loop<kBlockN, kCtas, false>(p, smem, full_bar, empty_bar, tfull_bar, tempty_bar, tmem_base, tile0,
                                              tile_step);
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue
    const int w = warp - 4;  // == warp % 4: the TMEM lane quarter this warp may read
    uint8_t* stg = staging + w * kStagingPerWarp;
    int buf = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int chunk_cols = p.d_fp32 ? 32 : 64;
    const int sw = lane & 7;
    for (int tile = tile0; tile < p.total_tiles; tile += tile_step) {
      const TileCoord tc = decode_tile<kCtas>(p, tile, cta_rank);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int row = tc.m0 + w * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t trow = tmem_base + acc * kBlockN + (static_cast<uint32_t>(w * 32) << 16);
      const long long res_row = tc.zh * p.res_s3 + tc.zl * p.res_s2 + static_cast<long long>(row) * p.res_ld;
      const long long aux_row = tc.zh * p.aux_s3 + tc.zl * p.aux_s2 + static_cast<long long>(row) * p.aux_ld;

      // glu_bwd: raw g / u of the NEXT 32 columns, requested one half-chunk ahead (a thread reads its own row: the
      // loads are latency-, not bandwidth-bound, and would otherwise sit on the epilogue's critical path)
      uint4 gq[4], uq[4];
      const long long glu_row = static_cast<long long>(row) * p.glu_ld;
      auto glu_issue = [&](int col0) {
        const int n = tc.n0 + col0;
        if (row_ok && p.N - n >= 32) {
          const uint4* gp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.glu_g) + glu_row + n);
          const uint4* up = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.glu_u) + glu_row + n);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            gq[i] = __ldg(gp + i);
            uq[i] = __ldg(up + i);
          }
        }
      };
      if (p.glu_bwd) glu_issue(0);

      for (int c = 0; c < p.n_per_tile; c += chunk_cols) {
        if (tc.n0 + c >= p.N) break;  // warp-uniform: fully out-of-range chunk
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
        uint8_t* sbuf = stg + buf * (32 * 128) + lane * 128;
        for (int h = 0; h < chunk_cols / 32; ++h) {
          const int col0 = c + h * 32;
          const int n = tc.n0 + col0;
          const int valid = p.N - n;  // may be <= 0 or >= 32
          uint32_t r[32];
          float v[32];
          tmem_ld_32x32(trow + col0, r);
          if (p.dual) {
            // act(gate) * up with the glu kernels' arithmetic (vec.cuh act_only: one ex2 + one rcp per element)
            uint32_t r2[32];
            tmem_ld_32x32(trow + 128 + col0, r2);
            tmem_ld_wait();
            float g[32], u[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              g[j] = __uint_as_float(r[j]) * p.alpha;
              u[j] = __uint_as_float(r2[j]) * p.alpha;
            }
            if (p.aux != nullptr && row_ok && valid > 0) {
              store_row32_any(p.aux, aux_row + n, p.d_fp32, valid, g);
              store_row32_any(p.aux2, aux_row + n, p.d_fp32, valid, u);
            }
            if (p.act == B200_ACT_SILU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = act_only<4>(g[j], 4) * u[j];
            } else if (p.act == B200_ACT_GELU_TANH) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = act_only<2>(g[j], 2) * u[j];
            } else {
              act_fwd_n(g, p.act);
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = g[j] * u[j];
            }
          } else if (p.glu_bwd) {
            // accumulator = dh (gradient of act(g) * u): D <- dg = dh * u * act'(g), d2 <- du = dh * act(g)
            float g[32], u[32];
            const long long grow = glu_row + n;
            if (row_ok && valid >= 32) {          // the prefetched raw packs of this half
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const __nv_bfloat162* gh = reinterpret_cast<const __nv_bfloat162*>(&gq[i]);
                const __nv_bfloat162* uh = reinterpret_cast<const __nv_bfloat162*>(&uq[i]);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 gf = __bfloat1622float2(gh[e]), uf = __bfloat1622float2(uh[e]);
                  g[i * 8 + 2 * e] = gf.x;
                  g[i * 8 + 2 * e + 1] = gf.y;
                  u[i * 8 + 2 * e] = uf.x;
                  u[i * 8 + 2 * e + 1] = uf.y;
                }
              }
            } else if (row_ok && valid > 0) {     // ragged last columns
              load_row32_any(p.glu_g, grow, 0, valid, g);
              load_row32_any(p.glu_u, grow, 0, valid, u);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) g[j] = u[j] = 0.0f;
            }
            if (col0 + 32 < p.n_per_tile) glu_issue(col0 + 32);
            tmem_ld_wait();
            if (p.act == B200_ACT_SILU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float dh = __uint_as_float(r[j]) * p.alpha;
                float f, df;
                act_pair<4>(g[j], 4, f, df);
                v[j] = dh * u[j] * df;
                u[j] = dh * f;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float dh = __uint_as_float(r[j]) * p.alpha;
                float f, df;
                act_pair<2>(g[j], 2, f, df);
                v[j] = dh * u[j] * df;
                u[j] = dh * f;
              }
            }
            if (row_ok && valid > 0) store_row32_any(p.d2, grow, 0, valid, u);
          } else {
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
            if (p.bias != nullptr && valid > 0) {
              float bv[32];
              load_row32_any(p.bias, n, p.bias_fp32, valid, bv);
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += bv[j];
            }
            if (p.aux != nullptr && row_ok && valid > 0) store_row32_any(p.aux, aux_row + n, p.d_fp32, valid, v);
            if (p.act != B200_ACT_NONE) act_fwd_n(v, p.act);     // ONE dispatch per 32 columns (see vec.cuh)
          }
          if (p.res != nullptr && row_ok && valid > 0) {
            float rv[32];
            load_row32_any(p.res, res_row + n, p.res_fp32, valid, rv);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += rv[j];
          }
          // registers -> 128B-swizzled staging row (16-byte chunk index XOR row%8)
          if (p.d_fp32) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<float4*>(sbuf + ((q ^ sw) << 4)) =
                  make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              *reinterpret_cast<uint4*>(sbuf + (((h * 4 + q) ^ sw) << 4)) =
                  make_uint4(pack_bf16(v[q * 8], v[q * 8 + 1]), pack_bf16(v[q * 8 + 2], v[q * 8 + 3]),
                             pack_bf16(v[q * 8 + 4], v[q * 8 + 5]), pack_bf16(v[q * 8 + 6], v[q * 8 + 7]));
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && tc.m0 + w * 32 < p.M) {
          tma_store_4d(&p.tmD, stg + buf * (32 * 128), tc.n0 + c, tc.m0 + w * 32, tc.zl, tc.zh);
          tma_store_commit();
        }
        buf ^= 1;
      }
      // this warp's TMEM reads of the tile are complete -> hand the accumulator back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (kCtas == 2) mbar_arrive_leader(&tempty_bar[acc]); else mbar_arrive(&tempty_bar[acc]);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  if (kCtas == 2)
    cluster_sync_all();  // the peer's smem / barriers stay alive until the leader's last MMA and arrive are done
  else
    __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (kCtas == 2)
      tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
    else
      tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// dims/strides in ELEMENTS (stride of dim 0 is 1); box = (b0, b1, 1, 1); SWIZZLE_128B, or the
// 32-byte-atom variant (SWIZZLE_128B_ATOM_32B) that MN-major tf32 operands require.
static int encode_4d(CUtensorMap* m, int fp32, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3,
                     int64_t s1, int64_t s2, int64_t s3, uint32_t b0, uint32_t b1, const char* what,
                     bool atom32 = false) {
  EncodeTiledFn fn = get_encode_fn();
  B200_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
  const uint64_t es = fp32 ? 4 : 2;
  cuuint64_t dims[4] = {d0, d1, d2, d3};
  // unit dims still need a legal (multiple of 16 B, non-zero) stride
  if (s1 <= 0) s1 = (int64_t)((d0 + 15) / 16 * 16);
  if (s2 <= 0) s2 = s1 * (int64_t)d1;
  if (s3 <= 0) s3 = s2 * (int64_t)d2;
  cuuint64_t strides[3] = {(cuuint64_t)s1 * es, (cuuint64_t)s2 * es, (cuuint64_t)s3 * es};
  cuuint32_t box[4] = {b0, b1, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  B200_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "%s: base pointer not 16-byte aligned", what);
  for (int i = 0; i < 3; ++i)
    B200_CHECK(strides[i] % 16 == 0, "%s: stride %d (%llu bytes) not a multiple of 16", what, i + 1,
               (unsigned long long)strides[i]);
  CUresult r = fn(m, fp32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                  const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_CHECK(r == CUDA_SUCCESS,
             "%s: cuTensorMapEncodeTiled failed (%d) dims=(%llu,%llu,%llu,%llu) strides=(%llu,%llu,%llu) box=(%u,%u)",
             what, (int)r, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
             (unsigned long long)d3, (unsigned long long)strides[0], (unsigned long long)strides[1],
             (unsigned long long)strides[2], b0, b1);
  return 0;
}

// MN-major operand as a 5-D map: (atom_elems, K, MN/atom_elems, z2, z3) with box (atom, bk, n_atoms, 1, 1):
// the box lands in smem as [atom][k-row][128 B], i.e. exactly the canonical MN-major SWIZZLE_128B tile.
static int encode_mn_5d(CUtensorMap* m, int fp32, const void* ptr, uint64_t mn, uint64_t k, uint64_t d3, uint64_t d4,
                        int64_t ld, int64_t s3, int64_t s4, uint32_t atom_elems, uint32_t bk, uint32_t n_atoms,
                        const char* what, bool atom32) {
  EncodeTiledFn fn = get_encode_fn();
  B200_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
  const uint64_t es = fp32 ? 4 : 2;
  cuuint64_t dims[5] = {atom_elems, k, mn / atom_elems, d3, d4};
  if (s3 <= 0) s3 = ld * (int64_t)k;
  if (s4 <= 0) s4 = s3 * (int64_t)d3;
  cuuint64_t strides[4] = {(cuuint64_t)ld * es, (cuuint64_t)atom_elems * es, (cuuint64_t)s3 * es, (cuuint64_t)s4 * es};
  cuuint32_t box[5] = {atom_elems, bk, n_atoms, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  B200_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "%s: base pointer not 16-byte aligned", what);
  for (int i = 0; i < 4; ++i)
    B200_CHECK(strides[i] % 16 == 0, "%s: stride %d not a multiple of 16 bytes", what, i + 1);
  CUresult r = fn(m, fp32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5,
                  const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_CHECK(r == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled(5d) failed (%d)", what, (int)r);
  return 0;
}

// SMs the persistent GEMM may occupy (0 = all).  The data-parallel overlap lowers it while a collective's CTAs hold
// some SMs: a persistent grid sized for ALL SMs would leave its last CTAs queued behind the collective and run their
// statically assigned tiles as a second wave.
static int g_sm_limit = 0;

template <int kBlockN, int kCtas>
static int launch_gemm(const GemmKParams& kp, cudaStream_t stream) {
  using Cfg = GemmCfg<kBlockN, kCtas>;
  static unsigned long long configured = 0;      // one bit per device ordinal: the attribute is per device
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || !((configured >> dev) & 1ull)) {
    B200_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<kBlockN, kCtas>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   Cfg::kSmemBytes));
    if (dev < 64) configured |= 1ull << dev;
  }
  const int sms = (g_sm_limit > 0 && g_sm_limit < num_sms()) ? g_sm_limit : num_sms();
  int units = sms / kCtas;                             // CTAs, or CTA pairs
  if (kp.total_tiles < units) units = kp.total_tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(units * kCtas);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCtas;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  B200_CUDA(cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<kBlockN, kCtas>, kp));
  B200_LAUNCH_OK();
  return 0;
}

static bool use_cta_pairs() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200_GEMM_2CTA");
    v = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  return v == 1;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_set_gemm_sm_limit(int sms) {
  g_sm_limit = sms < 0 ? 0 : sms;
  return 0;
}

int b200_gemm(const b200_gemm_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK(a != nullptr, "b200_gemm: null args");
  B200_CHECK(a->m > 0 && a->n > 0 && a->k > 0, "b200_gemm: empty problem m=%lld n=%lld k=%lld", (long long)a->m,
             (long long)a->n, (long long)a->k);
  const int fp32 = a->ab_dtype == B200_F32;
  const int es = fp32 ? 4 : 2;
  const int z_lo = a->z_lo > 0 ? a->z_lo : 1, z_hi = a->z_hi > 0 ? a->z_hi : 1;
  const int k_segs = a->k_segs > 0 ? a->k_segs : 1;

  int bn = a->block_n;
  if (a->dual_b) {
    bn = 256;
    B200_CHECK(!a->a_mn_major && !a->b_mn_major && a->b2 != nullptr, "b200_gemm: dual_b needs K-major A, B, B2");
  } else if (bn == 0) {
    bn = a->n <= 64 ? 64 : (a->n <= 128 ? 128 : 256);
  }
  B200_CHECK(bn == 64 || bn == 128 || bn == 256, "b200_gemm: block_n must be 64/128/256");

  GemmKParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.M = (int)a->m;
  kp.N = (int)a->n;
  kp.bk_elems = 128 / es;
  kp.atom_elems = 128 / es;
  kp.atom_bytes = kp.bk_elems * 128;
  kp.n_per_tile = a->dual_b ? 128 : bn;
  // CTA pairs (256-row tiles) for the large GEMMs: every MN-major operand must be 5-D describable
  const int atom_e = 128 / es;
  // (256 x 128 pair tiles were measured at 0.6x the 256 x 256 rate on the N = 3584 shapes: operand traffic per FLOP is
  // 1.5x, the kernel turns shared-memory bound; the partial last wave of 256-wide tiles costs less)
  const bool pair = use_cta_pairs() && bn == 256 && a->m >= 1024 && (!a->a_mn_major || a->m % atom_e == 0) &&
                    (!a->b_mn_major || a->n % atom_e == 0) && a->block_n >= 0;
  kp.m_blocks = (int)ceil_div(a->m, pair ? 2 * kBlockM : kBlockM);
  kp.n_blocks = (int)ceil_div(a->n, kp.n_per_tile);
  kp.tiles_per_batch = kp.m_blocks * kp.n_blocks;
  kp.total_tiles = kp.tiles_per_batch * z_lo * z_hi;
  kp.z_lo = z_lo;
  kp.kbps = (int)ceil_div(a->k, kp.bk_elems);
  kp.k_blocks = kp.kbps * k_segs;
  kp.a_div = a->a_div > 0 ? a->a_div : 1;
  kp.a_mul = a->a_mul;
  kp.a_seg = a->a_seg;
  kp.b_div = a->b_div > 0 ? a->b_div : 1;
  kp.b_mul = a->b_mul;
  kp.b_seg = a->b_seg;
  if (z_lo == 1 && a->a_mul == 0 && a->a_div <= 1) kp.a_mul = 1;
  if (z_lo == 1 && a->b_mul == 0 && a->b_div <= 1) kp.b_mul = 1;
  kp.a_mn = a->a_mn_major ? 1 : 0;
  kp.b_mn = a->b_mn_major ? 1 : 0;
  // K-major: advance 32 B inside the swizzle row; LBO unused (16).  MN-major: advance UMMA_K rows of 128 B,
  // LBO = distance between 128-byte MN atoms = one atom (bk_elems rows x 128 B).
  kp.a_kadv = kp.a_mn ? (32 / es) * 128 : 32;
  kp.b_kadv = kp.b_mn ? (32 / es) * 128 : 32;
  kp.a_lbo = kp.a_mn ? kp.atom_bytes : 16;
  kp.b_lbo = kp.b_mn ? kp.atom_bytes : 16;
  // MN-major tf32: SWIZZLE_128B_BASE32B atoms are 128 B x 4 k-rows -> SBO (next k-group) = 512 B
  const bool a_atom32 = kp.a_mn && fp32, b_atom32 = kp.b_mn && fp32;
  kp.a_sbo = a_atom32 ? 512 : 1024;
  kp.b_sbo = b_atom32 ? 512 : 1024;
  kp.a_lt = a_atom32 ? 1 : 2;
  kp.b_lt = b_atom32 ? 1 : 2;
  kp.ab_fp32 = fp32;
  kp.d_fp32 = a->d_dtype == B200_F32;
  kp.dual = a->dual_b ? 1 : 0;
  kp.glu_bwd = a->glu_bwd ? 1 : 0;
  kp.glu_g = a->glu_g;
  kp.glu_u = a->glu_u;
  kp.d2 = a->d2;
  kp.glu_ld = a->glu_ld;
  if (kp.glu_bwd) {
    B200_CHECK(!kp.dual && a->glu_g != nullptr && a->glu_u != nullptr && a->d2 != nullptr && a->d_dtype == B200_BF16 &&
                   a->residual == nullptr && a->bias == nullptr && a->aux == nullptr && z_lo == 1 && z_hi == 1,
               "b200_gemm: glu_bwd needs bf16 g / u / d2 and a plain (unbatched, no bias / residual / aux) GEMM");
    B200_CHECK(a->act == B200_ACT_SILU || a->act == B200_ACT_GELU_TANH, "b200_gemm: glu_bwd supports SiLU and tanh-GELU");
    B200_CHECK(a->glu_ld % 8 == 0 && ((reinterpret_cast<uintptr_t>(a->glu_g) | reinterpret_cast<uintptr_t>(a->glu_u) |
                                       reinterpret_cast<uintptr_t>(a->d2)) & 15) == 0,
               "b200_gemm: glu_bwd operands must be 16-byte aligned rows");
  }
  kp.alpha = a->alpha == 0.0f ? 1.0f : a->alpha;
  kp.bias = a->bias;
  kp.bias_fp32 = a->bias_dtype == B200_F32;
  kp.res = a->residual;
  kp.res_fp32 = a->res_dtype == B200_F32;
  kp.res_ld = a->res_ld;
  kp.res_s2 = a->res_s2;
  kp.res_s3 = a->res_s3;
  kp.aux = a->aux;
  kp.aux2 = a->aux2;
  kp.aux_ld = a->aux_ld;
  kp.aux_s2 = a->d_s2;
  kp.aux_s3 = a->d_s3;
  kp.act = a->act;
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* e = getenv("B200_GEMM_DEBUG");
      dbg = e ? atoi(e) : 0;
    }
    kp.debug = dbg;
  }
  if (kp.dual) B200_CHECK((kp.aux == nullptr) == (kp.aux2 == nullptr), "b200_gemm: dual_b aux and aux2 go together");
  if (kp.res) B200_CHECK(a->res_ld % (16 / (kp.res_fp32 ? 4 : 2)) == 0, "b200_gemm: residual ld not 16B-multiple");
  if (kp.aux) B200_CHECK(a->aux_ld % (16 / (kp.d_fp32 ? 4 : 2)) == 0, "b200_gemm: aux ld not 16B-multiple");

  const uint32_t a_z2 = a->a_z2 > 0 ? a->a_z2 : 1, b_z2 = a->b_z2 > 0 ? a->b_z2 : 1;
  // A
  if (!kp.a_mn) {
    if (encode_4d(&kp.tmA, fp32, a->a, a->k, a->m, a_z2, z_hi, a->a_ld, a->a_s2, a->a_s3, kp.bk_elems, kBlockM, "A"))
      return 1;
  } else if (a->m % kp.atom_elems == 0) {
    kp.a_5d = 1;
    if (encode_mn_5d(&kp.tmA, fp32, a->a, a->m, a->k, a_z2, z_hi, a->a_ld, a->a_s2, a->a_s3, kp.atom_elems,
                     kp.bk_elems, kBlockM / kp.atom_elems, "A(mn5)", a_atom32))
      return 1;
  } else {
    if (encode_4d(&kp.tmA, fp32, a->a, a->m, a->k, a_z2, z_hi, a->a_ld, a->a_s2, a->a_s3, kp.atom_elems, kp.bk_elems,
                  "A(mn)", a_atom32))
      return 1;
  }
  // B
  if (kp.dual) {
    if (encode_4d(&kp.tmB, fp32, a->b, a->k, a->n, 1, 1, a->b_ld, 0, 0, kp.bk_elems, 128, "B(gate)")) return 1;
    if (encode_4d(&kp.tmB2, fp32, a->b2, a->k, a->n, 1, 1, a->b_ld, 0, 0, kp.bk_elems, 128, "B2(up)")) return 1;
  } else if (!kp.b_mn) {
    if (encode_4d(&kp.tmB, fp32, a->b, a->k, a->n, b_z2, z_hi, a->b_ld, a->b_s2, a->b_s3, kp.bk_elems,
                  pair ? bn / 2 : bn, "B"))
      return 1;
  } else if (a->n % kp.atom_elems == 0) {
    kp.b_5d = 1;
    if (encode_mn_5d(&kp.tmB, fp32, a->b, a->n, a->k, b_z2, z_hi, a->b_ld, a->b_s2, a->b_s3, kp.atom_elems,
                     kp.bk_elems, (pair ? bn / 2 : bn) / kp.atom_elems, "B(mn5)", b_atom32))
      return 1;
  } else {
    if (encode_4d(&kp.tmB, fp32, a->b, a->n, a->k, b_z2, z_hi, a->b_ld, a->b_s2, a->b_s3, kp.atom_elems, kp.bk_elems,
                  "B(mn)", b_atom32))
      return 1;
  }
  // D: store boxes are 32 rows x 128 bytes
  if (encode_4d(&kp.tmD, kp.d_fp32, a->d, a->n, a->m, z_lo, z_hi, a->d_ld, a->d_s2, a->d_s3, kp.d_fp32 ? 32 : 64, 32,
                "D"))
    return 1;

  switch (bn) {
    case 64:
      return launch_gemm<64, 1>(kp, stream);
    case 128:
      return launch_gemm<128, 1>(kp, stream);
    default:
      return pair ? launch_gemm<256, 2>(kp, stream) : launch_gemm<256, 1>(kp, stream);
  }
}
