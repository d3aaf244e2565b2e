// Flash attention for sm_100a (bf16, head_dim <= 256, self-attention over a packed qkv buffer): forward and backward
// with the score block resident in TMEM — no [B, H, S, S] tensor exists at any point; the forward saves one fp32
// log-sum-exp per query row and the backward recomputes P from it.
//
//   flash_fwd_kernel<T>   one CTA works on T query tiles (128 rows each) that share a KV head.  Per 64-key block:
//                         S = Q K^T (tcgen05.mma -> TMEM, double buffered per tile), online softmax by one warpgroup per
//                         tile (one thread = one query row, no cross-thread exchange; the running maximum is only
//                         refreshed — and O rescaled in TMEM via tcgen05.ld/st — when it grows by more than 2^8),
//                         P -> bf16 -> 128B-swizzled smem, O += P V (tcgen05.mma, O stays in TMEM for the whole tile).
//                         T = 2 for head_dim <= 128 (two rows of work keep MUFU and the tensor pipe busy at the same
//                         time), T = 1 for head_dim 256 (TMEM: 2 x 64 score columns + 256 output columns).
//   flash_dq_kernel       CTA = (batch, head, query tile); per key block S and dP = dO V^T in TMEM, dS = scale * P *
//                         (dP - delta) -> smem, dQ += dS K accumulated in TMEM over all key blocks, written once.
//   flash_dkv_kernel      CTA = (batch, KV head, 128-key tile); loops over the G query heads of the group and the query
//                         blocks that see the tile.  Works on the transposed problem (S^T = K Q^T: TMEM lanes = keys) so
//                         that P^T / dS^T land in smem as K-major A operands: dV += P^T dO, dK += dS^T Q accumulate in
//                         TMEM over heads and query blocks — the GQA group reduction happens in the accumulator, no
//                         atomics, run-to-run deterministic.  head_dim 256 runs dV and dK as two items (TMEM budget).
//   attn_delta_kernel     delta[b,h,q] = rowsum(dO * O), the softmax-backward row term.
//
// Every operand tile is staged by TMA as 64-column atoms of [rows][128 B] (SWIZZLE_128B); the same smem image serves as
// a K-major operand (contraction over head_dim: S, dP) and as an MN-major operand (contraction over its rows: P V,
// dS K, P^T dO, dS^T Q), so K / V / Q / dO are loaded once per use site and never transposed.
//
// Mask rule (b200_softmax_fwd's): allowed(q, k) = k < S && (keymask == NULL || keymask[b, k]) &&
// (causal ? k <= q : true) && (bid == NULL || bid_k[b, k] <= bid_q[b, q]).
//
// Reference arithmetic replaced: F.scaled_dot_product_attention / eager softmax attention inside HF Qwen2 / Llama /
// Gemma / CLIP / SigLIP attention and its autograd backward (called from dexbotic_arch.py:55-62, clip_encoder.py:50-54,
// siglip_encoder.py:79-84; pi0's joint attention pi0_arch.py:185-192 with make_attn_mask pi0_arch.py:22-33).
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/dexbotic_b200_ops.h"
#include "common.h"
#define B200_SPIN_LIMIT (1u << 24)   // a protocol dead-lock traps within seconds instead of hanging the device
#include "ptx.cuh"

namespace b200 {

using bf16 = __nv_bfloat16;

constexpr int kFaMaxStages = 4;
constexpr int kAtom128 = 128 * 128;   // bytes of one 64-column atom of a 128-row tile
constexpr int kAtom64 = 64 * 128;     // ... of a 64-row block
constexpr float kRescaleThreshold = 8.0f;   // log2 units: P <= 2^8 under a stale maximum

struct FaParams {
  CUtensorMap tmQ, tmK, tmV, tmDO;
  CUtensorMap tmO;      // fwd: out rows, dq: dq rows (box 64 columns x 32 rows: one warp's slice of an atom)
  int B, H, KVH, G, S, hd;
  int katoms;    // ceil(hd / 64): 64-column atoms per tile row
  int ksteps;    // ceil(hd / 16): UMMA K steps of a head_dim contraction
  int n_hd;      // hd rounded up to 16: MMA N of the products whose output is head_dim wide
  int o_chunks;  // ceil(hd / 32): 32-column TMEM chunks of such an output
  int m_tiles;   // ceil(S / 128)
  int n_qblk;    // ceil(S / 64)
  int stages;
  int causal;
  int n_items;
  int n_modes;   // dkv: 1 (dV and dK together) or 2 (head_dim > 128: separate items)
  int tma_epi;   // dq: output rows leave through smem staging + TMA (needs 32 KB of shared memory)
  float scale, sl2;
  const uint8_t* keymask;
  const int* bid_q;
  const int* bid_k;
  const int4* mask_ws;  // [B][n_qblk][2] summary of the key / query side masks (attn_mask_prep_kernel), NULL: no such masks
  bf16* out;            // fwd: O rows; bwd: unused
  float* lse;           // [B, H, S] log2-domain log-sum-exp
  const float* delta;   // [B, H, S]
  bf16 *dq, *dk, *dv;
  long long o_ld, o_sh, o_sb;      // element strides of out (row, head, batch)
  long long g_ld, gq_sh, gkv_sh, g_sb;   // strides of dq / dk / dv
  long long* trace;     // debugging: clock64() stamps of CTA 0 (b200_flash_attn_set_trace), NULL in production
};

// slot-major [slot][64] table of clock64() stamps taken by CTA 0 for its first 64 steps (tools/trace_flash.py)
#define FA_TRACE(slot, idx)                                                                      \
  do {                                                                                           \
    if (p.trace != nullptr && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (idx) < 64u)        \
      p.trace[(slot) * 64 + (idx)] = clock64();                                                  \
  } while (0)


__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// 0xffff in the halves of a packed bf16 pair whose bits (c, c + 1) of `bits` are set
__device__ __forceinline__ uint32_t pair_mask(uint32_t bits, int c) {
  return (((bits >> c) & 1u) ? 0x0000ffffu : 0u) | (((bits >> (c + 1)) & 1u) ? 0xffff0000u : 0u);
}
__device__ __forceinline__ uint32_t low_mask(int n) { return n >= 32 ? 0xffffffffu : (n <= 0 ? 0u : ((1u << n) - 1u)); }

constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO 1024, descriptor version 1, SWIZZLE_128B
constexpr uint32_t kLoKMajor = (16u >> 4) << 16;                        // LBO unused for K-major operands
constexpr uint32_t kLoMN64 = (8192u >> 4) << 16;                        // MN-major, 64-row atoms: next atom 8 KB on

// One tcgen05.mma (bf16, fp32 accumulate in TMEM) with a compile-time accumulate flag: the issuing thread does two
// adds and the issue per instruction (both operands share the descriptor high word).
template <bool kAcc>
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 da, db;\nmov.b64 da, {%1, %3};\nmov.b64 db, {%2, %3};\nsetp.ne.b32 p, %5, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(kDescHi), "r"(idesc), "n"(kAcc ? 1 : 0)
      : "memory");
}
// D = A B^T contracted over head_dim (overwrites D); A / B tiles are [rows][hd] atom images, used K-major.
// a_lo / b_lo: descriptor low words of the tiles' first atoms; *_atom16: atom size in 16-byte units.
__device__ __forceinline__ void mma_over_hd(uint32_t d_tmem, uint32_t a_lo, uint32_t a_atom16, uint32_t b_lo,
                                            uint32_t b_atom16, int ksteps, uint32_t idesc) {
  umma_bf16<false>(d_tmem, a_lo, b_lo, idesc);
  int ks = 1;
  for (; ks < 4 && ks < ksteps; ++ks) umma_bf16<true>(d_tmem, a_lo + 2u * ks, b_lo + 2u * ks, idesc);
  for (; ks + 4 <= ksteps; ks += 4) {
    a_lo += a_atom16;
    b_lo += b_atom16;
    umma_bf16<true>(d_tmem, a_lo, b_lo, idesc);
    umma_bf16<true>(d_tmem, a_lo + 2u, b_lo + 2u, idesc);
    umma_bf16<true>(d_tmem, a_lo + 4u, b_lo + 4u, idesc);
    umma_bf16<true>(d_tmem, a_lo + 6u, b_lo + 6u, idesc);
  }
  if (ks < ksteps) {          // partial last atom (head_dim 72 / 96 ...)
    a_lo += a_atom16;
    b_lo += b_atom16;
    for (int r = 0; ks < ksteps; ++ks, ++r) umma_bf16<true>(d_tmem, a_lo + 2u * r, b_lo + 2u * r, idesc);
  }
}
// D (+)= A B contracted over the 64 rows of B: A = one [128][64] K-major atom (P, dS, P^T, dS^T), B = a 64-row block
// [64][hd] read MN-major (N = head_dim).  a_lo: K-major low word of A; b_lo: MN-major low word of B.
__device__ __forceinline__ void mma_over_rows64(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                                bool accumulate) {
  if (accumulate)
    umma_bf16<true>(d_tmem, a_lo, b_lo, idesc);
  else
    umma_bf16<false>(d_tmem, a_lo, b_lo, idesc);
  umma_bf16<true>(d_tmem, a_lo + 2u, b_lo + 128u, idesc);
  umma_bf16<true>(d_tmem, a_lo + 4u, b_lo + 256u, idesc);
  umma_bf16<true>(d_tmem, a_lo + 6u, b_lo + 384u, idesc);
}
__device__ __forceinline__ uint32_t lo_kmajor(uint32_t addr) { return kLoKMajor | (addr >> 4); }
__device__ __forceinline__ uint32_t lo_mn64(uint32_t addr) { return kLoMN64 | (addr >> 4); }

__device__ __forceinline__ int visible_blocks(const FaParams& p, int tile) {
  const int keys = p.causal ? min(p.S, tile * 128 + 128) : p.S;
  return (keys + 63) >> 6;
}

// Mask summary per (batch row, 64-key block), written by attn_mask_prep_kernel once per call, so that the softmax warps
// read ONE 16-byte word per block instead of evaluating the mask inputs key by key:
//   entry 0 (key side):   x / y = validity bits (in range and not padding) of keys [64j, 64j+32) / [64j+32, 64j+64),
//                         z / w = max / min block id (pi0 rule) over the valid keys (INT_MIN / INT_MAX when none)
//   entry 1 (query side): x / y = min / max block id of queries [64j, 64j+32), z / w = the same of [64j+32, 64j+64)
// Without a keymask and block ids the table is not needed (mask_ws == NULL): the bits follow from S alone.
__global__ void __launch_bounds__(128) attn_mask_prep_kernel(const uint8_t* __restrict__ keymask,
                                                             const int* __restrict__ bid_q,
                                                             const int* __restrict__ bid_k, int B, int S, int n_blk,
                                                             int4* __restrict__ ws) {
  const int w = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (w >= B * n_blk) return;
  const int b = w / n_blk, j = w - b * n_blk;
  int4 ke, qe;
  int kmax = (int)0x80000000, kmin = 0x7fffffff;
  uint32_t bits[2];
  int qmin[2], qmax[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int k = j * 64 + h * 32 + lane;
    const bool in = k < S;
    const bool ok = in && (keymask == nullptr || keymask[(size_t)b * S + k] != 0);
    bits[h] = __ballot_sync(0xffffffffu, ok);
    if (bid_k != nullptr && ok) {
      const int v = bid_k[(size_t)b * S + k];
      kmax = max(kmax, v);
      kmin = min(kmin, v);
    }
    int lo = 0x7fffffff, hi = (int)0x80000000;
    if (bid_q != nullptr && in) lo = hi = bid_q[(size_t)b * S + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    qmin[h] = lo;
    qmax[h] = hi;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
  }
  if (lane == 0) {
    ke.x = (int)bits[0];
    ke.y = (int)bits[1];
    ke.z = kmax;
    ke.w = kmin;
    qe.x = qmin[0];
    qe.y = qmax[0];
    qe.z = qmin[1];
    qe.w = qmax[1];
    ws[(size_t)w * 2] = ke;
    ws[(size_t)w * 2 + 1] = qe;
  }
}

// Read-only global loads that are ISSUED where they are written: a plain (invariant) load is sunk by the compiler to its
// first use — a whole pipeline step later — which exposes the L2 / DRAM latency these prefetches are meant to hide.
__device__ __forceinline__ int4 ldg_now(const int4* ptr) {
  int4 r;
  asm volatile("ld.global.nc.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(ptr));
  return r;
}
__device__ __forceinline__ float ldg_now(const float* ptr) {
  float r;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(r) : "l"(ptr));
  return r;
}
__device__ __forceinline__ int ldg_now(const int* ptr) {
  int r;
  asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(r) : "l"(ptr));
  return r;
}
// key-side entry of block j (registers only when there is no table)
__device__ __forceinline__ int4 load_kblk(const FaParams& p, int b, int j) {
  if (p.mask_ws != nullptr) return ldg_now(p.mask_ws + ((size_t)b * p.n_qblk + j) * 2);
  const int n = p.S - j * 64;
  return make_int4((int)low_mask(n), (int)low_mask(n - 32), (int)0x80000000, 0x7fffffff);
}
// Validity bits of the 64 keys of block j for query row q (row-wise kernels: fwd, dq).  The pi0 block-id rule only
// costs a per-key evaluation in blocks whose valid keys straddle the row's block id.
__device__ __forceinline__ void row_mask(const FaParams& p, const int4 ki, int b, int q, int bq, int k0, uint32_t& lo,
                                         uint32_t& hi) {
  lo = (uint32_t)ki.x;
  hi = (uint32_t)ki.y;
  if (p.causal) {
    lo &= low_mask(q + 1 - k0);
    hi &= low_mask(q - 31 - k0);
  }
  if (p.bid_k != nullptr) {
    // ki.z / ki.w = max / min block id over the block's valid keys: rows at or past the max see all of them, rows
    // before the min none; only a block whose keys straddle some row's id is evaluated key by key (whole warp: the keys'
    // ids are loaded once, coalesced, and broadcast by shuffles)
    const bool mixed = bq < ki.z && bq >= ki.w;
    if (__any_sync(0xffffffffu, mixed)) {
      const int lane = threadIdx.x & 31;
      const int* bk = p.bid_k + (size_t)b * p.S + k0;
      const int v0 = k0 + lane < p.S ? bk[lane] : 0x7fffffff;
      const int v1 = k0 + 32 + lane < p.S ? bk[32 + lane] : 0x7fffffff;
      uint32_t w0 = 0u, w1 = 0u;
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        if (__shfl_sync(0xffffffffu, v0, c) <= bq) w0 |= 1u << c;
        if (__shfl_sync(0xffffffffu, v1, c) <= bq) w1 |= 1u << c;
      }
      lo &= w0;
      hi &= w1;
    } else if (bq < ki.w) {
      lo = hi = 0u;
    }
  }
}

// ---------------------------------------------------------------------------------------------------- forward
struct FwdItem {
  int b, kvh, n;
  int h[2], tile[2], nblk[2];
};
template <int kTiles>
__device__ __forceinline__ FwdItem fwd_item(const FaParams& p, int it) {
  FwdItem r;
  if (kTiles == 2) {
    const int BK = p.B * p.KVH;
    const int pr = it / BK, bk = it - pr * BK;
    r.b = bk / p.KVH;
    r.kvh = bk - r.b * p.KVH;
    const int entries = p.G * p.m_tiles;
    r.n = 0;
    for (int t = 0; t < 2; ++t) {
      const int e = 2 * pr + t;
      if (e < entries) {
        const int tq = e / p.G;
        r.tile[t] = p.m_tiles - 1 - tq;          // widest (most key blocks under a causal mask) first
        r.h[t] = r.kvh * p.G + (e - tq * p.G);
        r.nblk[t] = visible_blocks(p, r.tile[t]);
      } else {
        r.tile[t] = 0;
        r.h[t] = 0;
        r.nblk[t] = 0;
      }
      r.n = max(r.n, r.nblk[t]);
    }
  } else {
    const int BH = p.B * p.H;
    const int pr = it / BH, bh = it - pr * BH;
    r.b = bh / p.H;
    r.h[0] = bh - r.b * p.H;
    r.kvh = r.h[0] / p.G;
    r.tile[0] = p.m_tiles - 1 - pr;
    r.nblk[0] = visible_blocks(p, r.tile[0]);
    r.n = r.nblk[0];
    r.h[1] = r.tile[1] = r.nblk[1] = 0;
  }
  return r;
}

// store one 32-column TMEM chunk of an output row as bf16 (columns >= hd are not written)
__device__ __forceinline__ void store_chunk_bf16(bf16* row, int c, int hd, const uint32_t (&o)[32], float mul) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int col = c * 32 + g * 8;
    if (col < hd) {
      uint4 u;
      u.x = pack2(__uint_as_float(o[g * 8]) * mul, __uint_as_float(o[g * 8 + 1]) * mul);
      u.y = pack2(__uint_as_float(o[g * 8 + 2]) * mul, __uint_as_float(o[g * 8 + 3]) * mul);
      u.z = pack2(__uint_as_float(o[g * 8 + 4]) * mul, __uint_as_float(o[g * 8 + 5]) * mul);
      u.w = pack2(__uint_as_float(o[g * 8 + 6]) * mul, __uint_as_float(o[g * 8 + 7]) * mul);
      *reinterpret_cast<uint4*>(row + col) = u;
    }
  }
}
// rows of a [128 x hd] fp32 TMEM accumulator -> bf16 global rows; chunks c0, c0 + cstep, ... two TMEM loads in flight
__device__ __forceinline__ void store_acc_rows(uint32_t acc_tmem, bf16* row, bool row_ok, int c0, int cstep, int o_chunks,
                                               int hd, float mul) {
  for (int c = c0; c < o_chunks; c += 2 * cstep) {
    uint32_t oa[32], ob[32];
    const bool two = c + cstep < o_chunks;
    tmem_ld_32x32(acc_tmem + c * 32, oa);
    if (two) tmem_ld_32x32(acc_tmem + (c + cstep) * 32, ob);
    tmem_ld_wait();
    if (row_ok) {
      store_chunk_bf16(row, c, hd, oa, mul);
      if (two) store_chunk_bf16(row, c + cstep, hd, ob, mul);
    }
  }
}
// The same through shared memory and TMA: one warp's 32 rows of a [128 x hd] fp32 accumulator -> bf16 -> the warp's 4 KB
// slice of a 128B-swizzled [128][128 B] atom image -> cp.async.bulk.tensor store (full 128-byte lines; rows >= S and
// columns >= hd are clipped by the tensor map).  Atoms a0, a0 + astep, ...  `stage` must be 1024-byte aligned.
__device__ __forceinline__ void store_acc_tma(const CUtensorMap* tm, uint32_t acc_tmem, uint8_t* stage, int lane, int a0,
                                              int astep, int katoms, int o_chunks, float mul, int row0, int c2, int c3,
                                              bool any_row) {
  uint8_t* srow = stage + lane * 128;
  const int sw = lane & 7;
  for (int a = a0; a < katoms; a += astep) {
    uint32_t oa[32], ob[32];
    const bool two = 2 * a + 1 < o_chunks;
    tmem_ld_32x32(acc_tmem + a * 64, oa);
    if (two) tmem_ld_32x32(acc_tmem + a * 64 + 32, ob);
    tmem_ld_wait();
    if (lane == 0) tma_store_wait_read<0>();     // the previous store out of this slice has been read
    __syncwarp();
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint4 u;
      u.x = pack2(__uint_as_float(oa[g * 8]) * mul, __uint_as_float(oa[g * 8 + 1]) * mul);
      u.y = pack2(__uint_as_float(oa[g * 8 + 2]) * mul, __uint_as_float(oa[g * 8 + 3]) * mul);
      u.z = pack2(__uint_as_float(oa[g * 8 + 4]) * mul, __uint_as_float(oa[g * 8 + 5]) * mul);
      u.w = pack2(__uint_as_float(oa[g * 8 + 6]) * mul, __uint_as_float(oa[g * 8 + 7]) * mul);
      *reinterpret_cast<uint4*>(srow + ((g ^ sw) << 4)) = u;
    }
    if (two) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 u;
        u.x = pack2(__uint_as_float(ob[g * 8]) * mul, __uint_as_float(ob[g * 8 + 1]) * mul);
        u.y = pack2(__uint_as_float(ob[g * 8 + 2]) * mul, __uint_as_float(ob[g * 8 + 3]) * mul);
        u.z = pack2(__uint_as_float(ob[g * 8 + 4]) * mul, __uint_as_float(ob[g * 8 + 5]) * mul);
        u.w = pack2(__uint_as_float(ob[g * 8 + 6]) * mul, __uint_as_float(ob[g * 8 + 7]) * mul);
        *reinterpret_cast<uint4*>(srow + (((4 + g) ^ sw) << 4)) = u;
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0 && any_row) {
      tma_store_4d(tm, stage, a * 64, row0, c2, c3);
      tma_store_commit();
    }
  }
}
// One arrival per warp: every lane has fenced its own writes / TMEM reads, __syncwarp orders them before lane 0's
// (release) arrive.  256 threads arriving on one mbarrier serialise (~32 cycles per warp instruction on one address).
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void release(uint64_t* bar, bool used) {
  if (used)
    umma_commit(bar);     // arrives when this thread's MMAs (the readers of the buffer) have retired
  else
    mbar_arrive(bar);
}

// warp roles: 0 TMA producer | 1 MMA issuer of tile 0 | 2 TMEM allocator | 3 MMA issuer of tile 1 | 4.. softmax
template <int kTiles>
__global__ void __launch_bounds__(128 + 128 * kTiles, 1) flash_fwd_kernel(const __grid_constant__ FaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar_q_full, bar_q_empty, bar_kv_full[kFaMaxStages], bar_kv_empty[kFaMaxStages];
  __shared__ uint64_t bar_s_full[2][2], bar_p_full[2], bar_pv_done[2];
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* sQ = smem;                                   // kTiles x katoms x [128][128 B]
  uint8_t* sP = sQ + kTiles * p.katoms * kAtom128;      // kTiles x [128][128 B]
  uint8_t* sRing = sP + kTiles * kAtom128;              // stages x (K: katoms x [64][128 B] | V: the same)
  const int stage_bytes = 2 * p.katoms * kAtom64;

  if (warp == 1 && lane == 0) {
    mbar_init(&bar_q_full, 1);
    mbar_init(&bar_q_empty, kTiles);
    for (int i = 0; i < kFaMaxStages; ++i) {
      mbar_init(&bar_kv_full[i], 1);
      mbar_init(&bar_kv_empty[i], kTiles);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&bar_s_full[t][0], 1);
      mbar_init(&bar_s_full[t][1], 1);
      mbar_init(&bar_p_full[t], 4);        // one arrival per softmax warp
      mbar_init(&bar_pv_done[t], 1);
    }
    mbar_fence_init();
    fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
  }
  if (warp == 2) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int stages = p.stages;

  if (warp == 0) {
    // ------------------------------------------------ TMA producer: the whole warp runs the loop, one elected lane issues
    {
      int stage = 0, qi = 0;
      uint32_t phase = 0;
      for (int it = blockIdx.x; it < p.n_items; it += gridDim.x, ++qi) {
        const FwdItem im = fwd_item<kTiles>(p, it);
        mbar_wait_u(&bar_q_empty, (uint32_t)(qi & 1) ^ 1u);
        if (elect_one()) {
          int nvalid = 0;
          for (int t = 0; t < kTiles; ++t) nvalid += im.nblk[t] > 0 ? 1 : 0;
          mbar_expect_tx(&bar_q_full, (uint32_t)(nvalid * p.katoms * kAtom128));
          for (int t = 0; t < kTiles; ++t) {
            if (im.nblk[t] == 0) continue;
            for (int a = 0; a < p.katoms; ++a)
              tma_load_4d(sQ + (t * p.katoms + a) * kAtom128, &p.tmQ, &bar_q_full, a * 64, im.tile[t] * 128, im.h[t],
                          im.b);
          }
        }
        __syncwarp();
        for (int j = 0; j < im.n; ++j) {
          mbar_wait_u(&bar_kv_empty[stage], phase ^ 1u);
          if (elect_one()) {
            mbar_expect_tx(&bar_kv_full[stage], (uint32_t)stage_bytes);
            uint8_t* sK = sRing + stage * stage_bytes;
            uint8_t* sV = sK + p.katoms * kAtom64;
            for (int a = 0; a < p.katoms; ++a) {
              tma_load_4d(sK + a * kAtom64, &p.tmK, &bar_kv_full[stage], a * 64, j * 64, im.kvh, im.b);
              tma_load_4d(sV + a * kAtom64, &p.tmV, &bar_kv_full[stage], a * 64, j * 64, im.kvh, im.b);
            }
          }
          __syncwarp();
          if (++stage == stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1 || (kTiles == 2 && warp == 3)) {
    // -------------------------------------- MMA issuer of tile t: the two tiles of a CTA run decoupled from each other.
    // The whole warp runs the loop (warp-uniform control flow); one elected lane issues MMAs and commits.
    {
      const int t = warp == 1 ? 0 : 1;
      const uint32_t idesc_s = umma_idesc(1u, 0, 0, 128, 64);
      const uint32_t idesc_pv = umma_idesc(1u, 0, 1, 128, (uint32_t)p.n_hd);
      const uint32_t q_lo = lo_kmajor(smem_u32(sQ) + t * p.katoms * kAtom128);
      const uint32_t p_lo = lo_kmajor(smem_u32(sP) + t * kAtom128);
      const uint32_t ring_addr = smem_u32(sRing);
      const uint32_t s_tmem = tmem + t * 128, o_tmem = tmem + (kTiles == 2 ? 256u + t * 128u : 128u);
      uint32_t blk = 0;   // running block counter of this tile: S buffer = blk & 1
      int stage = 0, qi = 0;
      uint32_t phase = 0;
      for (int it = blockIdx.x; it < p.n_items; it += gridDim.x, ++qi) {
        const FwdItem im = fwd_item<kTiles>(p, it);
        const int nb = im.nblk[t], n = im.n;
        mbar_wait_u(&bar_q_full, (uint32_t)(qi & 1));
        if (nb > 0) {
          mbar_wait_u(&bar_kv_full[stage], phase);
          tc_fence_after();
        }
        if (elect_one()) {
          if (nb > 0) {
            mma_over_hd(s_tmem + (blk & 1u) * 64, q_lo, kAtom128 >> 4, lo_kmajor(ring_addr + stage * stage_bytes),
                        kAtom64 >> 4, p.ksteps, idesc_s);
            umma_commit(&bar_s_full[t][blk & 1u]);
          }
          if (n == 1) release(&bar_q_empty, nb > 0);
        }
        __syncwarp();
        for (int j = 0; j < n; ++j) {
          int ns = stage + 1;
          uint32_t nph = phase;
          if (ns == stages) {
            ns = 0;
            nph ^= 1u;
          }
          if (j + 1 < n) {   // the next block's scores first: the softmax warps never wait for the P V product
            if (j + 1 < nb) {
              mbar_wait_u(&bar_kv_full[ns], nph);
              tc_fence_after();
            }
            if (elect_one()) {
              if (j + 1 < nb) {
                mma_over_hd(s_tmem + ((blk + 1u) & 1u) * 64, q_lo, kAtom128 >> 4,
                            lo_kmajor(ring_addr + ns * stage_bytes), kAtom64 >> 4, p.ksteps, idesc_s);
                umma_commit(&bar_s_full[t][(blk + 1u) & 1u]);
              }
              if (j + 2 == n) release(&bar_q_empty, nb > 0);   // every S product of this item has been issued
            }
            __syncwarp();
          }
          if (j < nb) {
            if (t == 0) FA_TRACE(16, blk);
            mbar_wait_u(&bar_p_full[t], blk & 1u);
            tc_fence_after();
            if (t == 0) FA_TRACE(17, blk);
          } else {
            mbar_wait_u(&bar_kv_full[stage], phase);   // a stage this tile does not use: still no running ahead of the ring
          }
          if (elect_one()) {
            if (j < nb) {
              mma_over_rows64(o_tmem, p_lo, lo_mn64(ring_addr + stage * stage_bytes + p.katoms * kAtom64), idesc_pv,
                              j > 0);
              umma_commit(&bar_pv_done[t]);
            }
            release(&bar_kv_empty[stage], j < nb);
          }
          __syncwarp();
          if (j < nb) {
            if (t == 0) FA_TRACE(18, blk);
            ++blk;
          }
          stage = ns;
          phase = nph;
        }
      }
    }
  } else if (warp >= 4) {
    // ----------------------------------------------------- softmax warpgroup of tile t: one thread = one query row
    const int t = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
    const uint32_t o_tmem = lane_base + (kTiles == 2 ? 256u + t * 128u : 128u);
    uint8_t* pslice = sP + t * kAtom128 + quarter * 4096;    // this warp's 32 rows of the P tile (also its output stage)
    uint8_t* prow = pslice + lane * 128;
    const int sw = r & 7;
    uint32_t cnt = 0;
    for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
      const FwdItem im = fwd_item<kTiles>(p, it);
      const int nblk = im.nblk[t];
      if (nblk == 0) continue;
      const int b = im.b, h = im.h[t];
      const int q = im.tile[t] * 128 + r;
      const bool row_ok = q < p.S;
      const int bq = (p.bid_q != nullptr && row_ok) ? p.bid_q[(size_t)b * p.S + q] : 0;
      float m_ref = -INFINITY, l = 0.0f;
      int4 kin = load_kblk(p, b, 0);
      for (int j = 0; j < nblk; ++j) {
        const uint32_t buf = cnt & 1u;
        const int4 kic = kin;
        if (j + 1 < nblk) kin = load_kblk(p, b, j + 1);    // next block's mask word: in flight during this block
        uint32_t bits0, bits1;
        row_mask(p, kic, b, q, bq, j * 64, bits0, bits1);
        const bool full = __all_sync(0xffffffffu, (bits0 & bits1) == 0xffffffffu);
        if (warp == 4) FA_TRACE(19, cnt);
        mbar_wait(&bar_s_full[t][buf], (cnt >> 1) & 1u);
        tc_fence_after();
        if (warp == 4) FA_TRACE(20, cnt);
        uint32_t s0[32], s1[32];
        tmem_ld_32x32(lane_base + t * 128 + buf * 64, s0);
        tmem_ld_32x32(lane_base + t * 128 + buf * 64 + 32, s1);
        tmem_ld_wait();
        if (warp == 4) FA_TRACE(21, cnt);
        if (!full) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (!((bits0 >> i) & 1u)) s0[i] = 0xff800000u;   // -inf
            if (!((bits1 >> i) & 1u)) s1[i] = 0xff800000u;
          }
        }
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // four independent chains
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          mx4[0] = fmaxf(mx4[0], __uint_as_float(s0[i]));
          mx4[1] = fmaxf(mx4[1], __uint_as_float(s0[16 + i]));
          mx4[2] = fmaxf(mx4[2], __uint_as_float(s1[i]));
          mx4[3] = fmaxf(mx4[3], __uint_as_float(s1[16 + i]));
        }
        const float m_blk = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * p.sl2;
        if (j == 0) {
          m_ref = m_blk;
        } else {
          const bool need = m_blk > m_ref + kRescaleThreshold;
          if (__any_sync(0xffffffffu, need)) {
            // refresh the reference maximum: O and l shrink by 2^(m_ref - m_new) (rows that do not need it: x1)
            const float m_new = need ? m_blk : m_ref;
            const float alpha = need ? ex2f(m_ref - m_new) : 1.0f;
            mbar_wait(&bar_pv_done[t], (cnt - 1u) & 1u);
            tc_fence_after();
            for (int c = 0; c < p.o_chunks; ++c) {
              uint32_t o[32];
              tmem_ld_32x32(o_tmem + c * 32, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st_32x32(o_tmem + c * 32, o);
            }
            tmem_st_wait();
            l *= alpha;
            m_ref = m_new;
          }
        }
        const float m_use = m_ref == -INFINITY ? 0.0f : m_ref;
        if (warp == 4) FA_TRACE(22, cnt);
        uint32_t pk[32];
        float sum0 = 0.0f, sum1 = 0.0f, sum2 = 0.0f, sum3 = 0.0f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float a0 = ex2f(fmaf(__uint_as_float(s0[2 * i]), p.sl2, -m_use));
          const float a1 = ex2f(fmaf(__uint_as_float(s0[2 * i + 1]), p.sl2, -m_use));
          const float c0 = ex2f(fmaf(__uint_as_float(s1[2 * i]), p.sl2, -m_use));
          const float c1 = ex2f(fmaf(__uint_as_float(s1[2 * i + 1]), p.sl2, -m_use));
          sum0 += a0;
          sum1 += a1;
          sum2 += c0;
          sum3 += c1;
          pk[i] = pack2(a0, a1);
          pk[16 + i] = pack2(c0, c1);
        }
        l += (sum0 + sum1) + (sum2 + sum3);
        if (warp == 4) FA_TRACE(23, cnt);
        if (j == 0) {
          if (lane == 0) tma_store_wait_read<0>();     // the previous item's output has left this slice of sP
          __syncwarp();
        } else {
          mbar_wait(&bar_pv_done[t], (cnt - 1u) & 1u);   // the P buffer's previous reader has retired
        }
        if (warp == 4) FA_TRACE(24, cnt);
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(prow + ((c ^ sw) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        fence_proxy_async_smem();
        tc_fence_before();
        warp_arrive(&bar_p_full[t], lane);
        if (warp == 4) FA_TRACE(25, cnt);
        ++cnt;
      }
      // ---- epilogue: O / l -> bf16 rows (through this warp's slice of sP and a TMA store), log-sum-exp
      if (warp == 4) FA_TRACE(26, cnt - 1u);
      mbar_wait(&bar_pv_done[t], (cnt - 1u) & 1u);
      tc_fence_after();
      if (warp == 4) FA_TRACE(27, cnt - 1u);
      const float inv = l > 0.0f ? 1.0f / l : 0.0f;
      const int row0 = im.tile[t] * 128 + quarter * 32;
      store_acc_tma(&p.tmO, o_tmem, pslice, lane, 0, 1, p.katoms, p.o_chunks, inv, row0, h, b, row0 < p.S);
      if (row_ok && p.lse != nullptr)
        p.lse[((size_t)b * p.H + h) * p.S + q] = l > 0.0f ? m_ref + log2f(l) : INFINITY;
      if (warp == 4) FA_TRACE(28, cnt - 1u);
      tc_fence_before();
    }
    if (lane == 0) tma_store_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// -------------------------------------------------------------------------------------------- backward: dQ
// smem: Q tile | dO tile | dS [128][64] | ring of (K block | V block).  TMEM: S @0/64, dP @128/192, dQ @256.
// warp roles: 0 TMA | 1 issuer A (S, dP of the next block) | 2 TMEM allocator | 3 issuer B (dQ += dS K) | 4-11 softmax
__global__ void __launch_bounds__(384, 1) flash_dq_kernel(const __grid_constant__ FaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar_qdo_full, bar_qdo_empty, bar_kv_full[kFaMaxStages], bar_kv_empty[kFaMaxStages];
  __shared__ uint64_t bar_sdp_full[2], bar_sdp_free[2], bar_ds_full, bar_ds_empty, bar_dq_full;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* sQ = smem;
  uint8_t* sDO = sQ + p.katoms * kAtom128;
  uint8_t* sDS = sDO + p.katoms * kAtom128;
  uint8_t* sRing = sDS + kAtom128;
  const int stage_bytes = 2 * p.katoms * kAtom64;
  const int stages = p.stages;

  if (warp == 1 && lane == 0) {
    mbar_init(&bar_qdo_full, 1);
    mbar_init(&bar_qdo_empty, 1);
    for (int i = 0; i < kFaMaxStages; ++i) {
      mbar_init(&bar_kv_full[i], 1);
      mbar_init(&bar_kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_sdp_full[i], 1);
      mbar_init(&bar_sdp_free[i], 8);     // one arrival per softmax warp
    }
    mbar_init(&bar_ds_full, 8);
    mbar_init(&bar_ds_empty, 1);
    mbar_init(&bar_dq_full, 1);
    mbar_fence_init();
    fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
    tma_prefetch_desc(&p.tmDO);
  }
  if (warp == 2) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int BH = p.B * p.H;

  if (warp == 0) {
    {   // TMA producer: whole warp, one elected lane issues
      int stage = 0, qi = 0;
      uint32_t phase = 0;
      for (int it = blockIdx.x; it < p.n_items; it += gridDim.x, ++qi) {
        const int pr = it / BH, bh = it - pr * BH;
        const int b = bh / p.H, h = bh - b * p.H, kvh = h / p.G, tile = p.m_tiles - 1 - pr;
        const int n = visible_blocks(p, tile);
        mbar_wait_u(&bar_qdo_empty, (uint32_t)(qi & 1) ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(&bar_qdo_full, (uint32_t)(2 * p.katoms * kAtom128));
          for (int a = 0; a < p.katoms; ++a) {
            tma_load_4d(sQ + a * kAtom128, &p.tmQ, &bar_qdo_full, a * 64, tile * 128, h, b);
            tma_load_4d(sDO + a * kAtom128, &p.tmDO, &bar_qdo_full, a * 64, tile * 128, h, b);
          }
        }
        __syncwarp();
        for (int j = 0; j < n; ++j) {
          mbar_wait_u(&bar_kv_empty[stage], phase ^ 1u);
          if (elect_one()) {
            mbar_expect_tx(&bar_kv_full[stage], (uint32_t)stage_bytes);
            uint8_t* sK = sRing + stage * stage_bytes;
            uint8_t* sV = sK + p.katoms * kAtom64;
            for (int a = 0; a < p.katoms; ++a) {
              tma_load_4d(sK + a * kAtom64, &p.tmK, &bar_kv_full[stage], a * 64, j * 64, kvh, b);
              tma_load_4d(sV + a * kAtom64, &p.tmV, &bar_kv_full[stage], a * 64, j * 64, kvh, b);
            }
          }
          __syncwarp();
          if (++stage == stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // issuer A: S = Q K^T and dP = dO V^T of block c into buffer c & 1, as far ahead as the two buffers allow
    // (whole warp in the loop, one elected lane issues: see ptx.cuh elect_one)
    {
      const uint32_t idesc_s = umma_idesc(1u, 0, 0, 128, 64);
      const uint32_t q_lo = lo_kmajor(smem_u32(sQ)), do_lo = lo_kmajor(smem_u32(sDO));
      const uint32_t ring_addr = smem_u32(sRing);
      uint32_t cnt = 0;
      int stage = 0, qi = 0;
      uint32_t phase = 0;
      for (int it = blockIdx.x; it < p.n_items; it += gridDim.x, ++qi) {
        const int n = visible_blocks(p, p.m_tiles - 1 - it / BH);
        mbar_wait_u(&bar_qdo_full, (uint32_t)(qi & 1));
        for (int j = 0; j < n; ++j) {
          const uint32_t buf = cnt & 1u;
          FA_TRACE(0, cnt);
          if (cnt >= 2) mbar_wait_u(&bar_sdp_free[buf], ((cnt >> 1) - 1u) & 1u);   // the buffer's previous block was read
          FA_TRACE(1, cnt);
          mbar_wait_u(&bar_kv_full[stage], phase);
          tc_fence_after();
          FA_TRACE(2, cnt);
          if (elect_one()) {
            const uint32_t k_lo = lo_kmajor(ring_addr + stage * stage_bytes);
            mma_over_hd(tmem + buf * 64, q_lo, kAtom128 >> 4, k_lo, kAtom64 >> 4, p.ksteps, idesc_s);
            mma_over_hd(tmem + 128 + buf * 64, do_lo, kAtom128 >> 4, k_lo + ((p.katoms * kAtom64) >> 4), kAtom64 >> 4,
                        p.ksteps, idesc_s);
            umma_commit(&bar_sdp_full[buf]);
            if (j + 1 == n) umma_commit(&bar_qdo_empty);
          }
          __syncwarp();
          FA_TRACE(3, cnt);
          if (++stage == stages) {
            stage = 0;
            phase ^= 1u;
          }
          ++cnt;
        }
      }
    }
  } else if (warp == 3) {
    // issuer B: dQ += dS K as soon as the softmax warps have written dS
    {
      const uint32_t idesc_dq = umma_idesc(1u, 0, 1, 128, (uint32_t)p.n_hd);
      const uint32_t ds_lo = lo_kmajor(smem_u32(sDS));
      const uint32_t ring_addr = smem_u32(sRing);
      uint32_t cnt = 0;
      int stage = 0;
      for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
        const int n = visible_blocks(p, p.m_tiles - 1 - it / BH);
        for (int j = 0; j < n; ++j) {
          mbar_wait_u(&bar_ds_full, cnt & 1u);
          tc_fence_after();
          FA_TRACE(4, cnt);
          if (elect_one()) {
            mma_over_rows64(tmem + 256, ds_lo, lo_mn64(ring_addr + stage * stage_bytes), idesc_dq, j > 0);
            umma_commit(&bar_ds_empty);
            umma_commit(&bar_kv_empty[stage]);
            if (j + 1 == n) umma_commit(&bar_dq_full);
          }
          __syncwarp();
          FA_TRACE(5, cnt);
          if (++stage == stages) stage = 0;
          ++cnt;
        }
      }
    }
  } else if (warp >= 4) {
    // 8 warps: lane quarter = warp & 3, column half = (warp - 4) / 4 of each 64-key block; no row reductions in backward.
    // dS is written WITHOUT the softmax scale (it is applied once to the dQ accumulator in the epilogue).
    const int quarter = warp & 3, half = (warp - 4) >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
    uint8_t* dsrow = sDS + r * 128;
    uint8_t* stage_out = sRing + stages * stage_bytes + (warp - 4) * 4096;   // TMA staging of this warp's dQ rows
    const int sw = r & 7;
    uint32_t cnt = 0;
    int qi = 0;
    for (int it = blockIdx.x; it < p.n_items; it += gridDim.x, ++qi) {
      const int pr = it / BH, bh = it - pr * BH;
      const int b = bh / p.H, h = bh - b * p.H, tile = p.m_tiles - 1 - pr;
      const int n = visible_blocks(p, tile);
      const int q = tile * 128 + r;
      const bool row_ok = q < p.S;
      const size_t stat = ((size_t)b * p.H + h) * p.S + q;
      const float L = row_ok ? ldg_now(p.lse + stat) : INFINITY;
      const float dl = row_ok ? ldg_now(p.delta + stat) : 0.0f;
      const int bq = (p.bid_q != nullptr && row_ok) ? p.bid_q[(size_t)b * p.S + q] : 0;
      int4 kin = load_kblk(p, b, 0);
      for (int j = 0; j < n; ++j) {
        const uint32_t buf = cnt & 1u;
        const int4 kic = kin;
        if (j + 1 < n) kin = load_kblk(p, b, j + 1);
        uint32_t b0, b1;
        row_mask(p, kic, b, q, bq, j * 64, b0, b1);
        const uint32_t bits = half ? b1 : b0;
        const bool full = __all_sync(0xffffffffu, bits == 0xffffffffu);
        if (warp == 4) FA_TRACE(6, cnt);
        mbar_wait(&bar_sdp_full[buf], (cnt >> 1) & 1u);
        tc_fence_after();
        if (warp == 4) FA_TRACE(7, cnt);
        uint32_t s[32], dp[32];
        tmem_ld_32x32(lane_base + buf * 64 + half * 32, s);
        tmem_ld_32x32(lane_base + 128 + buf * 64 + half * 32, dp);
        tmem_ld_wait();
        tc_fence_before();
        warp_arrive(&bar_sdp_free[buf], lane);
        if (warp == 4) FA_TRACE(8, cnt);
        // every element is computed; masked ones are then cleared in the packed words (an exponential of a masked score
        // may be inf, its product NaN: the AND removes the bit pattern, nothing is multiplied by it)
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float p0 = ex2f(fmaf(__uint_as_float(s[2 * i]), p.sl2, -L));
          const float p1 = ex2f(fmaf(__uint_as_float(s[2 * i + 1]), p.sl2, -L));
          pk[i] = pack2(p0 * (__uint_as_float(dp[2 * i]) - dl), p1 * (__uint_as_float(dp[2 * i + 1]) - dl));
        }
        if (!full) {
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] &= pair_mask(bits, 2 * i);
        }
        if (warp == 4) FA_TRACE(9, cnt);
        if (cnt > 0) mbar_wait(&bar_ds_empty, (cnt - 1u) & 1u);
        if (warp == 4) FA_TRACE(10, cnt);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(dsrow + (((half * 4 + c) ^ sw) << 4)) =
              make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        fence_proxy_async_smem();
        warp_arrive(&bar_ds_full, lane);
        if (warp == 4) FA_TRACE(11, cnt);
        ++cnt;
      }
      if (warp == 4) FA_TRACE(12, cnt - 1u);
      mbar_wait(&bar_dq_full, (uint32_t)(qi & 1));
      tc_fence_after();
      if (warp == 4) FA_TRACE(13, cnt - 1u);
      if (p.tma_epi) {
        const int row0 = tile * 128 + quarter * 32;
        store_acc_tma(&p.tmO, lane_base + 256, stage_out, lane, half, 2, p.katoms, p.o_chunks, p.scale, row0, h, b,
                      row0 < p.S);
      } else {
        bf16* drow = p.dq + (long long)b * p.g_sb + (long long)q * p.g_ld + (long long)h * p.gq_sh;
        store_acc_rows(lane_base + 256, drow, row_ok, half, 2, p.o_chunks, p.hd, p.scale);
      }
      if (warp == 4) FA_TRACE(14, cnt - 1u);
      tc_fence_before();
      // all eight warps have left the dQ accumulator before issuer B may overwrite it: their next arrival on ds_full
      // (block 0 of the next item) comes after this point in program order
    }
    if (lane == 0) tma_store_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// --------------------------------------------------------------------------------------- backward: dK and dV
// Transposed problem: TMEM lanes = the 128 keys of the tile, columns = 64 query rows of a block.
// smem: K tile | V tile | P^T [128][64] | dS^T [128][64] | ring of (Q block | dO block).
// TMEM: S^T @0/64, dP^T @128/192, accumulators @256 (mode 0: dV @256, dK @384; mode 1: dV only; mode 2: dK only).
// warp roles: 0 TMA | 1 issuer A (S^T, dP^T) | 2 TMEM allocator | 3 issuer B (dV += P^T dO, dK += dS^T Q) | 4-11 softmax
struct DkvItem {
  int b, kvh, jt, mode, ib0, nsteps;
};
__device__ __forceinline__ DkvItem dkv_item(const FaParams& p, int it) {
  DkvItem r;
  const int per = p.B * p.KVH * p.n_modes;
  r.jt = it / per;                       // key tile 0 sees the most query blocks under a causal mask: first
  int x = it - r.jt * per;
  r.mode = p.n_modes == 2 ? 1 + (x % 2) : 0;
  x /= p.n_modes;
  r.b = x / p.KVH;
  r.kvh = x - r.b * p.KVH;
  r.ib0 = p.causal ? 2 * r.jt : 0;
  r.nsteps = p.G * (p.n_qblk - r.ib0);
  return r;
}

__global__ void __launch_bounds__(384, 1) flash_dkv_kernel(const __grid_constant__ FaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar_kv_full, bar_kv_empty, bar_ring_full[kFaMaxStages], bar_ring_empty[kFaMaxStages];
  __shared__ uint64_t bar_sdp_full[2], bar_sdp_free[2], bar_pds_full, bar_pds_empty, bar_acc_full;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float stats[8][3][32];   // per softmax warp: lse, delta, block id of its 32 query columns

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* sK = smem;
  uint8_t* sV = sK + p.katoms * kAtom128;
  uint8_t* sPT = sV + p.katoms * kAtom128;
  uint8_t* sDST = p.n_modes == 2 ? sPT : sPT + kAtom128;   // head_dim > 128: an item needs only one of the two
  uint8_t* sRing = sDST + kAtom128;
  const int stage_bytes = 2 * p.katoms * kAtom64;
  const int stages = p.stages;

  if (warp == 1 && lane == 0) {
    mbar_init(&bar_kv_full, 1);
    mbar_init(&bar_kv_empty, 1);
    for (int i = 0; i < kFaMaxStages; ++i) {
      mbar_init(&bar_ring_full[i], 1);
      mbar_init(&bar_ring_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_sdp_full[i], 1);
      mbar_init(&bar_sdp_free[i], 8);
    }
    mbar_init(&bar_pds_full, 8);
    mbar_init(&bar_pds_empty, 1);
    mbar_init(&bar_acc_full, 1);
    mbar_fence_init();
    fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
    tma_prefetch_desc(&p.tmDO);
  }
  if (warp == 2) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    {   // TMA producer: whole warp, one elected lane issues
      int stage = 0, qi = 0;
      uint32_t phase = 0;
      for (int it = blockIdx.x; it < p.n_items; it += gridDim.x, ++qi) {
        const DkvItem im = dkv_item(p, it);
        mbar_wait_u(&bar_kv_empty, (uint32_t)(qi & 1) ^ 1u);
        const bool need_v = im.mode != 1;
        if (elect_one()) {
          mbar_expect_tx(&bar_kv_full, (uint32_t)((need_v ? 2 : 1) * p.katoms * kAtom128));
          for (int a = 0; a < p.katoms; ++a) {
            tma_load_4d(sK + a * kAtom128, &p.tmK, &bar_kv_full, a * 64, im.jt * 128, im.kvh, im.b);
            if (need_v) tma_load_4d(sV + a * kAtom128, &p.tmV, &bar_kv_full, a * 64, im.jt * 128, im.kvh, im.b);
          }
        }
        __syncwarp();
        for (int hh = 0; hh < p.G; ++hh) {
          const int h = im.kvh * p.G + hh;
          for (int ib = im.ib0; ib < p.n_qblk; ++ib) {
            mbar_wait_u(&bar_ring_empty[stage], phase ^ 1u);
            if (elect_one()) {
              mbar_expect_tx(&bar_ring_full[stage], (uint32_t)stage_bytes);
              uint8_t* sQb = sRing + stage * stage_bytes;
              uint8_t* sDOb = sQb + p.katoms * kAtom64;
              for (int a = 0; a < p.katoms; ++a) {
                tma_load_4d(sQb + a * kAtom64, &p.tmQ, &bar_ring_full[stage], a * 64, ib * 64, h, im.b);
                tma_load_4d(sDOb + a * kAtom64, &p.tmDO, &bar_ring_full[stage], a * 64, ib * 64, h, im.b);
              }
            }
            __syncwarp();
            if (++stage == stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // issuer A: S^T = K Q^T and dP^T = V dO^T of step c into buffer c & 1 (whole warp in the loop, elected lane issues)
    {
      const uint32_t idesc_st = umma_idesc(1u, 0, 0, 128, 64);
      const uint32_t k_lo = lo_kmajor(smem_u32(sK)), v_lo = lo_kmajor(smem_u32(sV));
      const uint32_t ring_addr = smem_u32(sRing);
      uint32_t cnt = 0;
      int stage = 0, qi = 0;
      uint32_t phase = 0;
      for (int it = blockIdx.x; it < p.n_items; it += gridDim.x, ++qi) {
        const DkvItem im = dkv_item(p, it);
        const int n = im.nsteps;
        const bool do_dk = im.mode != 1;
        mbar_wait_u(&bar_kv_full, (uint32_t)(qi & 1));
        for (int sidx = 0; sidx < n; ++sidx) {
          const uint32_t buf = cnt & 1u;
          FA_TRACE(32, cnt);
          if (cnt >= 2) mbar_wait_u(&bar_sdp_free[buf], ((cnt >> 1) - 1u) & 1u);
          FA_TRACE(33, cnt);
          mbar_wait_u(&bar_ring_full[stage], phase);
          tc_fence_after();
          FA_TRACE(34, cnt);
          if (elect_one()) {
            const uint32_t qb_lo = lo_kmajor(ring_addr + stage * stage_bytes);
            mma_over_hd(tmem + buf * 64, k_lo, kAtom128 >> 4, qb_lo, kAtom64 >> 4, p.ksteps, idesc_st);
            if (do_dk)
              mma_over_hd(tmem + 128 + buf * 64, v_lo, kAtom128 >> 4, qb_lo + ((p.katoms * kAtom64) >> 4), kAtom64 >> 4,
                          p.ksteps, idesc_st);
            umma_commit(&bar_sdp_full[buf]);
            if (sidx + 1 == n) umma_commit(&bar_kv_empty);
          }
          __syncwarp();
          FA_TRACE(35, cnt);
          if (++stage == stages) {
            stage = 0;
            phase ^= 1u;
          }
          ++cnt;
        }
      }
    }
  } else if (warp == 3) {
    // issuer B: the accumulating products, as soon as P^T / dS^T of a step are in smem
    {
      const uint32_t idesc_acc = umma_idesc(1u, 0, 1, 128, (uint32_t)p.n_hd);
      const uint32_t pt_lo = lo_kmajor(smem_u32(sPT)), dst_lo = lo_kmajor(smem_u32(sDST));
      const uint32_t ring_addr = smem_u32(sRing);
      uint32_t cnt = 0;
      int stage = 0;
      for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
        const DkvItem im = dkv_item(p, it);
        const bool do_dv = im.mode != 2, do_dk = im.mode != 1;
        const uint32_t dv_tmem = tmem + 256, dk_tmem = tmem + (im.mode == 0 ? 384u : 256u);
        for (int sidx = 0; sidx < im.nsteps; ++sidx) {
          mbar_wait_u(&bar_pds_full, cnt & 1u);
          tc_fence_after();
          FA_TRACE(36, cnt);
          if (elect_one()) {
            const uint32_t qb_addr = ring_addr + stage * stage_bytes;
            if (do_dv) mma_over_rows64(dv_tmem, pt_lo, lo_mn64(qb_addr + p.katoms * kAtom64), idesc_acc, sidx > 0);
            if (do_dk) mma_over_rows64(dk_tmem, dst_lo, lo_mn64(qb_addr), idesc_acc, sidx > 0);
            umma_commit(&bar_pds_empty);
            umma_commit(&bar_ring_empty[stage]);
            if (sidx + 1 == im.nsteps) umma_commit(&bar_acc_full);
          }
          __syncwarp();
          FA_TRACE(37, cnt);
          if (++stage == stages) stage = 0;
          ++cnt;
        }
      }
    }
  } else if (warp >= 4) {
    const int quarter = warp & 3, half = (warp - 4) >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
    float (*st)[32] = stats[warp - 4];
    uint8_t* ptrow = sPT + r * 128;
    uint8_t* dsrow = sDST + r * 128;
    const int sw = r & 7;
    uint32_t cnt = 0;
    int qi = 0;
    for (int it = blockIdx.x; it < p.n_items; it += gridDim.x, ++qi) {
      const DkvItem im = dkv_item(p, it);
      const bool do_dv = im.mode != 2, do_dk = im.mode != 1;
      const int b = im.b;
      const int key = im.jt * 128 + r;
      const bool key_in = key < p.S;
      const bool key_ok = key_in && (p.keymask == nullptr || p.keymask[(size_t)b * p.S + key] != 0);
      const int bk = (p.bid_k != nullptr && key_in) ? p.bid_k[(size_t)b * p.S + key] : 0;
      // per-column statistics of a step (lse, delta, block id of query q0 + lane), loaded one step ahead
      auto load_stats = [&](int hh, int ib, float& o_l, float& o_d, int& o_b) {
        const int qq = ib * 64 + half * 32 + lane;
        const bool qok = qq < p.S;
        const size_t stat = ((size_t)b * p.H + (im.kvh * p.G + hh)) * p.S + qq;
        o_l = qok ? ldg_now(p.lse + stat) : INFINITY;
        o_d = (qok && do_dk) ? ldg_now(p.delta + stat) : 0.0f;
        o_b = (p.bid_q != nullptr && qok) ? ldg_now(p.bid_q + (size_t)b * p.S + qq) : 0;
      };
      float n_l, n_d;
      int n_b;
      load_stats(0, im.ib0, n_l, n_d, n_b);
      for (int hh = 0; hh < p.G; ++hh) {
        for (int ib = im.ib0; ib < p.n_qblk; ++ib) {
          const uint32_t buf = cnt & 1u;
          const int q0 = ib * 64 + half * 32;
          __syncwarp();
          st[0][lane] = n_l;
          st[1][lane] = n_d;
          st[2][lane] = __int_as_float(n_b);
          __syncwarp();
          {
            int nh = hh, nib = ib + 1;
            if (nib == p.n_qblk) {
              nib = im.ib0;
              ++nh;
            }
            if (nh < p.G) load_stats(nh, nib, n_l, n_d, n_b);
          }
          uint32_t bits = key_ok ? 0xffffffffu : 0u;
          if (p.causal) bits &= ~low_mask(key - q0);      // column c is visible iff q0 + c >= key
          if (p.bid_k != nullptr) {
            // query-side summary of this half block: every query at or after the key's block id, or none, or mixed
            const int4 qe = ldg_now(p.mask_ws + ((size_t)b * p.n_qblk + ib) * 2 + 1);
            const int qmin = half ? qe.z : qe.x, qmax = half ? qe.w : qe.y;
            if (bk > qmin) {
              if (bk > qmax) {
                bits = 0u;
              } else {
                uint32_t w2 = 0;
#pragma unroll
                for (int c = 0; c < 32; ++c)
                  if (bk <= __float_as_int(st[2][c])) w2 |= 1u << c;
                bits &= w2;
              }
            }
          }
          const bool full = __all_sync(0xffffffffu, bits == 0xffffffffu);
          if (warp == 4) FA_TRACE(38, cnt);
          mbar_wait(&bar_sdp_full[buf], (cnt >> 1) & 1u);
          tc_fence_after();
          if (warp == 4) FA_TRACE(39, cnt);
          uint32_t s[32], dp[32];
          tmem_ld_32x32(lane_base + buf * 64 + half * 32, s);
          if (do_dk) tmem_ld_32x32(lane_base + 128 + buf * 64 + half * 32, dp);
          tmem_ld_wait();
          tc_fence_before();
          warp_arrive(&bar_sdp_free[buf], lane);
          if (warp == 4) FA_TRACE(40, cnt);
          // P^T = exp2(S^T * sl2 - lse[q]); dS^T = P^T * (dP^T - delta[q]) (the softmax scale goes into the dK epilogue)
          uint32_t pp[16], pd[16];
          if (do_dk) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 Lv = *reinterpret_cast<const float4*>(&st[0][4 * i]);
              const float4 Dv = *reinterpret_cast<const float4*>(&st[1][4 * i]);
              const float p0 = ex2f(fmaf(__uint_as_float(s[4 * i]), p.sl2, -Lv.x));
              const float p1 = ex2f(fmaf(__uint_as_float(s[4 * i + 1]), p.sl2, -Lv.y));
              const float p2 = ex2f(fmaf(__uint_as_float(s[4 * i + 2]), p.sl2, -Lv.z));
              const float p3 = ex2f(fmaf(__uint_as_float(s[4 * i + 3]), p.sl2, -Lv.w));
              pp[2 * i] = pack2(p0, p1);
              pp[2 * i + 1] = pack2(p2, p3);
              pd[2 * i] = pack2(p0 * (__uint_as_float(dp[4 * i]) - Dv.x), p1 * (__uint_as_float(dp[4 * i + 1]) - Dv.y));
              pd[2 * i + 1] = pack2(p2 * (__uint_as_float(dp[4 * i + 2]) - Dv.z), p3 * (__uint_as_float(dp[4 * i + 3]) - Dv.w));
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 Lv = *reinterpret_cast<const float4*>(&st[0][4 * i]);
              pp[2 * i] = pack2(ex2f(fmaf(__uint_as_float(s[4 * i]), p.sl2, -Lv.x)),
                                ex2f(fmaf(__uint_as_float(s[4 * i + 1]), p.sl2, -Lv.y)));
              pp[2 * i + 1] = pack2(ex2f(fmaf(__uint_as_float(s[4 * i + 2]), p.sl2, -Lv.z)),
                                    ex2f(fmaf(__uint_as_float(s[4 * i + 3]), p.sl2, -Lv.w)));
              pd[2 * i] = pd[2 * i + 1] = 0u;
            }
          }
          if (!full) {      // masked elements: clear the packed halves (see flash_dq_kernel)
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const uint32_t m = pair_mask(bits, 2 * i);
              pp[i] &= m;
              pd[i] &= m;
            }
          }
          if (warp == 4) FA_TRACE(41, cnt);
          if (cnt > 0) mbar_wait(&bar_pds_empty, (cnt - 1u) & 1u);
          if (warp == 4) FA_TRACE(42, cnt);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int off = ((half * 4 + c) ^ sw) << 4;
            if (do_dv) *reinterpret_cast<uint4*>(ptrow + off) = make_uint4(pp[4 * c], pp[4 * c + 1], pp[4 * c + 2], pp[4 * c + 3]);
            if (do_dk) *reinterpret_cast<uint4*>(dsrow + off) = make_uint4(pd[4 * c], pd[4 * c + 1], pd[4 * c + 2], pd[4 * c + 3]);
          }
          fence_proxy_async_smem();
          warp_arrive(&bar_pds_full, lane);
          if (warp == 4) FA_TRACE(43, cnt);
          ++cnt;
        }
      }
      // ---- epilogue: the accumulated dV / dK rows of this key tile
      mbar_wait(&bar_acc_full, (uint32_t)(qi & 1));
      tc_fence_after();
      const long long roff = (long long)b * p.g_sb + (long long)key * p.g_ld + (long long)im.kvh * p.gkv_sh;
      if (do_dv) store_acc_rows(lane_base + 256, p.dv + roff, key_in, half, 2, p.o_chunks, p.hd, 1.0f);
      if (do_dk)
        store_acc_rows(lane_base + (im.mode == 0 ? 384u : 256u), p.dk + roff, key_in, half, 2, p.o_chunks, p.hd, p.scale);
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// delta[b, h, q] = sum_d dO[b, q, h, d] * O[b, q, h, d].  HBM-bound (reads O and dO once): one thread per 16-byte
// chunk, kLanes = head_dim chunks (rounded up to a power of two) lanes per row, shuffle reduction inside the group.
template <int kLanes>
__global__ void __launch_bounds__(256) attn_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ dout,
                                                         float* __restrict__ delta, int B, int H, int S, int hd,
                                                         long long ld, long long sh, long long sb) {
  const long long rows = (long long)B * S * H;
  const int sub = threadIdx.x % kLanes;
  const long long row0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / kLanes;
  const long long stride = (long long)gridDim.x * blockDim.x / kLanes;
  for (long long row = row0;; row += stride) {   // whole warps stay converged for the shuffles
    const bool live = row < rows;
    float acc = 0.0f;
    int b = 0, h = 0, q = 0;
    if (live) {
      h = (int)(row % H);
      const long long bq = row / H;
      q = (int)(bq % S);
      b = (int)(bq / S);
      const long long off = (long long)b * sb + (long long)q * ld + (long long)h * sh;
      for (int c = sub * 8; c < hd; c += kLanes * 8) {
        const uint4 x = __ldg(reinterpret_cast<const uint4*>(o + off + c));
        const uint4 y = __ldg(reinterpret_cast<const uint4*>(dout + off + c));
        const __nv_bfloat162* xh = reinterpret_cast<const __nv_bfloat162*>(&x);
        const __nv_bfloat162* yh = reinterpret_cast<const __nv_bfloat162*>(&y);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 u = __bfloat1622float2(xh[i]), v = __bfloat1622float2(yh[i]);
          acc += u.x * v.x + u.y * v.y;
        }
      }
    }
#pragma unroll
    for (int o2 = kLanes / 2; o2 > 0; o2 >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o2);
    if (live && sub == 0) delta[((size_t)b * H + h) * S + q] = acc;
    if (!__any_sync(0xffffffffu, row + stride < rows)) break;
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFnFa)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// (hd, S, heads, B) view with element strides (1, ld, s_head, s_batch); box = (64, box_rows): one 64-column atom
static int encode_rows(CUtensorMap* m, const void* ptr, uint64_t hd, uint64_t S, uint64_t heads, uint64_t B, int64_t ld,
                       int64_t s_head, int64_t s_batch, uint32_t box_rows, const char* what) {
  static EncodeTiledFnFa fn = nullptr;
  if (fn == nullptr) {
    void* pfn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &pfn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFnFa>(pfn);
  }
  B200_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[4] = {hd, S, heads, B};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)s_head * 2, (cuuint64_t)s_batch * 2};
  cuuint32_t box[4] = {64, box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < 3; ++i)
    B200_CHECK(strides[i] % 16 == 0 && strides[i] > 0, "%s: stride %d not a positive multiple of 16 bytes", what, i + 1);
  B200_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "%s: base pointer not 16-byte aligned", what);
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_CHECK(r == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled failed (%d)", what, (int)r);
  return 0;
}

constexpr int kSmemBudget = 227 * 1024 - 4096;   // dynamic bytes available next to the static barriers / stats

static long long* g_fa_trace = nullptr;

static int fill_common(FaParams& kp, int64_t B, int64_t H, int64_t KVH, int64_t S, int64_t hd, float scale, int causal,
                       const uint8_t* keymask, const int32_t* bid_q, const int32_t* bid_k) {
  B200_CHECK(hd % 8 == 0 && hd >= 16 && hd <= 256, "flash_attn: head_dim=%lld unsupported (multiple of 8, <= 256)",
             (long long)hd);
  B200_CHECK(B > 0 && S > 0 && H > 0 && KVH > 0 && H % KVH == 0, "flash_attn: bad geometry");
  B200_CHECK((bid_q == nullptr) == (bid_k == nullptr), "flash_attn: bid_q and bid_k go together");
  B200_CHECK(B * H * ceil_div(S, 64) < (1ll << 30), "flash_attn: problem too large");
  memset(&kp, 0, sizeof(kp));
  kp.B = (int)B;
  kp.H = (int)H;
  kp.KVH = (int)KVH;
  kp.G = (int)(H / KVH);
  kp.S = (int)S;
  kp.hd = (int)hd;
  kp.katoms = (int)ceil_div(hd, 64);
  kp.ksteps = (int)ceil_div(hd, 16);
  kp.n_hd = (int)ceil_div(hd, 16) * 16;
  kp.o_chunks = (int)ceil_div(hd, 32);
  kp.m_tiles = (int)ceil_div(S, 128);
  kp.n_qblk = (int)ceil_div(S, 64);
  kp.causal = causal;
  kp.scale = scale;
  kp.sl2 = scale * 1.4426950408889634f;
  kp.keymask = keymask;
  kp.bid_q = bid_q;
  kp.bid_k = bid_k;
  kp.n_modes = 1;
  kp.trace = g_fa_trace;
  return 0;
}

// mask summary table (see attn_mask_prep_kernel): needed when a key mask or the block-id rule is present
static int prep_masks(FaParams& kp, int32_t* mask_ws, cudaStream_t stream) {
  kp.mask_ws = nullptr;
  if (kp.keymask == nullptr && kp.bid_k == nullptr) return 0;
  B200_CHECK(mask_ws != nullptr, "flash_attn: mask_ws (int32 [B, ceil(S/64), 8]) is required with a key mask / block ids");
  B200_CHECK((reinterpret_cast<uintptr_t>(mask_ws) & 15) == 0, "flash_attn: mask_ws must be 16-byte aligned");
  const int warps = kp.B * kp.n_qblk;
  attn_mask_prep_kernel<<<(unsigned)ceil_div((int64_t)warps * 32, 128), 128, 0, stream>>>(
      kp.keymask, kp.bid_q, kp.bid_k, kp.B, kp.S, kp.n_qblk, reinterpret_cast<int4*>(mask_ws));
  B200_LAUNCH_OK();
  kp.mask_ws = reinterpret_cast<const int4*>(mask_ws);
  return 0;
}

template <typename K>
static int set_smem(K kernel, int bytes) {
  B200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return 0;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_flash_attn_set_trace(void* buf) {
  b200::g_fa_trace = reinterpret_cast<long long*>(buf);
  return 0;
}

extern "C" int b200_flash_attn_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int64_t B,
                                   int64_t H, int64_t KVH, int64_t S, int64_t head_dim, int64_t qkv_ld,
                                   int64_t qkv_s_head, int64_t qkv_s_batch, int64_t o_ld, int64_t o_s_head,
                                   int64_t o_s_batch, float scale, int causal, const uint8_t* keymask,
                                   const int32_t* bid_q, const int32_t* bid_k, int32_t* mask_ws, void* stream) {
  FaParams kp;
  if (fill_common(kp, B, H, KVH, S, head_dim, scale, causal, keymask, bid_q, bid_k)) return 1;
  if (prep_masks(kp, mask_ws, reinterpret_cast<cudaStream_t>(stream))) return 1;
  if (encode_rows(&kp.tmO, out, head_dim, S, H, B, o_ld, o_s_head, o_s_batch, 32, "flash O")) return 1;
  B200_CHECK(o_ld % 8 == 0 && o_s_head % 8 == 0 && o_s_batch % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
             "flash_attn_fwd: output rows must be 16-byte aligned");
  if (encode_rows(&kp.tmQ, q, head_dim, S, H, B, qkv_ld, qkv_s_head, qkv_s_batch, 128, "flash Q")) return 1;
  if (encode_rows(&kp.tmK, k, head_dim, S, KVH, B, qkv_ld, qkv_s_head, qkv_s_batch, 64, "flash K")) return 1;
  if (encode_rows(&kp.tmV, v, head_dim, S, KVH, B, qkv_ld, qkv_s_head, qkv_s_batch, 64, "flash V")) return 1;
  kp.out = (bf16*)out;
  kp.lse = lse;
  kp.o_ld = o_ld;
  kp.o_sh = o_s_head;
  kp.o_sb = o_s_batch;
  const int tiles = head_dim <= 128 ? 2 : 1;
  const int fixed = tiles * kp.katoms * kAtom128 + tiles * kAtom128;
  const int stage_bytes = 2 * kp.katoms * kAtom64;
  int stages = (kSmemBudget - 1024 - fixed) / stage_bytes;
  if (stages > kFaMaxStages) stages = kFaMaxStages;
  B200_CHECK(stages >= 2, "flash_attn_fwd: shared memory budget exceeded");
  kp.stages = stages;
  const int smem = 1024 + fixed + stages * stage_bytes;
  if (tiles == 2) {
    kp.n_items = (int)(B * KVH * ceil_div((int64_t)kp.G * kp.m_tiles, 2));
    if (set_smem(flash_fwd_kernel<2>, smem)) return 1;
    const unsigned grid = (unsigned)(kp.n_items < num_sms() ? kp.n_items : num_sms());
    flash_fwd_kernel<2><<<grid, 384, smem, reinterpret_cast<cudaStream_t>(stream)>>>(kp);
  } else {
    kp.n_items = (int)(B * H * kp.m_tiles);
    if (set_smem(flash_fwd_kernel<1>, smem)) return 1;
    const unsigned grid = (unsigned)(kp.n_items < num_sms() ? kp.n_items : num_sms());
    flash_fwd_kernel<1><<<grid, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(kp);
  }
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int b200_flash_attn_bwd(const void* q, const void* k, const void* v, const void* out, const void* dout,
                                   const float* lse, float* delta, void* dq, void* dk, void* dv, int64_t B, int64_t H,
                                   int64_t KVH, int64_t S, int64_t head_dim, int64_t qkv_ld, int64_t qkv_s_head,
                                   int64_t qkv_s_batch, int64_t o_ld, int64_t o_s_head, int64_t o_s_batch,
                                   int64_t g_ld, int64_t g_s_head, int64_t g_s_batch, float scale, int causal,
                                   const uint8_t* keymask, const int32_t* bid_q, const int32_t* bid_k, int32_t* mask_ws,
                                   void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FaParams kp;
  if (fill_common(kp, B, H, KVH, S, head_dim, scale, causal, keymask, bid_q, bid_k)) return 1;
  if (prep_masks(kp, mask_ws, stream)) return 1;
  B200_CHECK(lse != nullptr && delta != nullptr, "flash_attn_bwd: lse and the delta workspace are required");
  B200_CHECK(o_ld % 8 == 0 && o_s_head % 8 == 0 && o_s_batch % 8 == 0 && g_ld % 8 == 0 && g_s_head % 8 == 0 &&
                 g_s_batch % 8 == 0,
             "flash_attn_bwd: rows must be 16-byte aligned");
  B200_CHECK(((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(dq) |
               reinterpret_cast<uintptr_t>(dk) | reinterpret_cast<uintptr_t>(dv)) & 15) == 0,
             "flash_attn_bwd: pointers must be 16-byte aligned");
  // 1. delta = rowsum(dO * O)
  {
    const long long rows = (long long)B * S * H;
    const int chunks = (int)(head_dim / 8);
    const int lanes = chunks <= 4 ? 4 : (chunks <= 8 ? 8 : (chunks <= 16 ? 16 : 32));
    long long blocks = ceil_div(rows * lanes, 256);
    if (blocks > (long long)num_sms() * 8) blocks = (long long)num_sms() * 8;
#define B200_DELTA(L)                                                                                              \
  attn_delta_kernel<L><<<(unsigned)blocks, 256, 0, stream>>>((const bf16*)out, (const bf16*)dout, delta, (int)B, (int)H, \
                                                             (int)S, (int)head_dim, o_ld, o_s_head, o_s_batch)
    if (lanes == 4) B200_DELTA(4); else if (lanes == 8) B200_DELTA(8); else if (lanes == 16) B200_DELTA(16); else B200_DELTA(32);
#undef B200_DELTA
    B200_LAUNCH_OK();
  }
  kp.lse = const_cast<float*>(lse);
  kp.delta = delta;
  kp.dq = (bf16*)dq;
  kp.dk = (bf16*)dk;
  kp.dv = (bf16*)dv;
  kp.g_ld = g_ld;
  kp.gq_sh = g_s_head;
  kp.gkv_sh = g_s_head;
  kp.g_sb = g_s_batch;
  const int stage_bytes = 2 * kp.katoms * kAtom64;
  // 2. dQ: CTA = (b, h, query tile)
  {
    FaParams a = kp;
    if (encode_rows(&a.tmQ, q, head_dim, S, H, B, qkv_ld, qkv_s_head, qkv_s_batch, 128, "flash dq Q")) return 1;
    if (encode_rows(&a.tmDO, dout, head_dim, S, H, B, o_ld, o_s_head, o_s_batch, 128, "flash dq dO")) return 1;
    if (encode_rows(&a.tmK, k, head_dim, S, KVH, B, qkv_ld, qkv_s_head, qkv_s_batch, 64, "flash dq K")) return 1;
    if (encode_rows(&a.tmV, v, head_dim, S, KVH, B, qkv_ld, qkv_s_head, qkv_s_batch, 64, "flash dq V")) return 1;
    if (encode_rows(&a.tmO, dq, head_dim, S, H, B, g_ld, g_s_head, g_s_batch, 32, "flash dq out")) return 1;
    const int fixed = 2 * a.katoms * kAtom128 + kAtom128;
    const int staging = 8 * 4096;       // one 32-row slice per softmax warp for the TMA-store epilogue
    a.tma_epi = (kSmemBudget - 1024 - fixed - staging) / stage_bytes >= 2 ? 1 : 0;
    int stages = (kSmemBudget - 1024 - fixed - (a.tma_epi ? staging : 0)) / stage_bytes;
    if (stages > kFaMaxStages) stages = kFaMaxStages;
    B200_CHECK(stages >= 1, "flash_attn_bwd(dq): shared memory budget exceeded");
    a.stages = stages;
    a.n_items = (int)(B * H * a.m_tiles);
    const int smem = 1024 + fixed + stages * stage_bytes + (a.tma_epi ? staging : 0);
    if (set_smem(flash_dq_kernel, smem)) return 1;
    const unsigned grid = (unsigned)(a.n_items < num_sms() ? a.n_items : num_sms());
    flash_dq_kernel<<<grid, 384, smem, stream>>>(a);
    B200_LAUNCH_OK();
  }
  // 3. dK, dV: CTA = (b, kv head, 128-key tile [, dV | dK])
  {
    FaParams a = kp;
    if (encode_rows(&a.tmQ, q, head_dim, S, H, B, qkv_ld, qkv_s_head, qkv_s_batch, 64, "flash dkv Q")) return 1;
    if (encode_rows(&a.tmDO, dout, head_dim, S, H, B, o_ld, o_s_head, o_s_batch, 64, "flash dkv dO")) return 1;
    if (encode_rows(&a.tmK, k, head_dim, S, KVH, B, qkv_ld, qkv_s_head, qkv_s_batch, 128, "flash dkv K")) return 1;
    if (encode_rows(&a.tmV, v, head_dim, S, KVH, B, qkv_ld, qkv_s_head, qkv_s_batch, 128, "flash dkv V")) return 1;
    a.n_modes = head_dim > 128 ? 2 : 1;
    const int fixed = 2 * a.katoms * kAtom128 + (a.n_modes == 2 ? 1 : 2) * kAtom128;
    int stages = (kSmemBudget - 1024 - fixed) / stage_bytes;
    if (stages > kFaMaxStages) stages = kFaMaxStages;
    B200_CHECK(stages >= 1, "flash_attn_bwd(dkv): shared memory budget exceeded");
    a.stages = stages;
    a.n_items = (int)(B * KVH * a.m_tiles * a.n_modes);
    const int smem = 1024 + fixed + stages * stage_bytes;
    if (set_smem(flash_dkv_kernel, smem)) return 1;
    const unsigned grid = (unsigned)(a.n_items < num_sms() ? a.n_items : num_sms());
    flash_dkv_kernel<<<grid, 384, smem, stream>>>(a);
    B200_LAUNCH_OK();
  }
  return 0;
}
