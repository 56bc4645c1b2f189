// Fused flash-style attention for sm_100a (head_dim 64, no mask, no dropout), forward and
// backward, on tcgen05 tensor cores with TMEM accumulators and TMA-fed operands.
// Replaces aten::scaled_dot_product_attention (+ backward) as called from HF
// modeling_vit.py:232-246 / integrations/sdpa_attention.py:92-102.
//
// Layouts: qkv bf16 [B,T,3,H,64] (= fused QKV projection output), o / do bf16 [B,T,H,64],
// lse fp32 [B,H,T].  Q/K/V tiles are 128 rows × 64 (one 128-byte swizzle atom wide):
//   as K-major operand  (contraction over head_dim): rows = M/N index
//   as MN-major operand (contraction over tokens)  : rows = K index, +2048 B per 16 tokens
// so the same TMA tile serves both roles.
//
// forward  : CTA = (image, head, pair of 128-query tiles) ping-ponging on one stream of 128-key K/V blocks; two softmax
//            warpgroups (thread = query row = TMEM lane) + MMA-issue warp + TMA-load warp, one CTA per SM.
// backward : CTA = (128 keys, head, image) loops over query blocks; Sᵀ and dPᵀ live in TMEM
//            (lane = key), dV/dK accumulate in TMEM across the loop, dQ partials are reduced
//            into an fp32 workspace with red.global.add.
#include <cuda.h>
#include <math.h>

#include <atomic>

#include "common.cuh"
#include "sm100_prims.cuh"
#include "tmap.cuh"
#include <stdlib.h>

namespace vitk {

// Optional per-phase timeline (vitk_debug_timeline): CTA (0,0,0) records clock64() stamps.
long long* g_timeline = nullptr;   // also stamped by the CTA-pair GEMM (gemm2.cu)
int g_timeline_seq = 0;            // GEMM launches since the buffer was set (per-launch min/max stamps)
#define VITK_FSTAMP(slot)                                            \
  do {                                                               \
    if (tl != nullptr && blockIdx.x == 0 && lane == 0) tl[6000 + (slot)] = clock64(); \
  } while (0)
#define VITK_STAMP(slot)                                             \
  do {                                                               \
    if (tl != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0) tl[(slot)] = clock64(); \
  } while (0)

constexpr int kTile = 128;   // queries per CTA (fwd) / keys per CTA (bwd) / tokens per inner block
constexpr int kDh = 64;
constexpr int kTileBytes = kTile * kDh * 2;  // 16 KB
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Row `row` of a K-major, 128-byte-swizzled [128 × 128] bf16 tile held as two [128 × 64]
// sub-tiles 16 KB apart: write 32 consecutive columns starting at col0 (multiple of 32).
__device__ __forceinline__ void store_row32_sw128(uint8_t* tile, int row, int col0, const float (&v)[32]) {
  uint8_t* base = tile + (col0 >> 6) * kTileBytes + row * 128;
  const int c0 = (col0 & 63) >> 3;  // first 16-byte chunk inside the 128-byte row
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 w;
    w.x = pack_bf16x2(v[8 * q + 0], v[8 * q + 1]);
    w.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
    w.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]);
    w.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
    *reinterpret_cast<uint4*>(base + (((c0 + q) ^ (row & 7)) << 4)) = w;
  }
}

// ============================================================================ forward
// Work unit = (image, head, PAIR of 128-query tiles) — or the single left-over tile when the tile count is odd.
// One CTA per SM runs a unit: both query tiles ping-pong on ONE stream of 128-key K/V blocks, so the tensor core
// computes tile 1's scores (and the P·V of the previous block) while tile 0's softmax runs, and vice versa:
//   warps 0-3 / 4-7 : softmax warpgroup of tile 0 / tile 1 (thread = query row = TMEM lane)
//   warp 8          : MMA issue        warp 9 : TMA loads (Q tiles once, then a 3-deep ring of K and of V blocks)
//   S_t  = Q_t·Kᵀ     SS-MMA, N = 128 keys (65 cycles per 16-wide k-step instead of 2 × 49 for two N = 64 MMAs)
//   P_t  (bf16)       written to its own TMEM columns with tcgen05.st (two per 32-bit column).  P does not alias S:
//                     the warpgroup releases S_t as soon as its last chunk sits in registers (≈ one chunk of math
//                     before P_t is complete), so S_t of the NEXT block is computed while this block's softmax finishes
//                     and the warpgroup never waits for the tensor core in steady state
//   O_t += P_t·V      TS-MMA: A = P straight from TMEM, B = V rows from smem (MN-major); O_t stays in TMEM for the
//                     whole key loop — no read-back per block, no P round trip through shared memory
// TMEM columns (all 512): S0 [0,128)  S1 [128,256)  P0 [256,320)  P1 [320,384)  O0 [384,448)  O1 [448,512).
// Epilogue: O_t/l → bf16 → the (now idle) Q_t tile in shared memory → one TMA store per tile (rows ≥ T clipped).
// Softmax is single-pass with a lazily updated reference maximum: probabilities are taken relative to m_ref; whenever
// a 32-column chunk exceeds it by more than 2^16 (always on the very first chunk, otherwise only for extreme logits)
// everything accumulated so far is rescaled by 2^(old−new) ≤ 1 — exact bookkeeping, m_ref cancels in O/l and in the
// LSE.  The last key block issues narrower MMAs (N, K rounded up to 16 valid keys) instead of computing masked
// columns.  exp2 runs on the MUFU pipe (16/clk/SM — the unit that bounds this kernel at head_dim 64) except for
// kPolyOf32 of every 32 elements, which take a Cody-Waite + degree-3 polynomial on the FMA pipe (relative error
// 1.0e-4, below bf16's 3.9e-3 rounding of P).
constexpr int kFwdWG = 128;                                // threads per softmax warpgroup
constexpr int kFwdThreads = 2 * kFwdWG + 64;               // + MMA-issue warp + load warp
constexpr int kKB = 128;                                   // keys per block
constexpr int kStages = 3;
constexpr int kFwdSmemQ = 0;                               // 2 × 16 KB
constexpr int kFwdSmemK = kFwdSmemQ + 2 * kTileBytes;      // kStages × 16 KB
constexpr int kFwdSmemV = kFwdSmemK + kStages * kTileBytes;
constexpr int kFwdSmemBar = kFwdSmemV + kStages * kTileBytes;
constexpr int kFwdSmemBytes = kFwdSmemBar + 256 + 1024;    // ≈ 129 KB
constexpr int kFwdTmemCols = 512;
constexpr float kLazyMaxLog2 = 16.0f;
#ifndef VITK_ATTN_POLY_OF32
#define VITK_ATTN_POLY_OF32 0
#endif
constexpr int kPolyOf32 = VITK_ATTN_POLY_OF32;             // elements per 32 whose exp2 runs on the FMA pipe

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// 2^x for x ≤ ~16 on the FMA/ALU pipes: n = round(x) via the 1.5·2^23 trick, f = x − n ∈ [−½, ½],
// 2^f ≈ 1 + f(c1 + f(c2 + f·c3)) (minimax, max relative error 1.01e-4), exponent patched in with an integer add.
__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  float p = fmaf(f, 0.05500865f, 0.24221093f);
  p = fmaf(p, f, 0.69328299f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <int I>
__device__ __forceinline__ float softmax_exp2(float x) {
  if constexpr (kPolyOf32 > 0 && (I % 32) * kPolyOf32 / 32 != ((I % 32) + 1) * kPolyOf32 / 32) return exp2_poly(x);
  return fast_exp2(x);
}

template <int I>
__device__ __forceinline__ void softmax_pairs(const uint32_t (&r)[32], float scale_log2, float m_ref, float& rs0, float& rs1,
                                              uint32_t (&pk)[16]) {
  if constexpr (I < 16) {
    const float p0 = softmax_exp2<2 * I>(fmaf(__uint_as_float(r[2 * I]), scale_log2, -m_ref));
    const float p1 = softmax_exp2<2 * I + 1>(fmaf(__uint_as_float(r[2 * I + 1]), scale_log2, -m_ref));
    rs0 += p0;
    rs1 += p1;
    pk[I] = pack_bf16x2(p0, p1);
    softmax_pairs<I + 1>(r, scale_log2, m_ref, rs0, rs1, pk);
  }
}

__global__ void __launch_bounds__(kFwdThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_kv,
                const __grid_constant__ CUtensorMap tma_o, float* __restrict__ lse, int T, int H, int BH, float scale_log2,
                long long* tl) {
  // tl (vitk_debug_timeline, ≥ 8192 int64): [3·cta + {0,1,2}] = globaltimer at entry / %smid / globaltimer at exit
  const unsigned lin_cta = blockIdx.x;
  if (tl != nullptr && threadIdx.x == 0 && lin_cta < 2000) {
    unsigned long long t; unsigned smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    tl[3 * lin_cta] = static_cast<long long>(t);
    tl[3 * lin_cta + 1] = smid;
  }
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + kFwdSmemQ;
  uint8_t* sK = smem + kFwdSmemK;
  uint8_t* sV = smem + kFwdSmemV;
  uint64_t* bar_q = reinterpret_cast<uint64_t*>(smem + kFwdSmemBar);   // [2] Q tile landed
  uint64_t* bar_k = bar_q + 2;                // [3] K block landed
  uint64_t* bar_v = bar_k + kStages;          // [3] V block landed
  uint64_t* bar_kfree = bar_v + kStages;      // [3] every S MMA reading this K block retired
  uint64_t* bar_vfree = bar_kfree + kStages;  // [3] every P·V MMA reading this V block retired
  uint64_t* bar_s = bar_vfree + kStages;      // [2] S_t ready
  uint64_t* bar_p = bar_s + 2;                // [2] P_t written by the 4 warps of warpgroup t
  uint64_t* bar_pv = bar_p + 2;               // [2] P·V_t of a block retired (only the rescale path waits on it)
  uint64_t* bar_o = bar_pv + 2;               // [2] last P·V_t retired
  uint64_t* bar_sfree = bar_o + 2;            // [2] S_t of the current block fully read into registers by warpgroup t
  constexpr int kNumBars = 2 + 4 * kStages + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_q + kNumBars);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // unit → (image·head, first tile, number of tiles): pair units first (they cost twice a single), singles last
  const int ntile = (T + kTile - 1) / kTile, npair = ntile >> 1;
  int bh, tile0, ntl;
  {
    const int u = blockIdx.x, n_pair_units = BH * npair;
    if (u < n_pair_units) { bh = u / npair; tile0 = 2 * (u - bh * npair); ntl = 2; }
    else { bh = u - n_pair_units; tile0 = ntile - 1; ntl = 1; }
  }
  const int b = bh / H, h = bh - b * H;
  const int nblk = (T + kKB - 1) / kKB;
  const int colq = h * kDh, colk = (H + h) * kDh, colv = (2 * H + h) * kDh;

  if (tid == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_kv);
    tma_prefetch_desc(&tma_o);
    for (int i = 0; i < kNumBars; ++i)
      mbar_init(bar_q + i, ((bar_q + i >= bar_p && bar_q + i < bar_pv) || bar_q + i >= bar_sfree) ? 4 : 1);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, kFwdTmemCols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  pdl_wait();
  pdl_launch_dependents();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // broadcast marks it warp-uniform: no per-MMA elect-broadcast-retry loop
  if (warp == 0) VITK_FSTAMP(200);                                    // prologue done (clock64, CTA 0)

  if (warp == 9) {
    // ------------------------------------------------------------------ load warp
    if (elect_one()) {
      for (int t = 0; t < ntl; ++t) {
        mbar_arrive_expect_tx(&bar_q[t], kTileBytes);
        tma_load_3d(sQ + t * kTileBytes, &tma_q, &bar_q[t], colq, (tile0 + t) * kTile, b);
      }
    }
    for (int j = 0; j < nblk; ++j) {
      const int st = j % kStages;
      const uint32_t ph = ((j / kStages) - 1) & 1;
      if (j >= kStages) mbar_wait(&bar_kfree[st], ph);
      if (elect_one()) {
        mbar_arrive_expect_tx(&bar_k[st], kTileBytes);
        tma_load_3d(sK + st * kTileBytes, &tma_kv, &bar_k[st], colk, j * kKB, b);
      }
      if (j >= kStages) mbar_wait(&bar_vfree[st], ph);
      if (elect_one()) {
        mbar_arrive_expect_tx(&bar_v[st], kTileBytes);
        tma_load_3d(sV + st * kTileBytes, &tma_kv, &bar_v[st], colv, j * kKB, b);
      }
    }
    __syncwarp();
  } else if (warp == 8) {
    // ------------------------------------------------------------------ MMA-issue warp
    // Every lane runs this loop so that descriptors and addresses stay warp-uniform (uniform registers), and the
    // issuing instructions sit under elect.sync: ptxas then emits back-to-back UTCHMMA on uniform registers.
    constexpr uint32_t kTile16 = kTileBytes >> 4;
    const uint64_t q_desc = umma_smem_desc(smem_u32(sQ), 0, 1024);
    const uint64_t k_desc = umma_smem_desc(smem_u32(sK), 0, 1024);
    const uint64_t v_desc = umma_smem_desc(smem_u32(sV), kTileBytes, 1024);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(kTile, kDh, 0, 1);
    auto blk_cols = [&](int j) { return (min(kKB, T - j * kKB) + 15) & ~15; };   // keys block j needs
    auto issue_s = [&](int t, int j) {        // S_t(j) = Q_t·K(j)ᵀ
      const uint64_t bk = k_desc + static_cast<uint64_t>((j % kStages) * kTile16);
      const uint64_t aq = q_desc + static_cast<uint64_t>(t * kTile16);
      const uint32_t idesc_s = umma_idesc_bf16(kTile, blk_cols(j), 0, 0);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kDh / 16; ++k) tc_mma_bf16(tmem_base + t * 128, aq + 2 * k, bk + 2 * k, idesc_s, k > 0);
        tc_commit(&bar_s[t]);
        if (t == ntl - 1) tc_commit(&bar_kfree[j % kStages]);
      }
    };
    mbar_wait(&bar_k[0], 0);
    for (int t = 0; t < ntl; ++t) {
      mbar_wait(&bar_q[t], 0);
      tc_fence_after_sync();
      issue_s(t, 0);
    }
    for (int j = 0; j < nblk; ++j) {
      const int st = j % kStages;
      const uint32_t ph = (j / kStages) & 1;
      const int ksteps = blk_cols(j) / 16;
      const uint64_t bv = v_desc + static_cast<uint64_t>(st * kTile16);
      mbar_wait(&bar_v[st], ph);
      if (j + 1 < nblk) mbar_wait(&bar_k[(j + 1) % kStages], ((j + 1) / kStages) & 1);
      for (int t = 0; t < ntl; ++t) {
        if (j + 1 < nblk) {                     // S_t(j+1) as soon as warpgroup t holds all of S_t(j) in registers
          VITK_FSTAMP(16 * j + 8 + 3 * t + 0);
          mbar_wait(&bar_sfree[t], j & 1);
          tc_fence_after_sync();
          issue_s(t, j + 1);
        }
        mbar_wait(&bar_p[t], j & 1);
        tc_fence_after_sync();
        VITK_FSTAMP(16 * j + 8 + 3 * t + 1);                      // got P_t(j)
        const uint32_t tm_p = tmem_base + 256 + t * 64, tm_o = tmem_base + 384 + t * kDh;
        if (elect_one()) {
          if (ksteps == kKB / 16) {             // full block: unrolled, MMAs issue back to back
#pragma unroll
            for (int k = 0; k < kKB / 16; ++k)
              tc_mma_bf16_ts(tm_o, tm_p + k * 8, bv + k * 128, idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
          } else {
            for (int k = 0; k < ksteps; ++k)
              tc_mma_bf16_ts(tm_o, tm_p + k * 8, bv + k * 128, idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit(&bar_pv[t]);
          if (t == ntl - 1) tc_commit(&bar_vfree[st]);
          if (j == nblk - 1) tc_commit(&bar_o[t]);
        }
        VITK_FSTAMP(16 * j + 8 + 3 * t + 2);                      // S_t(j+1), P·V_t(j) issued
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    const int t = warp >> 2;                  // tile of this warpgroup
    if (t < ntl) {
      const int row = tid - t * kFwdWG;       // query row inside the tile = TMEM lane
      const uint32_t lane_field = static_cast<uint32_t>((warp & 3) * 32) << 16;
      const uint32_t tm_s = tmem_base + lane_field + t * 128;
      const uint32_t tm_p = tmem_base + lane_field + 256 + t * 64;
      const uint32_t tm_o = tmem_base + lane_field + 384 + t * kDh;
      bool p_free = true;                     // P·V_t of the previous block has finished reading the P columns
      float m_ref = -INFINITY, l_run = 0.f;
      for (int j = 0; j < nblk; ++j) {
        const int nvalid = min(kKB, T - j * kKB);
        const int nchunk = (nvalid + 31) >> 5;
        if ((warp & 3) == 0) VITK_FSTAMP(16 * j + 4 * t + 0);      // waiting for S_t(j)
        mbar_wait(&bar_s[t], j & 1);
        tc_fence_after_sync();
        if ((warp & 3) == 0) VITK_FSTAMP(16 * j + 4 * t + 1);      // S_t(j) ready
        float rs0 = 0.f, rs1 = 0.f;
        // one 32-column chunk of the row: lazy-maximum bookkeeping, exp2, bf16 pack, P written over consumed S columns
        auto chunk = [&](uint32_t (&r)[32], int c) {
          const bool full = (c + 1) * 32 <= nvalid;
          // Probabilities are computed SPECULATIVELY against the current reference maximum while the chunk maximum is
          // being reduced (two independent instruction streams: the MUFU pipe starts at once instead of after a
          // 16-deep max chain + vote); the reference moves only on the first chunk of a row and for extreme logits,
          // and then the chunk is simply recomputed from the registers it still sits in.
          uint32_t pk[16];
          float s0 = 0.f, s1 = 0.f;
          float cm = -INFINITY;
          if (full) {
            softmax_pairs<0>(r, scale_log2, m_ref, s0, s1, pk);
#pragma unroll
            for (int i = 0; i < 32; i += 2) cm = fmax3(cm, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i < nvalid) cm = fmaxf(cm, __uint_as_float(r[i]));
          }
          const float m_c = cm * scale_log2;
          const bool fix = m_c > m_ref + kLazyMaxLog2;
          const bool moved = __any_sync(0xffffffffu, fix);
          if (moved) {
            const float m_new = fix ? m_c : m_ref;
            const float alpha = fix ? fast_exp2(m_ref - m_new) : 1.0f;
            l_run *= alpha;
            rs0 *= alpha;
            rs1 *= alpha;
            tmem_ld_wait();                         // a prefetched chunk may be in flight: let it land first
            tmem_st_wait();                         // P chunks of this block written so far
            if (j > 0) {
              // O_t holds P·V of blocks < j; P·V_t(j-1) may still be in flight
              if (!p_free) {
                mbar_wait(&bar_pv[t], (j - 1) & 1);
                tc_fence_after_sync();
                p_free = true;
              }
#pragma unroll 1
              for (int cc = 0; cc < 2; ++cc) {
                uint32_t ro[32];
                tmem_ld_32x32(tm_o + cc * 32, ro);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) ro[i] = __float_as_uint(__uint_as_float(ro[i]) * alpha);
                tmem_st_32x32(tm_o + cc * 32, ro);
              }
            }
#pragma unroll 1
            for (int cc = 0; cc < c; ++cc) {        // P chunks of this block already written
              uint32_t pq[16];
              tmem_ld_32x16(tm_p + cc * 16, pq);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pq[i]));
                pq[i] = pack_bf16x2(f.x * alpha, f.y * alpha);
              }
              tmem_st_32x16(tm_p + cc * 16, pq);
            }
            m_ref = m_new;
          }
          if (moved || !full) {
            s0 = s1 = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float p0 = fast_exp2(fmaf(__uint_as_float(r[2 * i]), scale_log2, -m_ref));
              float p1 = fast_exp2(fmaf(__uint_as_float(r[2 * i + 1]), scale_log2, -m_ref));
              if (c * 32 + 2 * i >= nvalid) p0 = 0.f;
              if (c * 32 + 2 * i + 1 >= nvalid) p1 = 0.f;
              s0 += p0;
              s1 += p1;
              pk[i] = pack_bf16x2(p0, p1);
            }
          }
          rs0 += s0;
          rs1 += s1;
          if (!p_free) {                        // first P chunk of this block: P·V_t(j-1) must be done with the P columns
            mbar_wait(&bar_pv[t], (j - 1) & 1);
            tc_fence_after_sync();
            p_free = true;
          }
          tmem_st_32x16(tm_p + c * 16, pk);     // chunk c of P → P columns [16c, 16c+16)
        };
        auto release_s = [&]() {                // every lane of this warp has its part of S_t(j) in registers
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_sfree[t]);
        };
        // two register buffers: the next chunk's TMEM load is in flight during this chunk's math
        uint32_t ra[32], rb[32];
        tmem_ld_32x32(tm_s, ra);
        for (int c = 0; c < nchunk; c += 2) {
          tmem_ld_wait();
          tmem_regs_ready(ra);
          if (c + 1 < nchunk) tmem_ld_32x32(tm_s + (c + 1) * 32, rb); else release_s();
          chunk(ra, c);
          if (c + 1 < nchunk) {
            tmem_ld_wait();
            tmem_regs_ready(rb);
            if (c + 2 < nchunk) tmem_ld_32x32(tm_s + (c + 2) * 32, ra); else release_s();
            chunk(rb, c + 1);
          }
        }
        p_free = false;
        l_run += rs0 + rs1;
        if ((warp & 3) == 0) VITK_FSTAMP(16 * j + 4 * t + 2);      // math done
        tmem_st_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_p[t]);
        if ((warp & 3) == 0) VITK_FSTAMP(16 * j + 4 * t + 3);      // P_t(j) published
      }
      mbar_wait(&bar_o[t], 0);
      tc_fence_after_sync();
      if ((warp & 3) == 0) VITK_FSTAMP(201 + 2 * t);                  // last P·V_t retired
      const int tq = (tile0 + t) * kTile + row;
      const float inv = 1.0f / l_run;
      // O_t / l → bf16 → Q_t's shared-memory tile (every S_t MMA that read it has retired), 128-byte swizzled rows
      uint8_t* srow = sQ + t * kTileBytes + row * 128;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t ro[32];
        tmem_ld_32x32(tm_o + c * 32, ro);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(ro[8 * q + 0]) * inv, __uint_as_float(ro[8 * q + 1]) * inv);
          w.y = pack_bf16x2(__uint_as_float(ro[8 * q + 2]) * inv, __uint_as_float(ro[8 * q + 3]) * inv);
          w.z = pack_bf16x2(__uint_as_float(ro[8 * q + 4]) * inv, __uint_as_float(ro[8 * q + 5]) * inv);
          w.w = pack_bf16x2(__uint_as_float(ro[8 * q + 6]) * inv, __uint_as_float(ro[8 * q + 7]) * inv);
          *reinterpret_cast<uint4*>(srow + (((4 * c + q) ^ (row & 7)) << 4)) = w;
        }
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync %0, %1;" ::"r"(1 + t), "r"(kFwdWG) : "memory");
      if (row == 0) {
        tma_store_3d(&tma_o, sQ + t * kTileBytes, colq, (tile0 + t) * kTile, b);
        tma_store_commit();
        tma_store_wait_read<0>();
      }
      if (tq < T) lse[(static_cast<long long>(b) * H + h) * T + tq] = (m_ref + log2f(l_run)) * kLn2;
      if ((warp & 3) == 0) VITK_FSTAMP(202 + 2 * t);                  // O_t stored
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, kFwdTmemCols);
  }
  if (tl != nullptr && threadIdx.x == 0 && lin_cta < 2000) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    tl[3 * lin_cta + 2] = static_cast<long long>(t);
  }
}

// ============================================================================ backward
// stats[b,h,t] = { lse·log2e, Δ = Σ_d dO·O } for t < T and { +inf, 0 } for the padding rows T ≤ t < Tpad
// (Tpad = multiple of 128), so the main kernel reads them with unconditional aligned float4 loads and
// padded queries get P = exp2(−inf) = 0.  One warp per token row.
__global__ void __launch_bounds__(256) attn_delta_kernel(const __nv_bfloat16* __restrict__ o,
                                                         const __nv_bfloat16* __restrict__ d_o, const float* __restrict__ lse,
                                                         int B, int T, int Tpad, int H, float* __restrict__ lse2,
                                                         float* __restrict__ delta) {
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);   // over B·Tpad
  const int lane = threadIdx.x & 31;
  if (row >= B * Tpad) return;
  const int b = row / Tpad, t = row - b * Tpad;
  if (t >= T) {
    for (int hh = lane; hh < H; hh += 32) {
      const long long idx = (static_cast<long long>(b) * H + hh) * Tpad + t;
      lse2[idx] = INFINITY;
      delta[idx] = 0.f;
    }
    return;
  }
  const int D = H * kDh;
  const long long grow = static_cast<long long>(b) * T + t;
  const uint4* po = reinterpret_cast<const uint4*>(o + grow * D);
  const uint4* pd = reinterpret_cast<const uint4*>(d_o + grow * D);
  for (int c0 = 0; c0 < D / 8; c0 += 32) {  // chunk c covers columns [8c, 8c+8) of head c/8
    const int c = c0 + lane;
    const bool valid = c < D / 8;
    float s = 0.f;
    if (valid) {
      const uint4 a = __ldg(po + c), g = __ldg(pd + c);
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&aw[i]));
        const float2 y = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gw[i]));
        s += x.x * y.x + x.y * y.y;
      }
    }
    // 8 consecutive lanes hold one head
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (valid && (lane & 7) == 0) {
      const int hh = c >> 3;
      const long long idx = (static_cast<long long>(b) * H + hh) * Tpad + t;
      delta[idx] = s;
      lse2[idx] = __ldg(lse + (static_cast<long long>(b) * H + hh) * T + t) * kLog2e;
    }
  }
}

// backward main kernel.  CTA = (128 keys, head, image) looping over the queries in sub-blocks of 64;
// 16 compute warps (TMEM lane quadrant = warp%4 = 32 keys, column group = warp/4 = 16 queries) + 1 control
// warp that owns TMA and MMA issue.  Per sub-block u (buffer x = u&1):
//   Sᵀ_x = K·Qᵀ, dPᵀ_x = V·dOᵀ       SS-MMAs (N = 64) into one of TWO score buffers in TMEM (lane = key)
//   Pᵀ = exp2(Sᵀ·c − lse), dSᵀ = Pᵀ∘(dPᵀ − Δ)   compute warps; both are written back (bf16) INTO the warp's own
//                                    consumed Sᵀ / dPᵀ columns with tcgen05.st, dSᵀ additionally to swizzled smem
//   dV += Pᵀ·dO, dK += dSᵀ·Q          TS-MMAs, A straight from TMEM (no shared-memory A reads)
//   dQ_i = dS·K                      once per 128-query block: SS-MMA, A = the dSᵀ smem tile read MN-major
// MMAs retire in issue order, so the scores of sub-block u+2 are queued right behind dV/dK(/dQ) of u and the
// tensor core works on buffer x while the compute warps work on buffer 1−x.  dQ is double-buffered in TMEM
// and reduced into the fp32 accumulator one block later with per-warp TMA reduce-adds.  Keys ≥ T need no
// masking (their dV/dK rows are never stored, their dQ contribution multiplies zero-filled K rows);
// queries ≥ T have lse = +inf in the padded statistics, hence P = dS = 0.
constexpr int kBwdComputeWarps = 16;
constexpr int kBwdThreads = (kBwdComputeWarps + 2) * 32;   // 16 compute warps + MMA-issue warp + TMA-load warp
constexpr int kQSub = 64;
constexpr int kBwdSmemK = 0;
constexpr int kBwdSmemV = kBwdSmemK + kTileBytes;
constexpr int kBwdSmemQ = kBwdSmemV + kTileBytes;        // ring of 4 sub-tiles of 64 queries (8 KB each)
constexpr int kBwdSmemDO = kBwdSmemQ + 2 * kTileBytes;   // ring of 4 sub-tiles
constexpr int kBwdSmemDSt = kBwdSmemDO + 2 * kTileBytes; // dSᵀ [128 keys × 128 q], one tile per 128-query block parity
constexpr int kBwdSmemDqS = kBwdSmemDSt + 4 * kTileBytes;   // per-warp dQ staging slabs: 16 × [32 q × 16 f32], 64 B swizzle
constexpr int kBwdSmemBar = kBwdSmemDqS + kBwdComputeWarps * 2048;
constexpr int kBwdSmemBytes = kBwdSmemBar + 128 + 1024;
// TMEM columns: buffer x: Sᵀ/Pᵀ [128x, 128x+64), dPᵀ/dSᵀ [128x+64, 128x+128); dV [256,320) dK [320,384) dQ[2] [384,512)
constexpr int kBwdTmemCols = 512;

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_q64,
                const __grid_constant__ CUtensorMap tma_do64, const __grid_constant__ CUtensorMap tma_dq, const float* __restrict__ lse2, const float* __restrict__ delta,
                __nv_bfloat16* __restrict__ dqkv, int T, int Tpad, int H, float scale, float scale_log2, long long* tl) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem + kBwdSmemK;
  uint8_t* sV = smem + kBwdSmemV;
  uint8_t* sQ = smem + kBwdSmemQ;
  uint8_t* sDO = smem + kBwdSmemDO;
  uint8_t* sDSt = smem + kBwdSmemDSt;
  uint64_t* bar_kv = reinterpret_cast<uint64_t*>(smem + kBwdSmemBar);
  uint64_t* bar_qd = bar_kv + 1;  // [4] Q / dO sub-tile landed
  uint64_t* bar_s = bar_kv + 5;   // [2] Sᵀ_x, dPᵀ_x ready
  uint64_t* bar_pd = bar_kv + 7;  // [2] Pᵀ_x / dSᵀ_x written by the 16 compute warps
  uint64_t* bar_g = bar_kv + 9;   // every MMA up to and including dQ of a 128-query block retired
  uint64_t* bar_free = bar_kv + 10;  // [4] dV/dK of the sub-block using this Q/dO sub-tile retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_kv + 14);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int nq = (T + kTile - 1) / kTile;
  const int nsub = (T + kQSub - 1) / kQSub;
  const int colq = h * kDh, colk = (H + h) * kDh, colv = (2 * H + h) * kDh;
  const int key0 = kb * kTile;

  if (tid == 0) {
    tma_prefetch_desc(&tma_qkv);
    tma_prefetch_desc(&tma_q64);
    tma_prefetch_desc(&tma_do64);
    tma_prefetch_desc(&tma_dq);
    for (int i = 0; i < 14; ++i) mbar_init(bar_kv + i, (i == 7 || i == 8) ? kBwdComputeWarps : 1);
    fence_mbar_init();
  }
  if (warp == kBwdComputeWarps) tmem_alloc(tmem_slot, kBwdTmemCols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  pdl_wait();
  pdl_launch_dependents();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // broadcast marks it warp-uniform: no per-MMA elect-broadcast-retry loop
  const uint32_t tm_dv = tmem_base + 256, tm_dk = tmem_base + 320, tm_dq = tmem_base + 384;

  if (warp == kBwdComputeWarps) {
    // ------------------------------------------------------------------ control warp: TMA + MMA issue
    // (all lanes run the loop so descriptors stay in uniform registers; the issuing instructions sit under
    // elect.sync so that ptxas emits them back to back instead of one elect/broadcast/retry loop per instruction)
    constexpr uint32_t idesc_sc = umma_idesc_bf16(kTile, kQSub, 0, 0);  // Sᵀ = K·Qᵀ, dPᵀ = V·dOᵀ (N = 64 queries)
    constexpr uint32_t idesc_km = umma_idesc_bf16(kTile, kDh, 0, 1);    // dV += Pᵀ·dO, dK += dSᵀ·Q
    constexpr uint32_t idesc_mm = umma_idesc_bf16(kTile, kDh, 1, 1);    // dQ = dS·K
    constexpr uint32_t kTile16 = kTileBytes >> 4;                        // descriptor address units are 16 B
    const uint64_t k_kmaj = umma_smem_desc(smem_u32(sK), 0, 1024);
    const uint64_t v_kmaj = umma_smem_desc(smem_u32(sV), 0, 1024);
    const uint64_t k_mnmaj = umma_smem_desc(smem_u32(sK), kTileBytes, 1024);
    constexpr uint32_t kSub16 = (kQSub * 128) >> 4;                      // one 64-query sub-tile = 8 KB
    const uint64_t q_kmaj = umma_smem_desc(smem_u32(sQ), 0, 1024);
    const uint64_t do_kmaj = umma_smem_desc(smem_u32(sDO), 0, 1024);
    const uint64_t q_mnmaj = umma_smem_desc(smem_u32(sQ), kQSub * 128, 1024);
    const uint64_t do_mnmaj = umma_smem_desc(smem_u32(sDO), kQSub * 128, 1024);
    const uint64_t dst_mnmaj = umma_smem_desc(smem_u32(sDSt), kTileBytes, 1024);
    auto issue_scores = [&](int u) {
      const int x = u & 1;
      const uint64_t boff = static_cast<uint64_t>((u & 3) * kSub16);
      mbar_wait(&bar_qd[u & 3], (u >> 2) & 1);
      tc_fence_after_sync();
      const uint32_t t_s = tmem_base + x * 128, t_dp = t_s + 64;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kDh / 16; ++k) tc_mma_bf16(t_s, k_kmaj + 2 * k, q_kmaj + boff + 2 * k, idesc_sc, k > 0);
#pragma unroll
        for (int k = 0; k < kDh / 16; ++k) tc_mma_bf16(t_dp, v_kmaj + 2 * k, do_kmaj + boff + 2 * k, idesc_sc, k > 0);
        tc_commit(&bar_s[x]);
      }
    };
    mbar_wait(bar_kv, 0);
    issue_scores(0);
    if (nsub > 1) issue_scores(1);
    for (int u = 0; u < nsub; ++u) {
      const int i = u >> 1, hq = u & 1, x = u & 1;
      const uint64_t boff = static_cast<uint64_t>((u & 3) * kSub16);
      mbar_wait(&bar_pd[x], (u >> 1) & 1);
      VITK_STAMP(16 * u + 0);
      tc_fence_after_sync();
      const uint32_t t_p = tmem_base + x * 128, t_ds = t_p + 64;
      const bool block_done = hq == 1 || u == nsub - 1;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kQSub / 16; ++k)   // dV[key,d] += Σ_q Pᵀ[key,q]·dO[q,d]; Pᵀ of column group k sits at Sᵀ column 16k
          tc_mma_bf16_ts(tm_dv, t_p + 16 * k, do_mnmaj + boff + 128 * k, idesc_km, (u > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < kQSub / 16; ++k)   // dK[key,d] += Σ_q dSᵀ[key,q]·Q[q,d]
          tc_mma_bf16_ts(tm_dk, t_ds + 16 * k, q_mnmaj + boff + 128 * k, idesc_km, (u > 0 || k > 0) ? 1u : 0u);
        tc_commit(&bar_free[u & 3]);
        if (block_done) {
#pragma unroll
          for (int k = 0; k < kTile / 16; ++k)   // dQ[q,d] = Σ_key dS[q,key]·K[key,d]; A = dSᵀ smem tile read MN-major
            tc_mma_bf16(tm_dq + (i & 1) * kDh, dst_mnmaj + static_cast<uint64_t>((i & 1) * 2 * kTile16) + 128 * k,
                        k_mnmaj + 128 * k, idesc_mm, k > 0);
          tc_commit(bar_g);
        }
      }
      VITK_STAMP(16 * u + 1);
      if (u + 2 < nsub) issue_scores(u + 2);   // queued right behind: MMAs retire in issue order
      VITK_STAMP(16 * u + 2);
      VITK_STAMP(16 * u + 3);
    }
    __syncwarp();
  } else if (warp == kBwdComputeWarps + 1) {
    // ------------------------------------------------------------------ load warp: K/V once, then the Q/dO ring.
    // (A separate warp: waiting here for a ring slot to be released — i.e. for dV/dK MMAs to retire — must not
    // hold up the MMA-issue warp, which would stall the whole score → softmax → dV/dK chain of the other buffer.)
    auto load_qd = [&](int u) {              // sub-block u → ring slot u&3
      const int slot = u & 3;
      if (elect_one()) {
        mbar_arrive_expect_tx(&bar_qd[slot], 2 * kQSub * 128);
        tma_load_3d(sQ + slot * (kQSub * 128), &tma_q64, &bar_qd[slot], colq, u * kQSub, b);
        tma_load_3d(sDO + slot * (kQSub * 128), &tma_do64, &bar_qd[slot], colq, u * kQSub, b);
      }
    };
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_kv, 2 * kTileBytes);
      tma_load_3d(sK, &tma_qkv, bar_kv, colk, key0, b);
      tma_load_3d(sV, &tma_qkv, bar_kv, colv, key0, b);
    }
    for (int u = 0; u < 4 && u < nsub; ++u) load_qd(u);
    for (int u = 4; u < nsub; ++u) {         // slot u&3 was last used by sub-block u−4: wait until its dV/dK retired
      mbar_wait(&bar_free[u & 3], ((u - 4) >> 2) & 1);
      load_qd(u);
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ compute warps
    const int quad = warp & 3, cg = warp >> 2;
    const int key_row = quad * 32 + lane;  // TMEM lane = key (Sᵀ, dPᵀ, dV, dK) or query (dQ)
    const uint32_t lane_field = static_cast<uint32_t>(quad * 32) << 16;
    const float* stat_lse = lse2 + (static_cast<long long>(b) * H + h) * Tpad + cg * 16;
    const float* stat_dlt = delta + (static_cast<long long>(b) * H + h) * Tpad + cg * 16;
    // this warp's 32 query rows × 16 head-dim columns of dQ_i: TMEM → swizzled slab → one TMA reduce-add into the
    // fp32 accumulator (rows ≥ T are clipped by the tensor map)
    uint8_t* dq_slab = smem + kBwdSmemDqS + warp * 2048;
    auto reduce_dq = [&](int i) {
      uint32_t r[16];
      tmem_ld_32x16(tm_dq + (i & 1) * kDh + lane_field + cg * 16, r);
      if (elect_one()) tma_store_wait_read<0>();   // previous block's reduce has finished reading the slab
      __syncwarp();
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(dq_slab + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) =
            make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (elect_one()) {
        tma_reduce_add_3d(&tma_dq, dq_slab, h * kDh + cg * 16, i * kTile + quad * 32, b);
        tma_store_commit();
      }
    };
    // statistics of this warp's 16 queries (padded arrays: unconditional aligned loads), fetched one sub-block ahead
    float4 l4[4], d4[4];
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      l4[y] = __ldg(reinterpret_cast<const float4*>(stat_lse) + y);
      d4[y] = __ldg(reinterpret_cast<const float4*>(stat_dlt) + y);
    }
    for (int u = 0; u < nsub; ++u) {
      const int i = u >> 1, hq = u & 1, x = u & 1;
      if (warp == 0) VITK_STAMP(16 * u + 8);
      mbar_wait(&bar_s[x], (u >> 1) & 1);
      if (warp == 0) VITK_STAMP(16 * u + 9);
      tc_fence_after_sync();
      const uint32_t t_s = tmem_base + x * 128 + lane_field + cg * 16, t_dp = t_s + 64;
      uint32_t rs[16], rp[16];
      tmem_ld_32x16(t_s, rs);
      tmem_ld_32x16(t_dp, rp);
      tmem_ld_wait();
      if (warp == 0) VITK_STAMP(16 * u + 10);
      uint32_t pk[8], dk[8];
#pragma unroll
      for (int y = 0; y < 4; ++y) {
        const float lv[4] = {l4[y].x, l4[y].y, l4[y].z, l4[y].w}, dv[4] = {d4[y].x, d4[y].y, d4[y].z, d4[y].w};
        float p[4], ds[4];
#pragma unroll
        for (int z = 0; z < 4; ++z) {
          p[z] = fast_exp2(fmaf(__uint_as_float(rs[4 * y + z]), scale_log2, -lv[z]));
          ds[z] = p[z] * (__uint_as_float(rp[4 * y + z]) - dv[z]);
        }
        pk[2 * y] = pack_bf16x2(p[0], p[1]);
        pk[2 * y + 1] = pack_bf16x2(p[2], p[3]);
        dk[2 * y] = pack_bf16x2(ds[0], ds[1]);
        dk[2 * y + 1] = pack_bf16x2(ds[2], ds[3]);
      }
      if (u + 1 < nsub) {           // next sub-block's statistics: in flight during the stores, fences and the next wait
#pragma unroll
        for (int y = 0; y < 4; ++y) {
          l4[y] = __ldg(reinterpret_cast<const float4*>(stat_lse + (u + 1) * kQSub) + y);
          d4[y] = __ldg(reinterpret_cast<const float4*>(stat_dlt + (u + 1) * kQSub) + y);
        }
      }
      if (warp == 0) VITK_STAMP(16 * u + 11);
      tmem_st_32x8(t_s, pk);        // Pᵀ  → this warp's own (consumed) Sᵀ columns: A operand of dV
      tmem_st_32x8(t_dp, dk);       // dSᵀ → this warp's own dPᵀ columns: A operand of dK
      {                             // dSᵀ → smem tile [128 keys × 128 q] (two 64-query halves), A operand of dQ
        // (double-buffered by block parity: dQ of block i may still be pending when block i+1's first half is written)
        uint8_t* rowp = sDSt + (i & 1) * 2 * kTileBytes + hq * kTileBytes + key_row * 128;
        *reinterpret_cast<uint4*>(rowp + (((2 * cg) ^ (key_row & 7)) << 4)) = make_uint4(dk[0], dk[1], dk[2], dk[3]);
        *reinterpret_cast<uint4*>(rowp + (((2 * cg + 1) ^ (key_row & 7)) << 4)) = make_uint4(dk[4], dk[5], dk[6], dk[7]);
      }
      fence_proxy_async_smem();
      tmem_st_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (warp == 0) VITK_STAMP(16 * u + 12);
      if (lane == 0) mbar_arrive(&bar_pd[x]);
      if (hq == 1 && i > 0) {       // dQ of the previous 128-query block: its MMAs were issued two sub-blocks ago
        mbar_wait(bar_g, (i - 1) & 1);
        tc_fence_after_sync();
        reduce_dq(i - 1);
      }
      if (warp == 0) VITK_STAMP(16 * u + 13);
    }
    if (nq > 1 && (nsub & 1)) {     // an odd tail sub-block skipped the hq == 1 step that drains block nq−2
      mbar_wait(bar_g, (nq - 2) & 1);
      tc_fence_after_sync();
      reduce_dq(nq - 2);
    }
    mbar_wait(bar_g, (nq - 1) & 1);
    tc_fence_after_sync();
    reduce_dq(nq - 1);
    if (elect_one()) tma_store_wait_all<0>();
    // dV (cg 0,1) and dK·scale (cg 2,3) → dqkv[b, key, 2|1, h, :]
    {
      const bool is_dv = cg < 2;
      const int c32 = (cg & 1) * 32;
      uint32_t r[32];
      tmem_ld_32x32((is_dv ? tm_dv : tm_dk) + lane_field + c32, r);
      tmem_ld_wait();
      const float mul = is_dv ? 1.0f : scale;
      const int key = key0 + key_row;
      if (key < T) {
        uint4* dst = reinterpret_cast<uint4*>(dqkv + ((static_cast<long long>(b) * T + key) * 3 + (is_dv ? 2 : 1)) * H * kDh +
                                              h * kDh + c32);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(r[8 * q + 0]) * mul, __uint_as_float(r[8 * q + 1]) * mul);
          w.y = pack_bf16x2(__uint_as_float(r[8 * q + 2]) * mul, __uint_as_float(r[8 * q + 3]) * mul);
          w.z = pack_bf16x2(__uint_as_float(r[8 * q + 4]) * mul, __uint_as_float(r[8 * q + 5]) * mul);
          w.w = pack_bf16x2(__uint_as_float(r[8 * q + 6]) * mul, __uint_as_float(r[8 * q + 7]) * mul);
          dst[q] = w;
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kBwdComputeWarps) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, kBwdTmemCols);
  }
}

// dq_acc fp32 [B,T,H,64] · scale → dqkv[b,t,0,h,:] bf16
__global__ void __launch_bounds__(256) attn_dq_store_kernel(const float* __restrict__ dq_acc, int BT, int H, float scale,
                                                            __nv_bfloat16* __restrict__ dqkv) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i8 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;  // 8 elements each
  const int D = H * kDh;
  const long long total = static_cast<long long>(BT) * D / 8;
  if (i8 >= total) return;
  const long long row = (i8 * 8) / D;
  const int col = static_cast<int>((i8 * 8) - row * D);
  const float4 a = __ldg(reinterpret_cast<const float4*>(dq_acc) + 2 * i8);
  const float4 c = __ldg(reinterpret_cast<const float4*>(dq_acc) + 2 * i8 + 1);
  uint4 w;
  w.x = pack_bf16x2(a.x * scale, a.y * scale);
  w.y = pack_bf16x2(a.z * scale, a.w * scale);
  w.z = pack_bf16x2(c.x * scale, c.y * scale);
  w.w = pack_bf16x2(c.z * scale, c.w * scale);
  *reinterpret_cast<uint4*>(dqkv + row * 3 * D + col) = w;
}

static int qkv_tensor_map(CUtensorMap* m, const void* base, int64_t B, int64_t T, int64_t row_elems, int rows = kTile) {
  const uint64_t dims[3] = {static_cast<uint64_t>(row_elems), static_cast<uint64_t>(T), static_cast<uint64_t>(B)};
  const uint64_t str[2] = {static_cast<uint64_t>(row_elems) * 2, static_cast<uint64_t>(T) * row_elems * 2};
  const uint32_t box[3] = {kDh, static_cast<uint32_t>(rows), 1};
  return make_tensor_map_bf16(m, base, 3, dims, str, box);
}

static int check_shape(const char* who, int64_t B, int64_t T, int64_t H) {
  VITK_REQUIRE(B > 0 && T > 0 && H > 0 && B < 65536 && H < 65536 && T < (1 << 24), VITK_EINVAL,
               "%s: unsupported shape B=%lld T=%lld H=%lld", who, (long long)B, (long long)T, (long long)H);
  return 0;
}

}  // namespace vitk

using namespace vitk;

extern "C" VITK_API int vitk_attn_fwd(const void* qkv, int64_t B, int64_t T, int64_t H, float scale, void* o, float* lse,
                             vitk_stream_t stream) {
  VITK_REQUIRE(qkv && o && lse, VITK_EINVAL, "attn_fwd: NULL argument");
  if (int rc = check_shape("attn_fwd", B, T, H)) return rc;
  VITK_REQUIRE(aligned16(qkv) && aligned16(o), VITK_EALIGN, "attn_fwd: buffers must be 16-byte aligned");
  VITK_REQUIRE(scale > 0.f, VITK_EINVAL, "attn_fwd: scale must be positive");
  CUtensorMap tm, tm_o;      // one box shape (128 rows × 64 columns) serves Q tiles, K/V blocks and the O store
  if (int rc = qkv_tensor_map(&tm, qkv, B, T, 3 * H * kDh)) return rc;
  if (int rc = qkv_tensor_map(&tm_o, o, B, T, H * kDh)) return rc;
  static std::atomic<int> attr_done{0};
  if (!attr_done.load()) {
    VITK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmemBytes));
    attr_done.store(1);
  }
  const int64_t ntile = (T + kTile - 1) / kTile;
  const int64_t units = B * H * ((ntile >> 1) + (ntile & 1));
  VITK_REQUIRE(units < (1LL << 31), VITK_EINVAL, "attn_fwd: too many work units");
  VITK_CUDA(launch_pdl(attn_fwd_kernel, dim3(static_cast<unsigned>(units)), dim3(kFwdThreads), kFwdSmemBytes,
                       static_cast<cudaStream_t>(stream), tm, tm, tm_o, lse, static_cast<int>(T),
                       static_cast<int>(H), static_cast<int>(B * H), scale * kLog2e, g_timeline));
  VITK_LAUNCH_CHECK("attn_fwd_kernel");
  return 0;
}

extern "C" VITK_API int vitk_debug_timeline(void* device_buf) {
  g_timeline = static_cast<long long*>(device_buf);
  g_timeline_seq = 0;
  return 0;
}

__global__ void debug_stamp_kernel(long long* dst) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  *dst = static_cast<long long>(t);
}
extern "C" VITK_API int vitk_debug_stamp(int64_t slot, void* stream) {
  VITK_REQUIRE(g_timeline != nullptr && slot >= 0, VITK_EINVAL, "vitk_debug_stamp: no timeline buffer set");
  static const int carve = [] {     // experiment: VITK_STAMP_CARVEOUT=100 keeps the SM in its max-shared-memory configuration
    const char* e = getenv("VITK_STAMP_CARVEOUT");
    const int c = e ? atoi(e) : -1;
    if (c >= 0) cudaFuncSetAttribute(debug_stamp_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, c);
    return c;
  }();
  (void)carve;
  debug_stamp_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(g_timeline + slot);
  VITK_LAUNCH_CHECK("debug_stamp_kernel");
  return 0;
}

static size_t bwd_stats_bytes(int64_t B, int64_t T, int64_t H) {
  const int64_t Tpad = (T + kTile - 1) / kTile * kTile;
  return (static_cast<size_t>(B) * H * Tpad * sizeof(float) + 255) / 256 * 256;
}
static size_t bwd_dq_bytes(int64_t B, int64_t T, int64_t H) {
  return (static_cast<size_t>(B) * T * H * kDh * sizeof(float) + 255) / 256 * 256;
}
// workspace = fp32 dQ accumulator [B,T,H,64] | lse·log2e [B,H,Tpad] | Δ [B,H,Tpad]
extern "C" VITK_API size_t vitk_attn_bwd_workspace_bytes(int64_t B, int64_t T, int64_t H) {
  return bwd_dq_bytes(B, T, H) + 2 * bwd_stats_bytes(B, T, H);
}

extern "C" VITK_API int vitk_attn_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, int64_t B, int64_t T,
                             int64_t H, float scale, void* dqkv, void* workspace, vitk_stream_t stream) {
  VITK_REQUIRE(qkv && o && d_o && lse && dqkv && workspace, VITK_EINVAL, "attn_bwd: NULL argument");
  if (int rc = check_shape("attn_bwd", B, T, H)) return rc;
  VITK_REQUIRE(aligned16(qkv) && aligned16(o) && aligned16(d_o) && aligned16(dqkv) && aligned16(workspace), VITK_EALIGN,
               "attn_bwd: buffers must be 16-byte aligned");
  VITK_REQUIRE(scale > 0.f, VITK_EINVAL, "attn_bwd: scale must be positive");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t dq_bytes = bwd_dq_bytes(B, T, H);
  const int Tpad = static_cast<int>((T + kTile - 1) / kTile * kTile);
  float* dq_acc = static_cast<float*>(workspace);
  float* lse2 = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + dq_bytes);
  float* delta = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + dq_bytes + bwd_stats_bytes(B, T, H));
  CUtensorMap tm_qkv, tm_q64, tm_do, tm_dq;
  if (int rc = qkv_tensor_map(&tm_qkv, qkv, B, T, 3 * H * kDh)) return rc;
  if (int rc = qkv_tensor_map(&tm_q64, qkv, B, T, 3 * H * kDh, kQSub)) return rc;
  if (int rc = qkv_tensor_map(&tm_do, d_o, B, T, H * kDh, kQSub)) return rc;
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(H * kDh), static_cast<uint64_t>(T), static_cast<uint64_t>(B)};
    const uint64_t str[2] = {static_cast<uint64_t>(H * kDh) * 4, static_cast<uint64_t>(T) * H * kDh * 4};
    const uint32_t box[3] = {16, 32, 1};
    if (int rc = get_tensor_map(&tm_dq, dq_acc, TM_F32, 3, dims, str, box, TM_SW64)) return rc;
  }
  static std::atomic<int> attr_done{0};
  if (!attr_done.load()) {
    VITK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmemBytes));
    attr_done.store(1);
  }
  VITK_CUDA(cudaMemsetAsync(dq_acc, 0, dq_bytes, s));
  const int BT = static_cast<int>(B * T);
  VITK_CUDA(launch_pdl(attn_delta_kernel, dim3((static_cast<int>(B) * Tpad + 7) / 8), dim3(256), 0, s,
                       static_cast<const __nv_bfloat16*>(o), static_cast<const __nv_bfloat16*>(d_o), lse, (int)B, (int)T, Tpad,
                       (int)H, lse2, delta));
  VITK_LAUNCH_CHECK("attn_delta_kernel");
  const dim3 grid(static_cast<unsigned>((T + kTile - 1) / kTile), static_cast<unsigned>(H), static_cast<unsigned>(B));
  VITK_CUDA(launch_pdl(attn_bwd_kernel, grid, dim3(kBwdThreads), kBwdSmemBytes, s, tm_qkv, tm_q64, tm_do, tm_dq,
                       static_cast<const float*>(lse2), static_cast<const float*>(delta), static_cast<__nv_bfloat16*>(dqkv), (int)T, Tpad,
                       (int)H, scale, scale * kLog2e, g_timeline));
  VITK_LAUNCH_CHECK("attn_bwd_kernel");
  const long long n8 = static_cast<long long>(BT) * H * kDh / 8;
  VITK_CUDA(launch_pdl(attn_dq_store_kernel, dim3(static_cast<unsigned>((n8 + 255) / 256)), dim3(256), 0, s,
                       static_cast<const float*>(dq_acc), BT, (int)H, scale, static_cast<__nv_bfloat16*>(dqkv)));
  VITK_LAUNCH_CHECK("attn_dq_store_kernel");
  return 0;
}
