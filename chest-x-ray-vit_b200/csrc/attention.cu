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
// forward  : CTA = (128 queries, head, image), 128 threads = 128 TMEM lanes = query rows,
//            two CTAs per SM so one CTA's softmax overlaps the other's MMAs.
// backward : CTA = (128 keys, head, image) loops over query blocks; Sᵀ and dPᵀ live in TMEM
//            (lane = key), dV/dK accumulate in TMEM across the loop, dQ partials are reduced
//            into an fp32 workspace with red.global.add.
#include <cuda.h>
#include <math.h>

#include <atomic>

#include "common.cuh"
#include "sm100_prims.cuh"
#include "tmap.cuh"

namespace vitk {

constexpr int kTile = 128;   // queries per CTA (fwd) / keys per CTA (bwd) / tokens per inner block
constexpr int kDh = 64;
constexpr int kTileBytes = kTile * kDh * 2;  // 16 KB
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
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
// CTA = (128 queries, head, image); 4 softmax warps (thread = query row = TMEM lane) + 1 control warp
// that owns TMA and MMA issue.  Two CTAs per SM.
//   keys are consumed in sub-blocks of 64 (two per 128-key TMA tile):
//   S_u = Q·K_uᵀ       SS-MMA (N = 64) into one of TWO S buffers in TMEM, so the tensor core computes
//                      S_{u+1} while the softmax warps work on S_u
//   P_u (bf16)         written back INTO the consumed S columns (two per 32-bit column) with tcgen05.st
//   O += P_u·V_u       TS-MMA: A = P straight from TMEM, B = V rows from smem (MN-major); O accumulates
//                      in TMEM across all sub-blocks — no per-block read-back, no P round trip via smem
// Softmax is single-pass with a lazily updated reference maximum: probabilities are taken relative to
// m_ref; whenever a 32-column chunk exceeds it by more than 2^16 (always on the first chunk, otherwise
// only for extreme logits) everything accumulated so far is rescaled by 2^(old−new) ≤ 1 — exact
// bookkeeping, m_ref cancels in O/l and in the LSE.  The last sub-block issues narrower MMAs
// (N, K rounded up to 16 valid keys) instead of computing masked columns.
constexpr int kFwdSoftmaxThreads = 128;
constexpr int kFwdThreads = kFwdSoftmaxThreads + 32;
constexpr int kSub = 64;                                   // keys per sub-block = one K/V TMA tile
constexpr int kSubBytes = kSub * kDh * 2;                  // 8 KB
constexpr int kFwdSmemQ = 0;
constexpr int kFwdSmemK = kFwdSmemQ + kTileBytes;          // 2 buffers of 64 keys
constexpr int kFwdSmemV = kFwdSmemK + 2 * kSubBytes;       // 2 buffers
constexpr int kFwdSmemBar = kFwdSmemV + 2 * kSubBytes;
constexpr int kFwdSmemBytes = kFwdSmemBar + 128 + 1024;    // ≈ 49 KB → four CTAs per SM
constexpr int kFwdTmemCols = 128;  // S/P: [0,64)   O: [64,128)
constexpr float kLazyMaxLog2 = 16.0f;

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

__global__ void __launch_bounds__(kFwdThreads, 4)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_kv,
                __nv_bfloat16* __restrict__ o, float* __restrict__ lse, int T, int H, float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + kFwdSmemQ;
  uint8_t* sK = smem + kFwdSmemK;
  uint8_t* sV = smem + kFwdSmemV;
  uint64_t* bar_q = reinterpret_cast<uint64_t*>(smem + kFwdSmemBar);
  uint64_t* bar_k = bar_q + 1;      // [2] K tile landed
  uint64_t* bar_v = bar_q + 3;      // [2] V tile landed
  uint64_t* bar_free = bar_q + 5;   // [2] K/V buffer consumed by its P·V
  uint64_t* bar_s = bar_q + 7;      // S ready
  uint64_t* bar_p = bar_q + 8;      // P written by the 4 softmax warps
  uint64_t* bar_o = bar_q + 9;      // last P·V retired
  uint64_t* bar_pv = bar_q + 10;    // P·V(u) retired (phase u): only the rescale path waits on it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_q + 11);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int nsub = (T + kSub - 1) / kSub;
  const int colq = h * kDh, colk = (H + h) * kDh, colv = (2 * H + h) * kDh;

  if (tid == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_kv);
    for (int i = 0; i < 11; ++i) mbar_init(bar_q + i, i == 8 ? 4 : 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, kFwdTmemCols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 64;

  if (warp == 4) {
    // ------------------------------------------------------------------ control warp: TMA + MMA issue
    if (lane == 0) {
      auto load_kv = [&](int u) {
        const int buf = u & 1;
        mbar_arrive_expect_tx(&bar_k[buf], kSubBytes);
        tma_load_3d(sK + buf * kSubBytes, &tma_kv, &bar_k[buf], colk, u * kSub, b);
        mbar_arrive_expect_tx(&bar_v[buf], kSubBytes);
        tma_load_3d(sV + buf * kSubBytes, &tma_kv, &bar_v[buf], colv, u * kSub, b);
      };
      mbar_arrive_expect_tx(bar_q, kTileBytes);
      tma_load_3d(sQ, &tma_q, bar_q, colq, qb * kTile, b);
      load_kv(0);
      if (nsub > 1) load_kv(1);
      auto sub_cols = [&](int u) { return (min(kSub, T - u * kSub) + 15) & ~15; };   // keys sub-block u needs
      auto issue_s = [&](int u) {
        const int buf = u & 1;
        mbar_wait(&bar_k[buf], (u >> 1) & 1);
        tc_fence_after_sync();
        const uint32_t aq = smem_u32(sQ), bk = smem_u32(sK + buf * kSubBytes);
        const uint32_t idesc_s = umma_idesc_bf16(kTile, sub_cols(u), 0, 0);
#pragma unroll
        for (int k = 0; k < kDh / 16; ++k)
          tc_mma_bf16(tmem_base, umma_smem_desc(aq + k * 32, 0, 1024), umma_smem_desc(bk + k * 32, 0, 1024), idesc_s, k > 0);
        tc_commit(bar_s);
      };
      mbar_wait(bar_q, 0);
      issue_s(0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(kTile, kDh, 0, 1);
      for (int u = 0; u < nsub; ++u) {
        const int buf = u & 1;
        const uint32_t ph = (u >> 1) & 1;
        mbar_wait(&bar_v[buf], ph);
        mbar_wait(bar_p, u & 1);
        tc_fence_after_sync();
        const uint32_t bv = smem_u32(sV + buf * kSubBytes);
        const int ksteps = sub_cols(u) / 16;
        for (int k = 0; k < ksteps; ++k)
          tc_mma_bf16_ts(tmem_o, tmem_base + k * 8, umma_smem_desc(bv + k * 2048, kSubBytes, 1024), idesc_pv,
                         (u > 0 || k > 0) ? 1u : 0u);
        tc_commit(bar_pv);
        tc_commit(&bar_free[buf]);
        if (u == nsub - 1) tc_commit(bar_o);
        if (u + 1 < nsub) issue_s(u + 1);     // queued right behind P·V(u): MMAs retire in issue order
        if (u + 2 < nsub) {                   // refill this K/V buffer once its P·V has retired
          mbar_wait(&bar_free[buf], ph);
          load_kv(u + 2);
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ softmax warps
    const uint32_t lane_field = static_cast<uint32_t>(warp * 32) << 16;
    float m_ref = -INFINITY, l_run = 0.f;
    for (int u = 0; u < nsub; ++u) {
      const int nvalid = min(kSub, T - u * kSub);
      const int nchunk = (nvalid + 31) >> 5;
      const uint32_t tm_s = tmem_base + lane_field;
      mbar_wait(bar_s, u & 1);
      tc_fence_after_sync();
      float rs0 = 0.f, rs1 = 0.f;
      for (int c = 0; c < nchunk; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tm_s + c * 32, r);
        tmem_ld_wait();
        const bool full = (c + 1) * 32 <= nvalid;
        float cm = -INFINITY;
        if (full) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) cm = fmax3(cm, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i < nvalid) cm = fmaxf(cm, __uint_as_float(r[i]));
        }
        const float m_c = cm * scale_log2;
        const bool fix = m_c > m_ref + kLazyMaxLog2;
        if (__any_sync(0xffffffffu, fix)) {
          const float m_new = fix ? m_c : m_ref;
          const float alpha = fix ? fast_exp2(m_ref - m_new) : 1.0f;
          l_run *= alpha;
          rs0 *= alpha;
          rs1 *= alpha;
          if (u > 0) {
            // O holds P·V of sub-blocks < u; P·V(u-1) may still be in flight
            mbar_wait(bar_pv, (u - 1) & 1);
            tc_fence_after_sync();
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
              uint32_t ro[32];
              tmem_ld_32x32(tmem_o + lane_field + cc * 32, ro);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) ro[i] = __float_as_uint(__uint_as_float(ro[i]) * alpha);
              tmem_st_32x32(tmem_o + lane_field + cc * 32, ro);
            }
          }
#pragma unroll 1
          for (int cc = 0; cc < c; ++cc) {
            uint32_t pk[16];
            tmem_ld_32x16(tm_s + cc * 16, pk);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk[i]));
              pk[i] = pack_bf16x2(f.x * alpha, f.y * alpha);
            }
            tmem_st_32x16(tm_s + cc * 16, pk);
          }
          m_ref = m_new;
        }
        uint32_t pk[16];
        if (full) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float p0 = fast_exp2(fmaf(__uint_as_float(r[2 * i]), scale_log2, -m_ref));
            const float p1 = fast_exp2(fmaf(__uint_as_float(r[2 * i + 1]), scale_log2, -m_ref));
            rs0 += p0;
            rs1 += p1;
            pk[i] = pack_bf16x2(p0, p1);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float p0 = fast_exp2(fmaf(__uint_as_float(r[2 * i]), scale_log2, -m_ref));
            float p1 = fast_exp2(fmaf(__uint_as_float(r[2 * i + 1]), scale_log2, -m_ref));
            if (c * 32 + 2 * i >= nvalid) p0 = 0.f;
            if (c * 32 + 2 * i + 1 >= nvalid) p1 = 0.f;
            rs0 += p0;
            rs1 += p1;
            pk[i] = pack_bf16x2(p0, p1);
          }
        }
        tmem_st_32x16(tm_s + c * 16, pk);     // P overwrites S columns already consumed
      }
      l_run += rs0 + rs1;
      tmem_st_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
    }
    mbar_wait(bar_o, 0);
    tc_fence_after_sync();
    const int t = qb * kTile + tid;
    const float inv = 1.0f / l_run;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_o + lane_field + c * 32, r);
      tmem_ld_wait();
      if (t < T) {
        uint4* dst = reinterpret_cast<uint4*>(o + ((static_cast<long long>(b) * T + t) * H + h) * kDh + c * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(r[8 * q + 0]) * inv, __uint_as_float(r[8 * q + 1]) * inv);
          w.y = pack_bf16x2(__uint_as_float(r[8 * q + 2]) * inv, __uint_as_float(r[8 * q + 3]) * inv);
          w.z = pack_bf16x2(__uint_as_float(r[8 * q + 4]) * inv, __uint_as_float(r[8 * q + 5]) * inv);
          w.w = pack_bf16x2(__uint_as_float(r[8 * q + 6]) * inv, __uint_as_float(r[8 * q + 7]) * inv);
          dst[q] = w;
        }
      }
    }
    if (t < T) lse[(static_cast<long long>(b) * H + h) * T + t] = (m_ref + log2f(l_run)) * kLn2;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, kFwdTmemCols);
  }
}

// ============================================================================ backward
// delta[b,h,t] = Σ_d dO·O  (one warp per token row)
__global__ void __launch_bounds__(256) attn_delta_kernel(const __nv_bfloat16* __restrict__ o,
                                                         const __nv_bfloat16* __restrict__ d_o, int BT, int T, int H,
                                                         float* __restrict__ delta) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= BT) return;
  const int b = row / T, t = row - b * T;
  const int D = H * kDh;
  const uint4* po = reinterpret_cast<const uint4*>(o + static_cast<long long>(row) * D);
  const uint4* pd = reinterpret_cast<const uint4*>(d_o + static_cast<long long>(row) * D);
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
    if (valid && (lane & 7) == 0) delta[(static_cast<long long>(b) * H + (c >> 3)) * T + t] = s;
  }
}

constexpr int kBwdThreads = 256;
constexpr int kBwdSmemK = 0;
constexpr int kBwdSmemV = kBwdSmemK + kTileBytes;
constexpr int kBwdSmemQ = kBwdSmemV + kTileBytes;        // 2 buffers
constexpr int kBwdSmemDO = kBwdSmemQ + 2 * kTileBytes;   // 2 buffers
constexpr int kBwdSmemPt = kBwdSmemDO + 2 * kTileBytes;  // Pᵀ  [128 keys × 128 q]
constexpr int kBwdSmemDSt = kBwdSmemPt + 2 * kTileBytes; // dSᵀ [128 keys × 128 q]
constexpr int kBwdSmemStat = kBwdSmemDSt + 2 * kTileBytes;  // lse2[2][128], delta[2][128]
constexpr int kBwdSmemBar = kBwdSmemStat + 4 * kTile * 4;
constexpr int kBwdSmemBytes = kBwdSmemBar + 128 + 1024;
constexpr int kBwdTmemCols = 512;  // Sᵀ [0,128) dPᵀ [128,256) dV [256,320) dK [320,384) dQ [384,448)

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_do,
                const float* __restrict__ lse, const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv,
                float* __restrict__ dq_acc, int T, int H, float scale, float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem + kBwdSmemK;
  uint8_t* sV = smem + kBwdSmemV;
  uint8_t* sQ = smem + kBwdSmemQ;
  uint8_t* sDO = smem + kBwdSmemDO;
  uint8_t* sPt = smem + kBwdSmemPt;
  uint8_t* sDSt = smem + kBwdSmemDSt;
  float* s_lse = reinterpret_cast<float*>(smem + kBwdSmemStat);  // [2][128]
  float* s_delta = s_lse + 2 * kTile;                            // [2][128]
  uint64_t* bar_kv = reinterpret_cast<uint64_t*>(smem + kBwdSmemBar);
  uint64_t* bar_qd = bar_kv + 1;  // [2]
  uint64_t* bar_s = bar_kv + 3;
  uint64_t* bar_g = bar_kv + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_kv + 5);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, col_half = warp >> 2;
  const int kb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int nq = (T + kTile - 1) / kTile;
  const int colq = h * kDh, colk = (H + h) * kDh, colv = (2 * H + h) * kDh;
  const int key0 = kb * kTile;

  if (tid == 0) {
    tma_prefetch_desc(&tma_qkv);
    tma_prefetch_desc(&tma_do);
    for (int i = 0; i < 5; ++i) mbar_init(bar_kv + i, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, kBwdTmemCols);

  auto load_stats = [&](int i, int buf) {
    if (tid < kTile) {
      const int q = i * kTile + tid;
      const long long idx = (static_cast<long long>(b) * H + h) * T + q;
      s_lse[buf * kTile + tid] = q < T ? __ldg(lse + idx) * kLog2e : INFINITY;
      s_delta[buf * kTile + tid] = q < T ? __ldg(delta + idx) : 0.f;
    }
  };
  load_stats(0, 0);

  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_st = tmem_base, tm_dpt = tmem_base + 128, tm_dv = tmem_base + 256, tm_dk = tmem_base + 320,
                 tm_dq = tmem_base + 384;

  constexpr uint32_t idesc_kk = umma_idesc_bf16(kTile, kTile, 0, 0);  // Sᵀ = K·Qᵀ, dPᵀ = V·dOᵀ
  constexpr uint32_t idesc_km = umma_idesc_bf16(kTile, kDh, 0, 1);    // dV += Pᵀ·dO, dK += dSᵀ·Q
  constexpr uint32_t idesc_mm = umma_idesc_bf16(kTile, kDh, 1, 1);    // dQ = dS·K

  auto issue_scores = [&](int buf) {  // thread 0 only
    const uint32_t ak = smem_u32(sK), av = smem_u32(sV);
    const uint32_t bq = smem_u32(sQ + buf * kTileBytes), bd = smem_u32(sDO + buf * kTileBytes);
#pragma unroll
    for (int k = 0; k < kDh / 16; ++k)
      tc_mma_bf16(tm_st, umma_smem_desc(ak + k * 32, 0, 1024), umma_smem_desc(bq + k * 32, 0, 1024), idesc_kk, k > 0);
#pragma unroll
    for (int k = 0; k < kDh / 16; ++k)
      tc_mma_bf16(tm_dpt, umma_smem_desc(av + k * 32, 0, 1024), umma_smem_desc(bd + k * 32, 0, 1024), idesc_kk, k > 0);
    tc_commit(bar_s);
  };

  if (tid == 0) {
    mbar_arrive_expect_tx(bar_kv, 2 * kTileBytes);
    tma_load_3d(sK, &tma_qkv, bar_kv, colk, key0, b);
    tma_load_3d(sV, &tma_qkv, bar_kv, colv, key0, b);
    mbar_arrive_expect_tx(&bar_qd[0], 2 * kTileBytes);
    tma_load_3d(sQ, &tma_qkv, &bar_qd[0], colq, 0, b);
    tma_load_3d(sDO, &tma_do, &bar_qd[0], colq, 0, b);
    mbar_wait(bar_kv, 0);
    mbar_wait(&bar_qd[0], 0);
    tc_fence_after_sync();
    issue_scores(0);
  }

  const int key_row = quad * 32 + lane;  // TMEM lane = key (Sᵀ, dPᵀ, dV, dK) or query (dQ)
  const bool key_valid = key0 + key_row < T;
  const uint32_t lane_field = static_cast<uint32_t>(quad * 32) << 16;

  for (int i = 0; i < nq; ++i) {
    const int buf = i & 1;
    if (tid == 0 && i + 1 < nq) {
      mbar_arrive_expect_tx(&bar_qd[buf ^ 1], 2 * kTileBytes);
      tma_load_3d(sQ + (buf ^ 1) * kTileBytes, &tma_qkv, &bar_qd[buf ^ 1], colq, (i + 1) * kTile, b);
      tma_load_3d(sDO + (buf ^ 1) * kTileBytes, &tma_do, &bar_qd[buf ^ 1], colq, (i + 1) * kTile, b);
    }
    if (i + 1 < nq) load_stats(i + 1, buf ^ 1);  // visible after the __syncthreads below

    __syncwarp();
    mbar_wait(bar_s, i & 1);
    tc_fence_after_sync();
    const float* lse2 = s_lse + buf * kTile;
    const float* dlt = s_delta + buf * kTile;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      const int col0 = col_half * 64 + c * 32;
      uint32_t rs[32], rp[32];
      tmem_ld_32x32(tm_st + lane_field + col0, rs);
      tmem_ld_32x32(tm_dpt + lane_field + col0, rp);
      tmem_ld_wait();
      float p[32], ds[32];
#pragma unroll
      for (int x = 0; x < 32; ++x) {
        const float e = fast_exp2(fmaf(__uint_as_float(rs[x]), scale_log2, -lse2[col0 + x]));
        p[x] = key_valid ? e : 0.f;
        ds[x] = p[x] * (__uint_as_float(rp[x]) - dlt[col0 + x]);
      }
      store_row32_sw128(sPt, key_row, col0, p);
      store_row32_sw128(sDSt, key_row, col0, ds);
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();

    if (tid == 0) {
      tc_fence_after_sync();
      const uint32_t apt = smem_u32(sPt), adst = smem_u32(sDSt);
      const uint32_t bdo = smem_u32(sDO + buf * kTileBytes), bq = smem_u32(sQ + buf * kTileBytes), bk = smem_u32(sK);
#pragma unroll
      for (int k = 0; k < kTile / 16; ++k)   // dV[key,d] += Σ_q Pᵀ[key,q]·dO[q,d]
        tc_mma_bf16(tm_dv, umma_smem_desc(apt + (k >> 2) * kTileBytes + (k & 3) * 32, 0, 1024),
                    umma_smem_desc(bdo + k * 2048, kTileBytes, 1024), idesc_km, (i > 0 || k > 0) ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < kTile / 16; ++k)   // dK[key,d] += Σ_q dSᵀ[key,q]·Q[q,d]
        tc_mma_bf16(tm_dk, umma_smem_desc(adst + (k >> 2) * kTileBytes + (k & 3) * 32, 0, 1024),
                    umma_smem_desc(bq + k * 2048, kTileBytes, 1024), idesc_km, (i > 0 || k > 0) ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < kTile / 16; ++k)   // dQ[q,d] = Σ_key dS[q,key]·K[key,d]; A = dSᵀ tile read MN-major
        tc_mma_bf16(tm_dq, umma_smem_desc(adst + k * 2048, kTileBytes, 1024),
                    umma_smem_desc(bk + k * 2048, kTileBytes, 1024), idesc_mm, k > 0);
      tc_commit(bar_g);
      if (i + 1 < nq) {
        mbar_wait(&bar_qd[buf ^ 1], ((i + 1) >> 1) & 1);
        tc_fence_after_sync();
        issue_scores(buf ^ 1);
      }
    }
    __syncwarp();
    mbar_wait(bar_g, i & 1);
    tc_fence_after_sync();
    {
      uint32_t r[32];
      tmem_ld_32x32(tm_dq + lane_field + col_half * 32, r);
      tmem_ld_wait();
      const int q = i * kTile + key_row;
      if (q < T) {
        float* dst = dq_acc + ((static_cast<long long>(b) * T + q) * H + h) * kDh + col_half * 32;
#pragma unroll
        for (int x = 0; x < 8; ++x)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * x),
                       "f"(__uint_as_float(r[4 * x + 0])), "f"(__uint_as_float(r[4 * x + 1])),
                       "f"(__uint_as_float(r[4 * x + 2])), "f"(__uint_as_float(r[4 * x + 3]))
                       : "memory");
      }
    }
    tc_fence_before_sync();
  }

  // dV (col_half 0) and dK·scale (col_half 1) → dqkv[b, key, 2|1, h, :]
  {
    const uint32_t src = (col_half == 0 ? tm_dv : tm_dk) + lane_field;
    const float mul = col_half == 0 ? 1.0f : scale;
    const int which = col_half == 0 ? 2 : 1;
    const int key = key0 + key_row;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(src + c * 32, r);
      tmem_ld_wait();
      if (key < T) {
        uint4* dst = reinterpret_cast<uint4*>(dqkv + ((static_cast<long long>(b) * T + key) * 3 + which) * H * kDh +
                                              h * kDh + c * 32);
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
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, kBwdTmemCols);
  }
}

// dq_acc fp32 [B,T,H,64] · scale → dqkv[b,t,0,h,:] bf16
__global__ void __launch_bounds__(256) attn_dq_store_kernel(const float* __restrict__ dq_acc, int BT, int H, float scale,
                                                            __nv_bfloat16* __restrict__ dqkv) {
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
  CUtensorMap tm, tm_kv;
  if (int rc = qkv_tensor_map(&tm, qkv, B, T, 3 * H * kDh)) return rc;
  if (int rc = qkv_tensor_map(&tm_kv, qkv, B, T, 3 * H * kDh, kSub)) return rc;
  static std::atomic<int> attr_done{0};
  if (!attr_done.load()) {
    VITK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmemBytes));
    attr_done.store(1);
  }
  const dim3 grid(static_cast<unsigned>((T + kTile - 1) / kTile), static_cast<unsigned>(H), static_cast<unsigned>(B));
  attn_fwd_kernel<<<grid, kFwdThreads, kFwdSmemBytes, static_cast<cudaStream_t>(stream)>>>(
      tm, tm_kv, static_cast<__nv_bfloat16*>(o), lse, static_cast<int>(T), static_cast<int>(H), scale * kLog2e);
  VITK_LAUNCH_CHECK("attn_fwd_kernel");
  return 0;
}

extern "C" VITK_API size_t vitk_attn_bwd_workspace_bytes(int64_t B, int64_t T, int64_t H) {
  const size_t dq = static_cast<size_t>(B) * T * H * kDh * sizeof(float);
  const size_t dl = (static_cast<size_t>(B) * H * T * sizeof(float) + 255) / 256 * 256;
  return dq + dl;
}

extern "C" VITK_API int vitk_attn_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, int64_t B, int64_t T,
                             int64_t H, float scale, void* dqkv, void* workspace, vitk_stream_t stream) {
  VITK_REQUIRE(qkv && o && d_o && lse && dqkv && workspace, VITK_EINVAL, "attn_bwd: NULL argument");
  if (int rc = check_shape("attn_bwd", B, T, H)) return rc;
  VITK_REQUIRE(aligned16(qkv) && aligned16(o) && aligned16(d_o) && aligned16(dqkv) && aligned16(workspace), VITK_EALIGN,
               "attn_bwd: buffers must be 16-byte aligned");
  VITK_REQUIRE(scale > 0.f, VITK_EINVAL, "attn_bwd: scale must be positive");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t dq_bytes = static_cast<size_t>(B) * T * H * kDh * sizeof(float);
  float* dq_acc = static_cast<float*>(workspace);
  float* delta = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + dq_bytes);
  CUtensorMap tm_qkv, tm_do;
  if (int rc = qkv_tensor_map(&tm_qkv, qkv, B, T, 3 * H * kDh)) return rc;
  if (int rc = qkv_tensor_map(&tm_do, d_o, B, T, H * kDh)) return rc;
  static std::atomic<int> attr_done{0};
  if (!attr_done.load()) {
    VITK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmemBytes));
    attr_done.store(1);
  }
  VITK_CUDA(cudaMemsetAsync(dq_acc, 0, dq_bytes, s));
  const int BT = static_cast<int>(B * T);
  attn_delta_kernel<<<(BT + 7) / 8, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(o),
                                                 static_cast<const __nv_bfloat16*>(d_o), BT, (int)T, (int)H, delta);
  VITK_LAUNCH_CHECK("attn_delta_kernel");
  const dim3 grid(static_cast<unsigned>((T + kTile - 1) / kTile), static_cast<unsigned>(H), static_cast<unsigned>(B));
  attn_bwd_kernel<<<grid, kBwdThreads, kBwdSmemBytes, s>>>(tm_qkv, tm_do, lse, delta, static_cast<__nv_bfloat16*>(dqkv),
                                                           dq_acc, (int)T, (int)H, scale, scale * kLog2e);
  VITK_LAUNCH_CHECK("attn_bwd_kernel");
  const long long n8 = static_cast<long long>(BT) * H * kDh / 8;
  attn_dq_store_kernel<<<static_cast<unsigned>((n8 + 255) / 256), 256, 0, s>>>(dq_acc, BT, (int)H, scale,
                                                                                static_cast<__nv_bfloat16*>(dqkv));
  VITK_LAUNCH_CHECK("attn_dq_store_kernel");
  return 0;
}
