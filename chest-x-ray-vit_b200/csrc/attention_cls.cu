// Attention for ONE query row per (image, head): the top encoder layer of a classifier.
//
// HF's ViTForImageClassification feeds only sequence_output[:, 0] — the CLS row of the last layer — to the classifier
// (modeling_vit.py:641), so of the last layer's attention only the CLS query's output is ever read, and in backward only
// that row carries a gradient.  K and V of all tokens are still needed.  For that layer the dense flash kernels
// (attention.cu: 44 µs forward, 102 + 18 µs backward at B = 16) are replaced by two small kernels, one CTA per
// (head, image): scores of 1 × T, a softmax, and rank-one outer products — a few hundred KFLOP per CTA, latency-bound.
// Same layouts as attention.cu: qkv bf16 [B,T,3,H,64], o / do bf16 [B,T,H·64] (only row t = 0 touched), lse fp32 [B,H,T]
// (only t = 0), dqkv bf16 [B,T,3,H,64] (dK, dV dense; dQ zero except row 0).  Result identical to the dense kernels
// restricted to query 0 up to fp32 summation order (tests/test_gpu_attention.py::test_cls_row_attention_*).
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"

namespace vitk {

constexpr int kClsThreads = 256;
constexpr int kClsDh = 64;

__device__ __forceinline__ float cls_block_reduce(float v, float* red, bool is_max) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float w = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, w) : v + w;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int w = 1; w < kClsThreads / 32; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  return r;
}

__device__ __forceinline__ float dot64(const float* __restrict__ a, const __nv_bfloat16* __restrict__ row) {
  const uint4* p = reinterpret_cast<const uint4*>(row);
  float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 q = __ldg(p + c);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
      acc0 = fmaf(a[8 * c + 2 * i], f.x, acc0);
      acc1 = fmaf(a[8 * c + 2 * i + 1], f.y, acc1);
    }
  }
  return acc0 + acc1;
}

// o[b,0,h,:] = softmax(scale · q_cls · Kᵀ) · V ; lse[b,h,0]
__global__ void __launch_bounds__(kClsThreads)
attn_cls_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, int T, int H, float scale, __nv_bfloat16* __restrict__ o,
                    float* __restrict__ lse) {
  extern __shared__ float sm[];          // [T] scores / probabilities, then [8][64] partial outputs
  __shared__ float q[kClsDh], red[kClsThreads / 32];
  pdl_wait();
  pdl_launch_dependents();
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long row_elems = 3LL * H * kClsDh;
  const __nv_bfloat16* base = qkv + static_cast<long long>(b) * T * row_elems;
  if (tid < kClsDh) q[tid] = __bfloat162float(base[h * kClsDh + tid]) * scale;
  __syncthreads();
  float mx = -INFINITY;
  for (int j = tid; j < T; j += kClsThreads) {
    const float s = dot64(q, base + j * row_elems + (H + h) * kClsDh);
    sm[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = cls_block_reduce(mx, red, true);
  float sum = 0.f;
  for (int j = tid; j < T; j += kClsThreads) {
    const float p = __expf(sm[j] - mx);
    sm[j] = p;
    sum += p;
  }
  sum = cls_block_reduce(sum, red, false);     // also orders the sm[] writes before the reads below
  // o = Σ_j p_j V_j: each warp takes keys warp, warp+8, …; a lane owns two adjacent head dims
  float a0 = 0.f, a1 = 0.f;
  for (int j = warp; j < T; j += kClsThreads / 32) {
    const float p = sm[j];
    const __nv_bfloat162 v = reinterpret_cast<const __nv_bfloat162*>(base + j * row_elems + (2 * H + h) * kClsDh)[lane];
    const float2 f = __bfloat1622float2(v);
    a0 = fmaf(p, f.x, a0);
    a1 = fmaf(p, f.y, a1);
  }
  float* part = sm + ((T + 3) & ~3);
  part[warp * kClsDh + 2 * lane] = a0;
  part[warp * kClsDh + 2 * lane + 1] = a1;
  __syncthreads();
  if (tid < kClsDh) {
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < kClsThreads / 32; ++w) acc += part[w * kClsDh + tid];
    o[(static_cast<long long>(b) * T * H + h) * kClsDh + tid] = __float2bfloat16(acc / sum);
  }
  if (tid == 0) lse[(static_cast<long long>(b) * H + h) * T] = mx + logf(sum);
}

// dqkv from the gradient of the CLS row only.  Phase 1: a thread per key computes the two 64-long dot products
// (s_j = q·K_j, dP_j = dO·V_j) and leaves p_j, dS_j in shared memory — no warp reductions on the per-key path.  Phase 2:
// a warp per key, lanes over the head dims, writes the rank-one rows dV_j = p_j·dO and dK_j = dS_j·q coalesced and
// accumulates dQ_cls = Σ_j dS_j·K_j in registers.
__global__ void __launch_bounds__(kClsThreads)
attn_cls_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ o,
                    const __nv_bfloat16* __restrict__ d_o, const float* __restrict__ lse, int T, int H, float scale,
                    __nv_bfloat16* __restrict__ dqkv) {
  extern __shared__ float sm[];               // [T] p_j, [T] dS_j
  __shared__ float q[kClsDh], g[kClsDh], red[kClsThreads / 32], dq_part[kClsThreads / 32][kClsDh];
  pdl_wait();
  pdl_launch_dependents();
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long row_elems = 3LL * H * kClsDh;
  const __nv_bfloat16* base = qkv + static_cast<long long>(b) * T * row_elems;
  __nv_bfloat16* dbase = dqkv + static_cast<long long>(b) * T * row_elems;
  const long long orow = (static_cast<long long>(b) * T * H + h) * kClsDh;      // CLS row of o / do for this head
  float* sp = sm;
  float* sds = sm + ((T + 3) & ~3);
  float dlt = 0.f;
  if (tid < kClsDh) {
    q[tid] = __bfloat162float(base[h * kClsDh + tid]) * scale;                  // scale folded into q
    g[tid] = __bfloat162float(d_o[orow + tid]);
    dlt = g[tid] * __bfloat162float(o[orow + tid]);
  }
  const float delta = cls_block_reduce(dlt, red, false);                         // Δ = dO·O (also publishes q, g)
  const float l = lse[(static_cast<long long>(b) * H + h) * T];
  for (int j = tid; j < T; j += kClsThreads) {
    const float s = dot64(q, base + j * row_elems + (H + h) * kClsDh);
    const float dp = dot64(g, base + j * row_elems + (2 * H + h) * kClsDh);
    const float p = __expf(s - l);
    sp[j] = p;
    sds[j] = p * (dp - delta);
  }
  __syncthreads();
  const float q0 = q[2 * lane], q1 = q[2 * lane + 1], g0 = g[2 * lane], g1 = g[2 * lane + 1];
  float dq0 = 0.f, dq1 = 0.f;                 // this lane's two head dims of dQ_cls, summed over the warp's keys
  for (int j = warp; j < T; j += kClsThreads / 32) {
    const float p = sp[j], ds = sds[j];
    const float2 kf = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(base + j * row_elems + (H + h) * kClsDh)[lane]);
    __nv_bfloat16* drow = dbase + j * row_elems + h * kClsDh;
    reinterpret_cast<__nv_bfloat162*>(drow + 2 * H * kClsDh)[lane] = __floats2bfloat162_rn(p * g0, p * g1);     // dV_j
    reinterpret_cast<__nv_bfloat162*>(drow + H * kClsDh)[lane] = __floats2bfloat162_rn(ds * q0, ds * q1);       // dK_j
    if (j > 0) reinterpret_cast<__nv_bfloat162*>(drow)[lane] = __floats2bfloat162_rn(0.f, 0.f);                 // dQ_j = 0
    dq0 = fmaf(ds * scale, kf.x, dq0);
    dq1 = fmaf(ds * scale, kf.y, dq1);
  }
  dq_part[warp][2 * lane] = dq0;
  dq_part[warp][2 * lane + 1] = dq1;
  __syncthreads();
  if (tid < kClsDh) {
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < kClsThreads / 32; ++w) acc += dq_part[w][tid];
    dbase[h * kClsDh + tid] = __float2bfloat16(acc);
  }
}

}  // namespace vitk

using namespace vitk;

static int cls_check(const char* who, int64_t B, int64_t T, int64_t H) {
  VITK_REQUIRE(B > 0 && T > 0 && H > 0 && B < 65536 && H < 65536 && T <= 4096, VITK_EINVAL,
               "%s: unsupported shape B=%lld T=%lld H=%lld (T <= 4096)", who, (long long)B, (long long)T, (long long)H);
  return 0;
}

extern "C" VITK_API int vitk_attn_cls_fwd(const void* qkv, int64_t B, int64_t T, int64_t H, float scale, void* o, float* lse,
                                          vitk_stream_t stream) {
  VITK_REQUIRE(qkv && o && lse, VITK_EINVAL, "attn_cls_fwd: NULL argument");
  if (int rc = cls_check("attn_cls_fwd", B, T, H)) return rc;
  VITK_REQUIRE(aligned16(qkv) && aligned16(o) && scale > 0.f, VITK_EALIGN, "attn_cls_fwd: buffers must be 16-byte aligned, scale > 0");
  const size_t smem = (((T + 3) & ~3) + (kClsThreads / 32) * kClsDh) * sizeof(float);
  VITK_CUDA(launch_pdl(attn_cls_fwd_kernel, dim3((unsigned)H, (unsigned)B), dim3(kClsThreads), smem, static_cast<cudaStream_t>(stream),
                       static_cast<const __nv_bfloat16*>(qkv), (int)T, (int)H, scale, static_cast<__nv_bfloat16*>(o), lse));
  VITK_LAUNCH_CHECK("attn_cls_fwd_kernel");
  return 0;
}

extern "C" VITK_API int vitk_attn_cls_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, int64_t B, int64_t T,
                                          int64_t H, float scale, void* dqkv, vitk_stream_t stream) {
  VITK_REQUIRE(qkv && o && d_o && lse && dqkv, VITK_EINVAL, "attn_cls_bwd: NULL argument");
  if (int rc = cls_check("attn_cls_bwd", B, T, H)) return rc;
  VITK_REQUIRE(aligned16(qkv) && aligned16(o) && aligned16(d_o) && aligned16(dqkv) && scale > 0.f, VITK_EALIGN,
               "attn_cls_bwd: buffers must be 16-byte aligned, scale > 0");
  const size_t smem = 2 * ((T + 3) & ~3) * sizeof(float);
  VITK_CUDA(launch_pdl(attn_cls_bwd_kernel, dim3((unsigned)H, (unsigned)B), dim3(kClsThreads), smem, static_cast<cudaStream_t>(stream),
                       static_cast<const __nv_bfloat16*>(qkv), static_cast<const __nv_bfloat16*>(o),
                       static_cast<const __nv_bfloat16*>(d_o), lse, (int)T, (int)H, scale, static_cast<__nv_bfloat16*>(dqkv)));
  VITK_LAUNCH_CHECK("attn_cls_bwd_kernel");
  return 0;
}
