// Multi-tensor-free optimizer step: parameters, gradients and both AdamW moments live in flat
// fp32 buffers, so one HBM-bound kernel updates all of them and refreshes the bf16 shadow the
// GEMMs read.  Follows torch.optim.AdamW (decoupled weight decay) as configured by HF Trainer
// (HF trainer.py:1143-1217, training_args.py:778-862); global-norm clipping (HF trainer.py:2489-2493)
// is a device-side scale factor so the step never synchronises with the host.
#include <cuda_bf16.h>

#include "common.cuh"

namespace vitk {

__global__ void __launch_bounds__(256)
adamw_kernel(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
             uint2* __restrict__ p16, long long n4, float lr, float b1, float b2, float eps, float wd, float bc1,
             float rsqrt_bc2, const float* __restrict__ grad_scale, int zero_grad, const float* __restrict__ bias_corr_dev) {
  const float gs = grad_scale ? __ldg(grad_scale) : 1.0f;
  if (bias_corr_dev) {              // CUDA-graph replay: the step counter lives on the device (adamw_tick_kernel)
    bc1 = __ldg(bias_corr_dev);
    rsqrt_bc2 = __ldg(bias_corr_dev + 1);
  }
  const float step = lr / bc1, decay = 1.0f - lr * wd;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 pp = p[i], mm = m[i], vv = v[i];
    const float4 gg = g[i];
    float* pa = &pp.x; float* ma = &mm.x; float* va = &vv.x;
    const float* ga = &gg.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gr = ga[k] * gs;
      ma[k] = fmaf(1.0f - b1, gr - ma[k], ma[k]);
      va[k] = fmaf(1.0f - b2, gr * gr - va[k], va[k]);
      const float denom = fmaf(sqrtf(va[k]), rsqrt_bc2, eps);
      pa[k] = fmaf(-step, __fdiv_rn(ma[k], denom), pa[k] * decay);
    }
    p[i] = pp; m[i] = mm; v[i] = vv;
    if (zero_grad) g[i] = make_float4(0.f, 0.f, 0.f, 0.f);   // next backward accumulates into a clean buffer: no memset pass
    if (p16) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(pp.x, pp.y), hi = __floats2bfloat162_rn(pp.z, pp.w);
      p16[i] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    }
  }
}

// Σ x² in a FIXED summation order: every block writes its partial sum to scratch[2 + block], takes a ticket, and the
// block that draws the last ticket adds the partials up in index order and accumulates the total into *out.  The result
// does not depend on which block finishes last, so data-parallel replicas that hold bit-identical gradients compute
// bit-identical clip coefficients and their parameters never drift apart (an atomicAdd per block would make the
// coefficient — hence every parameter — depend on the order of the atomics).  scratch[0] is the ticket counter
// (left at zero), scratch[1] is unused padding.
__global__ void __launch_bounds__(256) sumsq_kernel(const float4* __restrict__ x, long long n4, float* __restrict__ out,
                                                    float* __restrict__ scratch) {
  __shared__ float red[8];
  __shared__ bool is_last;
  float s = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 a = __ldg(x + i);
    s += (a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    scratch[2 + blockIdx.x] = t;
    __threadfence();
    const unsigned ticket = atomicInc(reinterpret_cast<unsigned*>(scratch), gridDim.x - 1);   // wraps to 0 after the last block
    is_last = ticket == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // fixed-order tree over the partials: thread t sums partials t, t+256, … in order, then a fixed shuffle/smem tree
  float t = 0.f;
  for (unsigned i = threadIdx.x; i < gridDim.x; i += 256) t += __ldcg(scratch + 2 + i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < 8; ++w) tot += red[w];
    *out += tot;
  }
}

// Owner's step of the peer-memory gradient all-reduce (parallel.PeerGradSync): this rank's shard of a bucket becomes the
// mean over all ranks — own[i] = (own[i] + Σ_p peer[p][i]) / world — where peer[p] are the copies of the same shard
// pulled from the other GPUs over NVLink by the copy engines.  Fixed summation order (own, then peers in rank order):
// every rank later copies these very bits, so replicas stay identical.  HBM-bound: (world + 1)·4 B per element.
__global__ void __launch_bounds__(256)
shard_mean_kernel(float4* __restrict__ own, const float4* __restrict__ peers, long long n4, long long stride4, int n_peers,
                  float inv_world) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 a = own[i];
    for (int p = 0; p < n_peers; ++p) {
      const float4 b = __ldg(peers + p * stride4 + i);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    own[i] = make_float4(a.x * inv_world, a.y * inv_world, a.z * inv_world, a.w * inv_world);
  }
}

// Device-side step counter for graph replay: t += inc; bc = {1 − β1^t, 1/sqrt(1 − β2^t)} (double pow, once per step).
__global__ void adamw_tick_kernel(long long* step, int inc, float b1, float b2, float* bc) {
  const long long t = *step + inc;
  if (inc) *step = t;
  bc[0] = static_cast<float>(1.0 - pow(static_cast<double>(b1), static_cast<double>(t)));
  bc[1] = static_cast<float>(1.0 / sqrt(1.0 - pow(static_cast<double>(b2), static_cast<double>(t))));
}

// torch.nn.utils.clip_grad_norm_: coef = min(1, max_norm / (‖g‖ + 1e-6))
__global__ void clip_scale_kernel(const float* sumsq, float max_norm, float* scale) {
  const float c = max_norm / (sqrtf(*sumsq) + 1e-6f);
  *scale = c < 1.0f ? c : 1.0f;
}

}  // namespace vitk

using namespace vitk;

extern "C" VITK_API int vitk_adamw(float* p, float* g, float* m, float* v, void* p_bf16, int64_t n, float lr,
                                   float beta1, float beta2, float eps, float weight_decay, float bias_corr1,
                                   float bias_corr2, const float* grad_scale, int zero_grad, const float* bias_corr_dev,
                                   vitk_stream_t stream) {
  VITK_REQUIRE(p && g && m && v && n > 0 && n % 4 == 0, VITK_EINVAL, "adamw: n must be a positive multiple of 4");
  if (bias_corr_dev) bias_corr1 = bias_corr2 = 1.0f;
  VITK_REQUIRE(aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v) && (reinterpret_cast<uintptr_t>(p_bf16) & 7) == 0,
               VITK_EALIGN, "adamw: buffers must be 16-byte aligned");
  VITK_REQUIRE(bias_corr1 > 0.f && bias_corr2 > 0.f, VITK_EINVAL, "adamw: bias corrections must be positive");
  const long long n4 = n / 4;
  long long blocks = (n4 + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  adamw_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<float4*>(p), reinterpret_cast<float4*>(g), reinterpret_cast<float4*>(m),
      reinterpret_cast<float4*>(v), static_cast<uint2*>(p_bf16), n4, lr, beta1, beta2, eps, weight_decay, bias_corr1,
      1.0f / sqrtf(bias_corr2), grad_scale, zero_grad, bias_corr_dev);
  VITK_LAUNCH_CHECK("adamw_kernel");
  return 0;
}

extern "C" VITK_API int vitk_adamw_tick(int64_t* step_dev, int increment, float beta1, float beta2, float* bias_corr_dev,
                                        vitk_stream_t stream) {
  VITK_REQUIRE(step_dev && bias_corr_dev && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f, VITK_EINVAL,
               "adamw_tick: bad argument");
  adamw_tick_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<long long*>(step_dev), increment ? 1 : 0,
                                                                   beta1, beta2, bias_corr_dev);
  VITK_LAUNCH_CHECK("adamw_tick_kernel");
  return 0;
}

extern "C" VITK_API int vitk_shard_mean(float* own, const float* peers, int64_t n, int64_t peer_stride, int n_peers,
                                        float inv_world, int max_ctas, vitk_stream_t stream) {
  VITK_REQUIRE(own && n > 0 && n % 4 == 0 && aligned16(own) && n_peers >= 0 && (n_peers == 0 || (peers && aligned16(peers))) &&
                   peer_stride % 4 == 0 && peer_stride >= (n_peers > 0 ? n : 0),
               VITK_EINVAL, "shard_mean: n and peer_stride must be multiples of 4, buffers 16-byte aligned");
  const long long n4 = n / 4;
  long long blocks = (n4 + 255) / 256;
  const long long cap = max_ctas > 0 ? max_ctas : static_cast<long long>(num_sms()) * 2;
  if (blocks > cap) blocks = cap;
  shard_mean_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<float4*>(own), reinterpret_cast<const float4*>(peers), n4, peer_stride / 4, n_peers, inv_world);
  VITK_LAUNCH_CHECK("shard_mean_kernel");
  return 0;
}

extern "C" VITK_API int64_t vitk_sumsq_scratch_floats(void) { return 2 + static_cast<int64_t>(num_sms()) * 8; }

extern "C" VITK_API int vitk_sumsq_f32(const float* x, int64_t n, float* out, float* scratch, vitk_stream_t stream) {
  VITK_REQUIRE(x && out && scratch && n > 0 && n % 4 == 0 && aligned16(x), VITK_EINVAL, "sumsq: n must be a positive multiple of 4");
  const long long n4 = n / 4;
  long long blocks = (n4 + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  sumsq_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(x), n4, out, scratch);
  VITK_LAUNCH_CHECK("sumsq_kernel");
  return 0;
}

extern "C" VITK_API int vitk_clip_scale(const float* sumsq, float max_norm, float* scale, vitk_stream_t stream) {
  VITK_REQUIRE(sumsq && scale && max_norm > 0.f, VITK_EINVAL, "clip_scale: bad argument");
  clip_scale_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(sumsq, max_norm, scale);
  VITK_LAUNCH_CHECK("clip_scale_kernel");
  return 0;
}
