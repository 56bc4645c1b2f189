// Element-wise math shared by the GEMM epilogues.
#pragma once
#include <cuda_runtime.h>

namespace vitk {

// erf via Abramowitz–Stegun 7.1.26 (|err| < 2e-7 in fp32 with MUFU ex2/rcp): the exact-erf GELU the
// reference uses (HF activations.py:85-86).  One rcp and one ex2 per element; the same
// e = exp(−u²/2) also gives the normal pdf needed by the derivative.
struct GeluParts {
  float cdf;  // Φ(u)
  float pdf;  // φ(u)
};
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ GeluParts gelu_parts(float u) {
  const float au = fabsf(u);
  const float t = rcp_approx(fmaf(0.3275911f * 0.7071067811865476f, au, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  const float e = ex2_approx(u * u * (-0.5f * 1.4426950408889634f));  // exp(−u²/2) = exp(−(u/√2)²)
  const float erf_abs = fmaf(-poly, e, 1.0f);
  GeluParts r;
  r.cdf = 0.5f + 0.5f * copysignf(erf_abs, u);
  r.pdf = 0.3989422804014327f * e;
  return r;
}
__device__ __forceinline__ float gelu_erf(float u) { return u * gelu_parts(u).cdf; }
__device__ __forceinline__ float gelu_erf_grad(float u) {
  const GeluParts g = gelu_parts(u);
  return fmaf(u, g.pdf, g.cdf);
}

}  // namespace vitk
