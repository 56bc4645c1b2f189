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
// Two elements at a time with packed fp32 arithmetic (fma.rn.f32x2 & co., sm_100): the same operations in the same
// order as gelu_parts — bit-identical results — in 9 packed + 6 scalar instructions per pair instead of 2 × 15.
// The packed forms have the FLOP rate of scalar FFMA; what they save is issue slots, which is what the GELU
// epilogue (2 MUFU + 13 FP32 ops per element next to TMEM loads, packing and stores) runs out of.
struct GeluParts2 {
  float2 cdf, pdf;
};
__device__ __forceinline__ GeluParts2 gelu_parts2(float2 u) {
  const float2 au = make_float2(fabsf(u.x), fabsf(u.y));
  const float2 den = __ffma2_rn(make_float2(0.3275911f * 0.7071067811865476f, 0.3275911f * 0.7071067811865476f), au,
                                make_float2(1.0f, 1.0f));
  const float2 t = make_float2(rcp_approx(den.x), rcp_approx(den.y));
  float2 poly = __ffma2_rn(make_float2(1.061405429f, 1.061405429f), t, make_float2(-1.453152027f, -1.453152027f));
  poly = __ffma2_rn(poly, t, make_float2(1.421413741f, 1.421413741f));
  poly = __ffma2_rn(poly, t, make_float2(-0.284496736f, -0.284496736f));
  poly = __ffma2_rn(poly, t, make_float2(0.254829592f, 0.254829592f));
  poly = __fmul2_rn(poly, t);
  const float2 uu = __fmul2_rn(u, u);
  const float2 arg = __fmul2_rn(uu, make_float2(-0.5f * 1.4426950408889634f, -0.5f * 1.4426950408889634f));
  const float2 e = make_float2(ex2_approx(arg.x), ex2_approx(arg.y));
  const float2 erf_abs = __ffma2_rn(make_float2(-poly.x, -poly.y), e, make_float2(1.0f, 1.0f));
  GeluParts2 r;
  r.cdf = __ffma2_rn(make_float2(0.5f, 0.5f), make_float2(copysignf(erf_abs.x, u.x), copysignf(erf_abs.y, u.y)),
                     make_float2(0.5f, 0.5f));
  r.pdf = __fmul2_rn(make_float2(0.3989422804014327f, 0.3989422804014327f), e);
  return r;
}
__device__ __forceinline__ float gelu_erf(float u) { return u * gelu_parts(u).cdf; }
__device__ __forceinline__ float gelu_erf_grad(float u) {
  const GeluParts g = gelu_parts(u);
  return fmaf(u, g.pdf, g.cdf);
}

}  // namespace vitk
