// Per-SM throughput of FFMA, packed FFMA2 (fma.rn.f32x2), MUFU.EX2 and MUFU.RCP on B200: 8 warps per SM sub-partition
// run long independent chains; clock64 per block.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/alu_bench tools/micro/alu_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, long long* cyc, int iters) {
  float a[8], b = threadIdx.x * 1e-3f + 1.0f, c = 0.5f;
  float2 p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = i + threadIdx.x; p[i] = make_float2(a[i], a[i] + 1.f); }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = fmaf(a[i], b, c);
      if (MODE == 1) p[i] = __ffma2_rn(p[i], make_float2(b, b), make_float2(c, c));
      if (MODE == 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 3) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 4) { a[i] = fmaf(a[i], b, c); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(p[i].x)); }
      if (MODE == 5) { unsigned& hx = reinterpret_cast<unsigned&>(a[i]); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(hx)); }
      if (MODE == 6) { unsigned& hx = reinterpret_cast<unsigned&>(a[i]); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(hx)); }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + p[i].x + p[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_iter_elems) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 4096;
  k<MODE><<<148, 1024>>>(out, cyc, iters);
  k<MODE><<<148, 1024>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double ops = 1024.0 * iters * 8 * per_iter_elems;
  printf("%-22s %8.1f scalar ops/clk/SM  (%s)\n", name, ops / h[0], cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("FFMA", 1);
  run<1>("FFMA2 (f32x2)", 2);
  run<2>("MUFU.EX2", 1);
  run<3>("MUFU.RCP", 1);
  run<4>("FFMA + MUFU.EX2 mix", 2);
  run<5>("MUFU.EX2 f16x2", 2);
  run<6>("MUFU.EX2 bf16x2", 2);
  return 0;
}
