// Issue-to-retire cost of single-CTA tcgen05.mma shapes / operand sources on B200 (one CTA, lane 0 issues 256
// back-to-back MMAs on zero-filled shared memory, clock64 around issue+commit+wait).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I chest-x-ray-vit_b200/csrc -o gpurun_out/mma_bench tools/micro/mma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "sm100_prims.cuh"
using namespace vitk;

template <int N, int A_MN, int B_MN, int TS>
__global__ void __launch_bounds__(128, 1) bench(long long* out, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N, A_MN, B_MN);
    const uint64_t ad = umma_smem_desc(smem_u32(smem), A_MN ? 8192 : 0, 1024);
    const uint64_t bd = umma_smem_desc(smem_u32(smem + 32768), B_MN ? 8192 : 0, 1024);
    uint32_t ph = 0;
    for (int warm = 0; warm < 2; ++warm) {
      const long long t0 = clock64();
      for (int i = 0; i < reps; ++i) {
        const int k = i & 3;
        if (TS) tc_mma_bf16_ts(tm, tm + 256 + k * 8, bd + (B_MN ? 128 * k : 2 * k), idesc, 1);
        else tc_mma_bf16(tm, ad + (A_MN ? 128 * k : 2 * k), bd + (B_MN ? 128 * k : 2 * k), idesc, 1);
      }
      const long long t1 = clock64();
      tc_commit(&bar);
      mbar_wait(&bar, ph);
      ph ^= 1;
      const long long t2 = clock64();
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

template <int N, int A_MN, int B_MN, int TS>
void run(const char* name) {
  long long* d;
  cudaMalloc(&d, 16);
  auto k = bench<N, A_MN, B_MN, TS>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  const int reps = 256;
  k<<<1, 128, 96 * 1024>>>(d, reps);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-34s N=%3d  issue %6.1f cyc/MMA   issue+retire %6.1f cyc/MMA   (nominal %d)  %s\n", name, N, h[0] / double(reps),
         h[1] / double(reps), 128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<64, 0, 0, 0>("SS  A K-major  B K-major");
  run<128, 0, 0, 0>("SS  A K-major  B K-major");
  run<256, 0, 0, 0>("SS  A K-major  B K-major");
  run<64, 0, 1, 0>("SS  A K-major  B MN-major");
  run<128, 0, 1, 0>("SS  A K-major  B MN-major");
  run<64, 1, 1, 0>("SS  A MN-major B MN-major");
  run<128, 1, 1, 0>("SS  A MN-major B MN-major");
  run<256, 1, 1, 0>("SS  A MN-major B MN-major");
  run<64, 0, 1, 1>("TS  A TMEM     B MN-major");
  run<64, 0, 0, 1>("TS  A TMEM     B K-major");
  run<128, 0, 0, 1>("TS  A TMEM     B K-major");
  run<16, 0, 0, 0>("SS  A K-major  B K-major");
  return 0;
}
