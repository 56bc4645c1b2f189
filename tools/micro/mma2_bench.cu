// Issue-to-retire cost of cta_group::2 tcgen05.mma (M = 256 over a CTA pair) on B200: the leader's lane 0 issues
// back-to-back MMAs on zero-filled shared memory of both CTAs, clock64 around issue + commit + wait.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I chest-x-ray-vit_b200/csrc -o /tmp/mma2_bench tools/micro/mma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "sm100_prims.cuh"
using namespace vitk;

template <int N, int A_MN, int B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) bench(long long* out, int reps, int kblocks_per_commit) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc_pair(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tm = slot;
  const bool leader = (blockIdx.x & 1) == 0;
  if (leader && threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(256, N, A_MN, B_MN);
    const uint64_t ad = umma_smem_desc(smem_u32(smem), A_MN ? 8192 : 0, 1024);
    const uint64_t bd = umma_smem_desc(smem_u32(smem + 32768), B_MN ? 8192 : 0, 1024);
    uint32_t ph = 0;
    for (int warm = 0; warm < 2; ++warm) {
      const long long t0 = clock64();
      for (int i = 0; i < reps; ++i) {
        const int k = i & 3;
        tc_mma_bf16_pair(tm, ad + (A_MN ? 128 * k : 2 * k), bd + (B_MN ? 128 * k : 2 * k), idesc, 1);
        if (kblocks_per_commit && k == 3 && i + 1 < reps) tc_commit_pair(&bar, 1), ph ^= 1;   // a commit per K block, like the GEMM
      }
      const long long t1 = clock64();
      tc_commit_pair(&bar, 1);
      mbar_wait(&bar, ph);
      ph ^= 1;
      const long long t2 = clock64();
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x < 32) tmem_dealloc_pair(tm, 512);
}

template <int N, int A_MN, int B_MN>
void run(const char* name, int commits) {
  long long* d;
  cudaMalloc(&d, 16);
  auto k = bench<N, A_MN, B_MN>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  const int reps = 256;
  k<<<2, 128, 96 * 1024>>>(d, reps, commits);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-30s N=%3d %s issue %6.1f cyc/MMA   issue+retire %6.1f cyc/MMA   (nominal %d)  %s\n", name, N,
         commits ? "commit/4" : "        ", h[0] / double(reps), h[1] / double(reps), 128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<64, 0, 0>("pair SS  A K  B K", 0);
  run<128, 0, 0>("pair SS  A K  B K", 0);
  run<192, 0, 0>("pair SS  A K  B K", 0);
  run<256, 0, 0>("pair SS  A K  B K", 0);
  run<256, 0, 0>("pair SS  A K  B K", 1);
  run<128, 0, 1>("pair SS  A K  B MN", 0);
  run<256, 0, 1>("pair SS  A K  B MN", 0);
  run<128, 1, 1>("pair SS  A MN B MN", 0);
  run<256, 1, 1>("pair SS  A MN B MN", 0);
  run<256, 1, 1>("pair SS  A MN B MN", 1);
  return 0;
}
