// L2 -> shared-memory operand delivery rate of TMA on B200, with and without cluster multicast (no MMAs).
// Every CTA fills a ring of 32 KB stages (one 128x64 bf16 "A" box + one 128x64 "B" box, SW128) from L2-resident
// matrices.  GA / GB = how many CTAs of the cluster share the A / B box: each loads 1/G of it and multicasts.
// Build on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I chest-x-ray-vit_b200/csrc -o /tmp/tma_mc_bench \
//        tools/micro/tma_mc_bench.cu chest-x-ray-vit_b200/csrc/tmap.cu chest-x-ray-vit_b200/csrc/common.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cuda.h>
#include "sm100_prims.cuh"
#include "tmap.cuh"
using namespace vitk;

constexpr int kStages = 6;
constexpr int kStageBytes = 32768;

__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}

template <int CL, int GA, int GB>
__global__ void __launch_bounds__(64, 1) deliver(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb,
                                                 int iters, int a_tiles, int b_tiles, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[kStages], empty[kStages];
  const uint32_t rank = CL > 1 ? cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / CL;
  // A shared by ranks that differ in bit 1 (GA=2); B shared by ranks that differ in bit 2 (GB=2) or bit 1 when GA=1
  constexpr int kABit = 2, kBBit = (GA == 2) ? 4 : 2;
  const uint32_t a_peer = GA == 2 ? (rank ^ kABit) : rank, b_peer = GB == 2 ? (rank ^ kBBit) : rank;
  constexpr int kArrivals = 1 + (GA == 2) + (GB == 2);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kArrivals); }
    fence_mbar_init();
  }
  __syncthreads();
  if (CL > 1) cluster_sync_all();
  const long long t0 = clock64();
  if (threadIdx.x == 0) {                          // producer
    const int a_piece = GA == 2 ? ((rank & kABit) != 0) : 0, b_piece = GB == 2 ? ((rank & kBBit) != 0) : 0;
    const uint16_t a_mask = uint16_t((1u << rank) | (1u << a_peer)), b_mask = uint16_t((1u << rank) | (1u << b_peer));
    // the ranks that share A must ask for the same A tile: derive it from the cluster id and the non-shared bits
    const int a_sel = cluster_id * 8 + (GA == 2 ? (rank & ~kABit) : rank);
    const int b_sel = cluster_id * 8 + (GB == 2 ? (rank & ~kBBit) : rank);
    for (int it = 0; it < iters; ++it) {
      const int s = it % kStages;
      const uint32_t ph = (it / kStages) & 1;
      mbar_wait(&empty[s], ph ^ 1);
      mbar_arrive_expect_tx(&full[s], kStageBytes);
      const int k0 = (it % 12) * 64, tile = it / 12;
      const int ra = ((a_sel * 5 + tile) % a_tiles) * 128, rb = ((b_sel * 3 + tile) % b_tiles) * 128;
      uint8_t* st = smem + s * kStageBytes;
      if (GA == 2) tma_load_2d_mc(st + a_piece * 8192, &ta, &full[s], k0, ra + a_piece * 64, a_mask);
      else tma_load_2d(st, &ta, &full[s], k0, ra);
      if (GB == 2) tma_load_2d_mc(st + 16384 + b_piece * 8192, &tb, &full[s], k0, rb + b_piece * 64, b_mask);
      else tma_load_2d(st + 16384, &tb, &full[s], k0, rb);
    }
  } else if (threadIdx.x == 32) {                  // consumer: releases the stage here and in the CTAs that write into it
    for (int it = 0; it < iters; ++it) {
      const int s = it % kStages;
      const uint32_t ph = (it / kStages) & 1;
      mbar_wait(&full[s], ph);
      mbar_arrive(&empty[s]);
      if (GA == 2) mbar_arrive_cluster(mapa_shared(smem_u32(&empty[s]), a_peer));
      if (GB == 2) mbar_arrive_cluster(mapa_shared(smem_u32(&empty[s]), b_peer));
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (CL > 1) cluster_sync_all();                  // nobody exits while a peer may still multicast into / arrive on it
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

static __nv_bfloat16* g_a;
static __nv_bfloat16* g_b;
constexpr int kRowsA = 9216, kRowsB = 3072, kK = 768;

template <int CL, int GA, int GB>
void run(const char* name, int iters) {
  CUtensorMap ta, tb;
  const uint64_t dims_a[2] = {kK, kRowsA}, dims_b[2] = {kK, kRowsB}, strides[1] = {kK * 2};
  const uint32_t box_a[2] = {64, 128 / GA}, box_b[2] = {64, 128 / GB};
  if (make_tensor_map_bf16(&ta, g_a, 2, dims_a, strides, box_a) || make_tensor_map_bf16(&tb, g_b, 2, dims_b, strides, box_b)) {
    printf("%s: tensor map failed\n", name);
    return;
  }
  auto kern = deliver<CL, GA, GB>;
  const int smem = kStages * kStageBytes + 1024;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (CL > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem;
  cfg.gridDim = dim3(CL);
  int max_clusters = 0;
  cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
  if (max_clusters <= 0) { printf("%s: no clusters fit (%s)\n", name, cudaGetErrorString(cudaGetLastError())); return; }
  const int ctas = max_clusters * CL;
  cfg.gridDim = dim3(ctas);
  long long* d_cycles;
  cudaMalloc(&d_cycles, ctas * sizeof(long long));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    cudaLaunchKernelEx(&cfg, kern, ta, tb, iters, kRowsA / 128, kRowsB / 128, d_cycles);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  std::vector<long long> h(ctas);
  cudaMemcpy(h.data(), d_cycles, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
  double sum = 0; long long mx = 0;
  for (long long c : h) { sum += c; if (c > mx) mx = c; }
  const double avg = sum / ctas;
  const double l2_bytes_per_cta = double(iters) * (16384.0 / GA + 16384.0 / GB);
  printf("%-44s ctas %3d  cycles/stage avg %6.1f max %6.1f  recv %5.1f B/clk/SM  L2 read %5.1f B/clk/SM  chip recv %5.2f TB/s (%.3f ms)\n",
         name, ctas, avg / iters, double(mx) / iters, kStageBytes * double(iters) / avg, l2_bytes_per_cta / avg,
         double(ctas) * iters * kStageBytes / (best * 1e-3) / 1e12, best);
  cudaFree(d_cycles);
}

int main() {
  cudaMalloc(&g_a, size_t(kRowsA) * kK * 2);
  cudaMalloc(&g_b, size_t(kRowsB) * kK * 2);
  cudaMemset(g_a, 0, size_t(kRowsA) * kK * 2);
  cudaMemset(g_b, 0, size_t(kRowsB) * kK * 2);
  const int iters = 12 * 200;
  run<1, 1, 1>("no cluster, unicast", iters);
  run<2, 1, 1>("cluster 2, unicast", iters);
  run<4, 1, 1>("cluster 4, unicast", iters);
  run<4, 2, 1>("cluster 4, A shared by 2 (multicast)", iters);
  run<4, 1, 2>("cluster 4, B shared by 2 (multicast)", iters);
  run<8, 1, 1>("cluster 8, unicast", iters);
  run<8, 2, 2>("cluster 8, A and B shared by 2 (multicast)", iters);
  return 0;
}
