// CUtensorMap construction for all TMA users of libvitk, with a process-wide cache: the engine
// launches the same GEMMs on the same buffers every step, so the (≈1–2 µs) driver encode call is
// paid once per distinct (pointer, shape, box) and afterwards a hash lookup returns the 128-byte map.
#include <cuda.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "tmap.cuh"

namespace vitk {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      ptr = nullptr;
    return reinterpret_cast<EncodeTiledFn>(ptr);
  }();
  return fn;
}

struct TmapKey {
  uint64_t v[12];
  bool operator==(const TmapKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (uint64_t x : k.v) { h ^= x; h *= 1099511628211ull; h ^= h >> 29; }
    return static_cast<size_t>(h);
  }
};

static std::mutex g_mu;
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_cache;

int get_tensor_map(CUtensorMap* out, const void* base, TmapDtype dtype, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, TmapSwizzle swizzle) {
  VITK_REQUIRE(rank >= 2 && rank <= 3, VITK_EINVAL, "tensor map rank %d unsupported", rank);
  TmapKey key;
  memset(&key, 0, sizeof(key));
  int dev = 0;
  cudaGetDevice(&dev);
  key.v[0] = reinterpret_cast<uint64_t>(base);
  key.v[1] = (static_cast<uint64_t>(dtype) << 40) | (static_cast<uint64_t>(swizzle) << 32) | (static_cast<uint64_t>(rank) << 24) |
             static_cast<uint64_t>(dev);
  for (int i = 0; i < rank; ++i) { key.v[2 + i] = dims[i]; key.v[8 + i] = box[i]; }
  for (int i = 0; i < rank - 1; ++i) key.v[5 + i] = strides_bytes[i];
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_cache.find(key);
    if (it != g_cache.end()) { *out = it->second; return 0; }
  }
  EncodeTiledFn fn = encode_fn();
  VITK_REQUIRE(fn != nullptr, VITK_EDRIVER, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[3];
  cuuint64_t gstr[2];
  cuuint32_t bx[3];
  cuuint32_t es[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 0; i < rank - 1; ++i) gstr[i] = strides_bytes[i];
  const CUtensorMapDataType dt = dtype == TM_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const CUtensorMapSwizzle sw = swizzle == TM_SW128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle == TM_SW64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(out, dt, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VITK_REQUIRE(r == CUDA_SUCCESS, VITK_EDRIVER, "cuTensorMapEncodeTiled failed (CUresult %d)", static_cast<int>(r));
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_cache.size() > 16384) g_cache.clear();
    g_cache.emplace(key, *out);
  }
  return 0;
}

int make_tensor_map_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box) {
  return get_tensor_map(out, base, TM_BF16, rank, dims, strides_bytes, box, TM_SW128);
}

}  // namespace vitk
