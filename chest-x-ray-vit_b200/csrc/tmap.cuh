// Cached CUtensorMap construction (tmap.cu).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace vitk {

enum TmapDtype { TM_BF16 = 0, TM_F32 = 1 };
enum TmapSwizzle { TM_SW128 = 0, TM_SW64 = 1, TM_SWNONE = 2 };

// dims / box innermost first; strides_bytes has rank-1 entries (dimension 0 is contiguous).
int get_tensor_map(CUtensorMap* out, const void* base, TmapDtype dtype, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, TmapSwizzle swizzle);
// bf16, 128-byte swizzle (operand tiles of the GEMM and attention kernels)
int make_tensor_map_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box);

}  // namespace vitk
