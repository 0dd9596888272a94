// countr_b200 — host-side TMA tensor-map construction.
//
// cuTensorMapEncodeTiled is a driver entry point; it is resolved at run time through
// cudaGetDriverEntryPoint so the library carries no link-time dependency on libcuda
// (it has to load on a CPU-only box for the symbol-export test).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace countr {

enum TmapSwizzle : int { TMAP_SW_NONE = 0, TMAP_SW_32 = 1, TMAP_SW_64 = 2, TMAP_SW_128 = 3 };

// Rank-4 tiled tensor map over 16-bit elements.
//   dims[0] is the contiguous dimension; strides[i] (elements) belongs to dims[i], strides[0]==1.
//   box[i] is the tile extent in dims[i].  Out-of-bounds elements are zero-filled on load,
//   which is what implements K/M/N tails and the 3x3 convolution halo.
int make_tmap_4d_16b(CUtensorMap* out, const void* base, const uint64_t dims[4],
                     const uint64_t strides[4], const uint32_t box[4], TmapSwizzle swizzle);

// Same for any element size (2: 16-bit payload, 4: fp32 — the fp32 type matters for TMA reduce-add stores).
int make_tmap_4d(CUtensorMap* out, const void* base, int elt_bytes, const uint64_t dims[4], const uint64_t strides[4],
                 const uint32_t box[4], TmapSwizzle swizzle);

int num_sms();

}  // namespace countr
