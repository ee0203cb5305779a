// Internal: TMA descriptor construction shared by the GEMM and attention translation units.
#pragma once
#include <cuda.h>

namespace rb {

CUtensorMap make_map_3d_bf16(const void* ptr, long long d0, long long d1, long long d2, long long s1_elems, long long s2_elems, int b0,
                             int b1, int b2);

// 2-D tensor map with the 128-byte swizzle: `inner` contiguous elements (esize 1 or 2 bytes) per row, `outer` rows `ld` elements apart
CUtensorMap make_map_2d_sw128(const void* ptr, long long inner, long long outer, long long ld, int box_inner, int box_outer, int esize);

}  // namespace rb
