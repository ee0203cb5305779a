// Internal: TMA descriptor construction shared by the GEMM and attention translation units.
#pragma once
#include <cuda.h>

namespace rb {

CUtensorMap make_map_3d_bf16(const void* ptr, long long d0, long long d1, long long d2, long long s1_elems, long long s2_elems, int b0,
                             int b1, int b2);

}  // namespace rb
