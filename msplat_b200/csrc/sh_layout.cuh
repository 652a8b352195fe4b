// msplat_b200/csrc/sh_layout.cuh -- how a [Cs, D] block of SH coefficients is split over lanes
// (shared by sh.cu and the fused render preprocess in render.cu).
#pragma once

namespace msb {

__host__ __device__ constexpr int sh_dim(int deg) { return (deg + 1) * (deg + 1); }
__host__ __device__ constexpr bool sh_vec(int deg) { return sh_dim(deg) % 4 == 0; }
// lanes per row
__host__ __device__ constexpr int sh_lpr(int deg) {
    return deg == 0 ? 1 : deg == 1 ? 1 : deg == 2 ? 4 : deg == 3 ? 4 : deg == 4 ? 8 : deg == 5 ? 4 :
           deg == 6 ? 16 : deg == 7 ? 16 : deg == 8 ? 32 : deg == 9 ? 8 : 32;
}
// units (float4 or float) per row and iterations per lane
__host__ __device__ constexpr int sh_units(int deg) { return sh_vec(deg) ? sh_dim(deg) / 4 : sh_dim(deg); }
__host__ __device__ constexpr int sh_iters(int deg) { return (sh_units(deg) + sh_lpr(deg) - 1) / sh_lpr(deg); }

template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

static inline int sh_degree_of(int D) {
    for (int d = 0; d <= 10; ++d)
        if ((d + 1) * (d + 1) == D) return d;
    return -1;
}

}  // namespace msb
