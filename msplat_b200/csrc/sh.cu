// msplat_b200/csrc/sh.cu -- spherical-harmonics colour evaluation, degree 0..10, any channel count.
//
// Replaces computeSHForward/Backward (/root/reference/msplat/src/compute_sh.cu:1600-1754):
//   value[p, c]    = sum_d Y_d(dir_p) * shs[p, c, d]                 (no +0.5, no clamp)
//   dL_dshs[p,c,d] = Y_d(dir_p) * dL_dvalue[p, c]
//   dL_ddir[p]     = sum_d (sum_c dL_dvalue[p,c] shs[p,c,d]) * grad Y_d(dir_p)
// Invisible Gaussians produce explicit zeros (the reference relies on pre-zeroed outputs).
//
// Design.  shs is [P, Cs, D] with D innermost, i.e. a stream of contiguous D-float rows.  A
// block owns GPB consecutive Gaussians.  Phase 1: one thread per Gaussian evaluates the basis
// ONCE (generated straight-line code, sh_eval.cuh) into shared memory, transposed [D][GS] so
// the writes are conflict-free.  Phase 2: a group of LPR lanes owns one Gaussian, keeps its
// slice of the basis in registers and walks the Gaussian's Cs rows; each row is read with
// 16-byte (D % 4 == 0) or coalesced 4-byte loads, reduced over the group with shuffles.  The
// backward pass additionally streams dL_dshs out and accumulates w_d = sum_c dL_dvalue * shs in
// registers -> shared memory; Phase 3: one thread per Gaussian contracts w with grad Y.
// HBM-bound for deg <= 3 (4*Cs*D bytes per Gaussian each way); all indexing is 64-bit
// (shs[1M,32,121] has 3.9e9 elements).
#include "sh_eval.cuh"
#include "sh_layout.cuh"

namespace msb {

constexpr int SH_NT = 256;

template <int DEG, bool BWD>
__global__ void __launch_bounds__(SH_NT) sh_kernel(int P, int Cs, int GPB, const float* __restrict__ shs,
                                                   const float* __restrict__ dirs,
                                                   const uint8_t* __restrict__ visible,
                                                   const float* __restrict__ dL_dvalue,  // BWD only
                                                   float* __restrict__ value,            // FWD only
                                                   float* __restrict__ dL_dshs,          // BWD only
                                                   float* __restrict__ dL_ddirs) {       // BWD only
    constexpr int D = sh_dim(DEG);
    constexpr bool VEC = sh_vec(DEG);
    constexpr int LPR = sh_lpr(DEG);
    constexpr int IT = sh_iters(DEG);
    constexpr int UNITS = sh_units(DEG);
    constexpr int W = VEC ? 4 : 1;
    constexpr int GROUPS = SH_NT / LPR;
    extern __shared__ float smem[];
    const int GS = GPB | 1;
    float* Bs = smem;            // [D][GS]
    float* Ws = smem + D * GS;   // [D][GS]  (BWD)
    const long long g0 = (long long)blockIdx.x * GPB;
    const int ng = (int)min((long long)GPB, (long long)P - g0);

    // ---- phase 1: basis, one thread per Gaussian ------------------------------------------------
    for (int t = threadIdx.x; t < ng; t += SH_NT) {
        const float* d = dirs + (g0 + t) * 3;
        sh_basis<DEG>(__ldg(d), __ldg(d + 1), __ldg(d + 2), Bs + t, GS);
    }
    __syncthreads();

    // ---- phase 2: one LPR-lane group per Gaussian -----------------------------------------------
    const int grp = threadIdx.x / LPR, s = threadIdx.x % LPR;
    for (int gl0 = 0; gl0 < ng; gl0 += GROUPS) {  // uniform trip count across the block
        const int gl = gl0 + grp;
        const bool act = gl < ng;
        const long long g = g0 + (act ? gl : 0);
        const bool vis = act && (visible == nullptr || visible[g] != 0);
        float b[IT * W];
#pragma unroll
        for (int it = 0; it < IT; ++it) {
            const int u = s + it * LPR;
#pragma unroll
            for (int k = 0; k < W; ++k) b[it * W + k] = (act && u < UNITS) ? Bs[(u * W + k) * GS + gl] : 0.f;
        }
        float wacc[IT * W];
#pragma unroll
        for (int i = 0; i < IT * W; ++i) wacc[i] = 0.f;
        const long long row0 = g * Cs;
        for (int c = 0; c < Cs; ++c) {
            const float* rp = shs + (row0 + c) * D;
            float sv[IT * W];
#pragma unroll
            for (int it = 0; it < IT; ++it) {
                const int u = s + it * LPR;
                if (VEC) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (vis && u < UNITS) v = ldg_stream4(reinterpret_cast<const float4*>(rp) + u);
                    sv[it * W + 0] = v.x;
                    sv[it * W + (W > 1 ? 1 : 0)] = v.y;
                    sv[it * W + (W > 2 ? 2 : 0)] = v.z;
                    sv[it * W + (W > 3 ? 3 : 0)] = v.w;
                } else {
                    sv[it * W] = (vis && u < UNITS) ? __ldg(rp + u) : 0.f;
                }
            }
            if (!BWD) {
                float acc = 0.f;
#pragma unroll
                for (int i = 0; i < IT * W; ++i) acc = fmaf(sv[i], b[i], acc);
                acc = group_sum<LPR>(acc);
                if (act && s == 0) value[row0 + c] = vis ? acc : 0.f;
            } else {
                const float dv = vis ? __ldg(dL_dvalue + row0 + c) : 0.f;
                float* op = dL_dshs + (row0 + c) * D;
#pragma unroll
                for (int it = 0; it < IT; ++it) {
                    const int u = s + it * LPR;
                    if (VEC) {
                        if (act && u < UNITS)
                            reinterpret_cast<float4*>(op)[u] =
                                make_float4(b[it * W] * dv, b[it * W + (W > 1 ? 1 : 0)] * dv,
                                            b[it * W + (W > 2 ? 2 : 0)] * dv, b[it * W + (W > 3 ? 3 : 0)] * dv);
                    } else {
                        if (act && u < UNITS) op[u] = b[it * W] * dv;
                    }
                }
#pragma unroll
                for (int i = 0; i < IT * W; ++i) wacc[i] = fmaf(sv[i], dv, wacc[i]);
            }
        }
        if (BWD) {
#pragma unroll
            for (int it = 0; it < IT; ++it) {
                const int u = s + it * LPR;
#pragma unroll
                for (int k = 0; k < W; ++k)
                    if (act && u < UNITS) Ws[(u * W + k) * GS + gl] = wacc[it * W + k];
            }
        }
    }
    if (!BWD) return;
    __syncthreads();

    // ---- phase 3 (BWD): dL_ddir, one thread per Gaussian ------------------------------------------
    for (int t = threadIdx.x; t < ng; t += SH_NT) {
        const long long g = g0 + t;
        float gx = 0.f, gy = 0.f, gz = 0.f;
        if (visible == nullptr || visible[g] != 0) {
            const float* d = dirs + g * 3;
            sh_basis_grad<DEG>(__ldg(d), __ldg(d + 1), __ldg(d + 2), Ws + t, GS, gx, gy, gz);
        }
        dL_ddirs[g * 3 + 0] = gx;
        dL_ddirs[g * 3 + 1] = gy;
        dL_ddirs[g * 3 + 2] = gz;
    }
}

static int sh_gpb(int deg) {
    const int D = sh_dim(deg), groups = SH_NT / sh_lpr(deg);
    int k = (24 * 1024) / (D * 4 * groups);
    if (k < 1) k = 1;
    if (k > 4) k = 4;
    return groups * k;
}

template <int DEG, bool BWD>
static int sh_launch(int P, int Cs, const float* shs, const float* dirs, const uint8_t* visible,
                     const float* dL_dvalue, float* value, float* dL_dshs, float* dL_ddirs, cudaStream_t st) {
    const int GPB = sh_gpb(DEG);
    const int GS = GPB | 1;
    const size_t smem = (size_t)sh_dim(DEG) * GS * sizeof(float) * (BWD ? 2 : 1);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(sh_kernel<DEG, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return set_error((int)e, "compute_sh: cudaFuncSetAttribute failed");
    }
    const unsigned grid = (unsigned)(((long long)P + GPB - 1) / GPB);
    sh_kernel<DEG, BWD><<<grid, SH_NT, smem, st>>>(P, Cs, GPB, shs, dirs, visible, dL_dvalue, value, dL_dshs,
                                                   dL_ddirs);
    return check_launch(BWD ? "compute_sh_bwd" : "compute_sh_fwd");
}

template <bool BWD>
static int sh_dispatch(int deg, int P, int Cs, const float* shs, const float* dirs, const uint8_t* visible,
                       const float* dL_dvalue, float* value, float* dL_dshs, float* dL_ddirs, cudaStream_t st) {
    switch (deg) {
#define MSB_SH_CASE(d) \
    case d:            \
        return sh_launch<d, BWD>(P, Cs, shs, dirs, visible, dL_dvalue, value, dL_dshs, dL_ddirs, st);
        MSB_SH_CASE(0)
        MSB_SH_CASE(1)
        MSB_SH_CASE(2)
        MSB_SH_CASE(3)
        MSB_SH_CASE(4)
        MSB_SH_CASE(5)
        MSB_SH_CASE(6)
        MSB_SH_CASE(7)
        MSB_SH_CASE(8)
        MSB_SH_CASE(9)
        MSB_SH_CASE(10)
#undef MSB_SH_CASE
    }
    return set_error(MSB_ERR_ARG, "compute_sh: D must be (deg+1)^2 with 0 <= deg <= 10");
}

}  // namespace msb

using namespace msb;

extern "C" {

int msb_compute_sh_fwd(const float* shs, const float* dirs, const uint8_t* visible, int P, int Cs, int D,
                       float* value, void* stream) {
    if (P == 0 || Cs == 0) return MSB_OK;
    if (!(P > 0 && Cs > 0 && shs && dirs && value)) return set_error(MSB_ERR_ARG, "compute_sh_fwd: bad argument");
    if ((reinterpret_cast<uintptr_t>(shs) & 15u) != 0) return set_error(MSB_ERR_ARG, "compute_sh_fwd: shs alignment");
    return sh_dispatch<false>(sh_degree_of(D), P, Cs, shs, dirs, visible, nullptr, value, nullptr, nullptr,
                              (cudaStream_t)stream);
}

int msb_compute_sh_bwd(const float* shs, const float* dirs, const uint8_t* visible, const float* dL_dvalue,
                       int P, int Cs, int D, float* dL_dshs, float* dL_ddirs, void* stream) {
    if (P == 0) return MSB_OK;
    if (!(P > 0 && Cs >= 0 && shs && dirs && dL_dvalue && dL_dshs && dL_ddirs))
        return set_error(MSB_ERR_ARG, "compute_sh_bwd: bad argument");
    if (((reinterpret_cast<uintptr_t>(shs) | reinterpret_cast<uintptr_t>(dL_dshs)) & 15u) != 0)
        return set_error(MSB_ERR_ARG, "compute_sh_bwd: shs/dL_dshs alignment");
    return sh_dispatch<true>(sh_degree_of(D), P, Cs, shs, dirs, visible, dL_dvalue, nullptr, dL_dshs, dL_ddirs,
                             (cudaStream_t)stream);
}

}  // extern "C"
