// msplat_b200/csrc/common.cuh -- shared device helpers for the sm_100a kernels.
//
// Numerics policy.  The library is compiled WITHOUT --use_fast_math but with -ftz=true.
// Stages whose outputs feed integer decisions that must be bit-identical to the reference
// (uv/depth -> sort keys, cov3d -> radius/tiles, blend power/alpha -> skip/terminate tests)
// are written with explicit round-to-nearest intrinsics (__fmul_rn/__fmaf_rn/__fadd_rn are
// never contracted or re-associated by nvcc) and explicit MUFU approximations, in the exact
// operation order the reference's `-O3 --use_fast_math` sm_100 build executes (SASS dataflow
// recorded in DESIGN.md "Numerics mirrored from the reference build").  Everything else is
// ordinary FP32 with FMA contraction.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define MSB_TILE 16          // reference: msplat/include/config.h:7-8 (BLOCK_X = BLOCK_Y = 16)
#define MSB_TILE_PIX 256

namespace msb {

// ---- host/device portability ---------------------------------------------------------------
// The per-Gaussian math headers (geom.cuh, sh_eval.cuh, blend_math.cuh) also compile for the
// host so that tools/host_check.cu can exercise exactly the same source on a CPU-only box
// (the -m "not gpu" tests).  On the host the MUFU approximations become IEEE operations, so
// host results are tolerance-level, not bit-level, checks.
#define MSB_HD __host__ __device__ __forceinline__

#ifdef __CUDA_ARCH__
// explicit round-to-nearest ops: never contracted / re-associated by nvcc
MSB_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
MSB_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
MSB_HD float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
MSB_HD int f2i_rz(float x) { return __float2int_rz(x); }
MSB_HD int f2i_ru(float x) { return __float2int_ru(x); }
MSB_HD float ldg_f(const float* p) { return __ldg(p); }
// MUFU approximations (what --use_fast_math lowers `/`, sqrtf, expf to)
MSB_HD float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
MSB_HD float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
MSB_HD float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// max/min with the FMNMX NaN rule (returns the non-NaN operand)
MSB_HD float fmax_ftz(float a, float b) {
    float y;
    asm("max.ftz.f32 %0, %1, %2;" : "=f"(y) : "f"(a), "f"(b));
    return y;
}
MSB_HD float fmin_ftz(float a, float b) {
    float y;
    asm("min.ftz.f32 %0, %1, %2;" : "=f"(y) : "f"(a), "f"(b));
    return y;
}
#else
}  // namespace msb
#include <math.h>
namespace msb {
MSB_HD float fmul(float a, float b) { volatile float r = a * b; return r; }
MSB_HD float fadd(float a, float b) { volatile float r = a + b; return r; }
MSB_HD float ffma(float a, float b, float c) { return fmaf(a, b, c); }
MSB_HD int f2i_sat(double x) {
    if (x != x) return 0;
    if (x >= 2147483647.0) return 2147483647;
    if (x <= -2147483648.0) return (int)(-2147483647 - 1);
    return (int)x;
}
MSB_HD int f2i_rz(float x) { return f2i_sat(trunc((double)x)); }
MSB_HD int f2i_ru(float x) { return f2i_sat(ceil((double)x)); }
MSB_HD float ldg_f(const float* p) { return *p; }
MSB_HD float rcp_approx(float x) { return 1.0f / x; }
MSB_HD float sqrt_approx(float x) { return sqrtf(x); }
MSB_HD float ex2_approx(float x) { return exp2f(x); }
MSB_HD float fmax_ftz(float a, float b) { return fmaxf(a, b); }
MSB_HD float fmin_ftz(float a, float b) { return fminf(a, b); }
#endif
MSB_HD int imin(int a, int b) { return a < b ? a : b; }
MSB_HD int imax(int a, int b) { return a > b ? a : b; }

// ---- packed FP32 pairs (sm_100: FFMA2 / FMUL2 / FADD2, one issue slot for two lanes of math) -----
// Round-to-nearest, FTZ: bit-identical to the scalar __fmaf_rn / __fmul_rn / __fadd_rn results.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// ---- streaming global access ------------------------------------------------------------
__device__ __forceinline__ float4 ldg_stream4(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}

// ---- cp.async (LDGSTS) ------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- vector reductions to global memory (fire and forget) ------------------------------------
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_v2(float* p, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) + mbarrier ----------------------------------
// One elected thread moves a contiguous, 16-byte aligned slab between global and shared memory
// through the async proxy: no per-thread address arithmetic, no registers, and the copy engine
// keeps the whole slab in flight.  Sizes must be multiples of 16 bytes.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "MBAR_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra MBAR_DONE;\n\t"
        "bra MBAR_WAIT;\n\t"
        "MBAR_DONE:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
// gmem_dst[i] += smem_src[i] (FP32), performed by the copy engine at the L2
__device__ __forceinline__ void bulk_s2g_add_f32(void* gmem_dst, const void* smem_src, unsigned bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gmem_dst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// makes this thread's generic-proxy shared-memory writes visible to the async proxy (before a bulk store)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- block-cooperative AoS slab staging ----------------------------------------------------
// A block owns rows [row0, row0+rows) of a row-major [P, K] float array.  The block's slab
// is contiguous in memory, so it is moved with fully coalesced 16-byte accesses (scalar for
// the <4-float tail of the last block) into shared memory, where each thread then reads its
// own row with stride K (conflict-free for odd K; rows with even K are read as vectors).
// This is the "vectorised 16-byte loads" path for the [P,2]/[P,3]/[P,4]/[P,6] tensors of
// the msplat API.  Requires g + first to be 16-byte aligned (base pointers are checked by
// the C-ABI entry points; first = blockIdx * 256 * K is a multiple of 4).
template <int NT>
__device__ __forceinline__ void slab_load(float* __restrict__ smem, const float* __restrict__ g,
                                          long long first, int count) {
    const float* src = g + first;
    const int nvec = count >> 2;
    const float4* v = reinterpret_cast<const float4*>(src);
    float4* d = reinterpret_cast<float4*>(smem);
    for (int i = threadIdx.x; i < nvec; i += NT) d[i] = ldg_stream4(v + i);
    for (int i = 4 * nvec + threadIdx.x; i < count; i += NT) smem[i] = __ldg(src + i);
}

template <int NT>
__device__ __forceinline__ void slab_store(float* __restrict__ g, const float* __restrict__ smem,
                                           long long first, int count) {
    float* dst = g + first;
    const int nvec = count >> 2;
    float4* v = reinterpret_cast<float4*>(dst);
    const float4* s = reinterpret_cast<const float4*>(smem);
    for (int i = threadIdx.x; i < nvec; i += NT) v[i] = s[i];
    for (int i = 4 * nvec + threadIdx.x; i < count; i += NT) dst[i] = smem[i];
}

// ---- slab sets: TMA bulk copies for full blocks, per-thread path for the ragged last block ----------
// d[i] = {shared-memory slab, global [P, k] array, k}.  A full block (rows == NT, 16-byte aligned
// bases) moves all slabs with cp.async.bulk issued by thread 0 and waits on one mbarrier (loads) or
// on the bulk group (stores); otherwise slab_load / slab_store.  Both end with the block in sync.
struct SlabIn {
    float* s;
    const float* g;
    int k;
};
struct SlabOut {
    const float* s;
    float* g;
    int k;
};
__device__ __forceinline__ bool al16(const void* p) { return (reinterpret_cast<unsigned long long>(p) & 15ull) == 0ull; }

template <int NT, int N>
__device__ __forceinline__ void slabs_load(unsigned long long* bar, const SlabIn (&d)[N], long long row0, int rows) {
    bool full = rows == NT;
#pragma unroll
    for (int i = 0; i < N; ++i) full = full && al16(d[i].g);
    if (threadIdx.x == 0 && full) mbar_init(bar, 1);
    __syncthreads();
    if (full) {
        if (threadIdx.x == 0) {
            unsigned bytes = 0;
#pragma unroll
            for (int i = 0; i < N; ++i) bytes += (unsigned)(d[i].k * NT * sizeof(float));
            mbar_expect_tx(bar, bytes);
#pragma unroll
            for (int i = 0; i < N; ++i)
                bulk_g2s(d[i].s, d[i].g + row0 * d[i].k, (unsigned)(d[i].k * NT * sizeof(float)), bar);
        }
        mbar_wait(bar, 0);
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) slab_load<NT>(d[i].s, d[i].g, row0 * d[i].k, rows * d[i].k);
        __syncthreads();
    }
}

// Call after the threads have written their rows into the shared-memory slabs (no barrier needed before).
template <int NT, int N>
__device__ __forceinline__ void slabs_store(const SlabOut (&d)[N], long long row0, int rows) {
    bool full = rows == NT;
#pragma unroll
    for (int i = 0; i < N; ++i) full = full && al16(d[i].g);
    if (full) {
        fence_async_smem();  // this thread's slab writes -> visible to the async proxy
        __syncthreads();
        if (threadIdx.x == 0) {
#pragma unroll
            for (int i = 0; i < N; ++i)
                bulk_s2g(d[i].g + row0 * d[i].k, d[i].s, (unsigned)(d[i].k * NT * sizeof(float)));
            bulk_commit();
            bulk_wait_read();  // shared memory must stay valid until the copy engine has read it
        }
    } else {
        __syncthreads();
#pragma unroll
        for (int i = 0; i < N; ++i) slab_store<NT>(d[i].g, d[i].s, row0 * d[i].k, rows * d[i].k);
    }
}

// ---- warp helpers -----------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }

}  // namespace msb

// ---- the public C ABI: every translation unit sees the prototypes it defines, so a definition that
// drifts from include/msplat_b200.h is a compile error (conflicting C-linkage declarations) ---------
#include "../../include/msplat_b200.h"

// ---- host-side error plumbing (defined in capi.cu) ------------------------------------------
namespace msb {
int set_error(int code, const char* msg);
int check_launch(const char* what);
}  // namespace msb

#define MSB_OK 0
#define MSB_ERR_ARG (-1)
#define MSB_ERR_WORKSPACE (-2)
#define MSB_ERR_RANGE (-3)
