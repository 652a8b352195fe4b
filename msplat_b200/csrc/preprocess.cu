// msplat_b200/csrc/preprocess.cu -- per-Gaussian kernels (HBM-bound streaming).
//
// Replaces the reference kernels K1..K6:
//   projectPointForward/Backward   /root/reference/msplat/src/project_point.cu:13-145
//   computeCov3DForward/Backward   /root/reference/msplat/src/compute_cov3d.cu:119-147
//   EWAProjectForward/Back         /root/reference/msplat/src/ewa_project.cu:16-252
// plus a fused forward (K1+K3+K5) and fused backward (K6+K4+K2) used by rasterization().
//
// Layout / roofline: every tensor of the msplat API is row-major [P, K] with small K.  A block
// of 256 threads owns 256 consecutive Gaussians; each [256, K] slab is contiguous, so it is
// moved HBM <-> shared memory by TMA bulk copies (common.cuh slabs_load/slabs_store; coalesced
// 16-byte per-thread accesses for the ragged last block) and
// each thread then works on its own row from shared memory.  Outputs for culled / invisible /
// degenerate Gaussians are written as explicit zeros (the reference relies on pre-zeroed
// tensors: project_point.cu:161-162, ewa_project.cu:274-276), so no memset pass is needed.
// Camera gradients are reduced warp -> block -> 16 atomics per block instead of 24-25 same-
// address atomics per thread (project_point.cu:107-144, ewa_project.cu:206-246).
#include "geom.cuh"

namespace msb {

constexpr int NT = 256;

// ------------------------------------------------------------------------------------------
// K1  project_point forward
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) project_point_fwd_kernel(int P, const float* __restrict__ xyz,
                                                               const float* __restrict__ intr,
                                                               const float* __restrict__ extr, int W, int H,
                                                               float nearest, float extent,
                                                               float* __restrict__ uv, float* __restrict__ depth) {
    __shared__ __align__(16) float s[NT * 3];
    const long long row0 = (long long)blockIdx.x * NT;
    const int rows = (int)min((long long)NT, P - row0);
    const Cam c = load_cam(intr, extr);
    __shared__ unsigned long long s_bar;
    {
        const SlabIn in[] = {{s, xyz, 3}};
        slabs_load<NT>(&s_bar, in, row0, rows);
    }
    const int t = threadIdx.x;
    float u = 0.f, v = 0.f, d = 0.f;
    if (t < rows) {
        const float px = s[3 * t], py = s[3 * t + 1], pz = s[3 * t + 2];
        if (!project_fwd(c, px, py, pz, W, H, nearest, extent, u, v, d)) u = v = d = 0.f;
    }
    __syncthreads();
    if (t < rows) {
        s[2 * t] = u;
        s[2 * t + 1] = v;
        depth[row0 + t] = d;
    }
    {
        const SlabOut out[] = {{s, uv, 2}};
        slabs_store<NT>(out, row0, rows);
    }
}

// ------------------------------------------------------------------------------------------
// K2  project_point backward
// ------------------------------------------------------------------------------------------
template <bool CAM>
__global__ void __launch_bounds__(NT) project_point_bwd_kernel(
    int P, const float* __restrict__ xyz, const float* __restrict__ intr, const float* __restrict__ extr,
    const float* __restrict__ depth, const float* __restrict__ dL_duv, const float* __restrict__ dL_ddepth,
    float* __restrict__ dL_dxyz, float* __restrict__ dL_dintr, float* __restrict__ dL_dextr) {
    __shared__ __align__(16) float s_xyz[NT * 3];
    __shared__ __align__(16) float s_guv[NT * 2];
    __shared__ float s_red[8 * 16];
    const long long row0 = (long long)blockIdx.x * NT;
    const int rows = (int)min((long long)NT, P - row0);
    const Cam c = load_cam(intr, extr);
    __shared__ unsigned long long s_bar;
    {
        const SlabIn in[] = {{s_xyz, xyz, 3}, {s_guv, dL_duv, 2}};
        slabs_load<NT>(&s_bar, in, row0, rows);
    }
    const int t = threadIdx.x;
    float cam[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) cam[i] = 0.f;
    float dx = 0.f, dy = 0.f, dz = 0.f;
    if (t < rows && depth[row0 + t] != 0.f) {  // project_point.cu:74-75: depth == 0 means culled
        project_bwd<CAM>(c, s_xyz[3 * t], s_xyz[3 * t + 1], s_xyz[3 * t + 2], s_guv[2 * t], s_guv[2 * t + 1],
                         dL_ddepth[row0 + t], dx, dy, dz, cam);
    }
    __syncthreads();
    if (t < rows) {
        s_xyz[3 * t] = dx;
        s_xyz[3 * t + 1] = dy;
        s_xyz[3 * t + 2] = dz;
    }
    {
        const SlabOut out[] = {{s_xyz, dL_dxyz, 3}};
        slabs_store<NT>(out, row0, rows);
    }
    if (CAM) cam_reduce_atomic(cam, dL_dintr, dL_dextr, s_red);
}

// ------------------------------------------------------------------------------------------
// K3 / K4  compute_cov3d forward / backward
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) cov3d_fwd_kernel(int P, const float* __restrict__ scale,
                                                       const float* __restrict__ quat,
                                                       const uint8_t* __restrict__ visible,
                                                       float* __restrict__ cov3d) {
    __shared__ __align__(16) float s[NT * 7];  // in: scale(3) + quat(4); out: cov(6)
    float* s_scale = s;
    float* s_quat = s + NT * 3;
    const long long row0 = (long long)blockIdx.x * NT;
    const int rows = (int)min((long long)NT, P - row0);
    __shared__ unsigned long long s_bar;
    {
        const SlabIn in[] = {{s_scale, scale, 3}, {s_quat, quat, 4}};
        slabs_load<NT>(&s_bar, in, row0, rows);
    }
    const int t = threadIdx.x;
    float cv[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (t < rows && (visible == nullptr || visible[row0 + t])) {
        const float4 q = reinterpret_cast<const float4*>(s_quat)[t];
        cov3d_fwd(s_scale[3 * t], s_scale[3 * t + 1], s_scale[3 * t + 2], q.x, q.y, q.z, q.w, cv);
    }
    __syncthreads();
    if (t < rows) {
        float2* o = reinterpret_cast<float2*>(s) + 3 * t;
        o[0] = make_float2(cv[0], cv[1]);
        o[1] = make_float2(cv[2], cv[3]);
        o[2] = make_float2(cv[4], cv[5]);
    }
    {
        const SlabOut out[] = {{s, cov3d, 6}};
        slabs_store<NT>(out, row0, rows);
    }
}

__global__ void __launch_bounds__(NT) cov3d_bwd_kernel(int P, const float* __restrict__ scale,
                                                       const float* __restrict__ quat,
                                                       const uint8_t* __restrict__ visible,
                                                       const float* __restrict__ dL_dcov3d,
                                                       float* __restrict__ dL_dscale,
                                                       float* __restrict__ dL_dquat) {
    __shared__ __align__(16) float s[NT * 13];
    float* s_scale = s;
    float* s_quat = s + NT * 3;
    float* s_g = s + NT * 7;
    const long long row0 = (long long)blockIdx.x * NT;
    const int rows = (int)min((long long)NT, P - row0);
    __shared__ unsigned long long s_bar;
    {
        const SlabIn in[] = {{s_scale, scale, 3}, {s_quat, quat, 4}, {s_g, dL_dcov3d, 6}};
        slabs_load<NT>(&s_bar, in, row0, rows);
    }
    const int t = threadIdx.x;
    float ds[3] = {0.f, 0.f, 0.f}, dq[4] = {0.f, 0.f, 0.f, 0.f};
    if (t < rows && (visible == nullptr || visible[row0 + t])) {
        const float4 q = reinterpret_cast<const float4*>(s_quat)[t];
        const float2* gp = reinterpret_cast<const float2*>(s_g) + 3 * t;
        const float2 g0 = gp[0], g1 = gp[1], g2 = gp[2];
        const float g[6] = {g0.x, g0.y, g1.x, g1.y, g2.x, g2.y};
        cov3d_bwd(s_scale[3 * t], s_scale[3 * t + 1], s_scale[3 * t + 2], q.x, q.y, q.z, q.w, g, ds, dq);
    }
    __syncthreads();
    if (t < rows) {
        s_scale[3 * t] = ds[0];
        s_scale[3 * t + 1] = ds[1];
        s_scale[3 * t + 2] = ds[2];
        reinterpret_cast<float4*>(s_quat)[t] = make_float4(dq[0], dq[1], dq[2], dq[3]);
    }
    {
        const SlabOut out[] = {{s_scale, dL_dscale, 3}, {s_quat, dL_dquat, 4}};
        slabs_store<NT>(out, row0, rows);
    }
}

// ------------------------------------------------------------------------------------------
// K5  ewa_project forward
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) ewa_fwd_kernel(int P, const float* __restrict__ xyz,
                                                     const float* __restrict__ cov3d,
                                                     const float* __restrict__ intr,
                                                     const float* __restrict__ extr,
                                                     const float* __restrict__ uv,
                                                     const uint8_t* __restrict__ visible, int gx, int gy,
                                                     float* __restrict__ conic, int* __restrict__ radius,
                                                     int* __restrict__ tiles) {
    __shared__ __align__(16) float s[NT * 11];
    float* s_xyz = s;
    float* s_cov = s + NT * 3;
    float* s_uv = s + NT * 9;
    const long long row0 = (long long)blockIdx.x * NT;
    const int rows = (int)min((long long)NT, P - row0);
    const Cam c = load_cam(intr, extr);
    __shared__ unsigned long long s_bar;
    {
        const SlabIn in[] = {{s_xyz, xyz, 3}, {s_cov, cov3d, 6}, {s_uv, uv, 2}};
        slabs_load<NT>(&s_bar, in, row0, rows);
    }
    const int t = threadIdx.x;
    float cx = 0.f, cy = 0.f, cz = 0.f;
    int rad = 0, til = 0;
    if (t < rows && (visible == nullptr || visible[row0 + t])) {
        const float2* cp = reinterpret_cast<const float2*>(s_cov) + 3 * t;
        const float2 c0 = cp[0], c1 = cp[1], c2 = cp[2];
        const float cv[6] = {c0.x, c0.y, c1.x, c1.y, c2.x, c2.y};
        const float2 p2 = reinterpret_cast<const float2*>(s_uv)[t];
        if (!ewa_fwd(c, s_xyz[3 * t], s_xyz[3 * t + 1], s_xyz[3 * t + 2], cv, p2.x, p2.y, gx, gy, cx, cy, cz,
                     rad, til)) {
            cx = cy = cz = 0.f;
            rad = til = 0;
        }
    }
    __syncthreads();
    if (t < rows) {
        s_xyz[3 * t] = cx;
        s_xyz[3 * t + 1] = cy;
        s_xyz[3 * t + 2] = cz;
        radius[row0 + t] = rad;
        tiles[row0 + t] = til;
    }
    {
        const SlabOut out[] = {{s_xyz, conic, 3}};
        slabs_store<NT>(out, row0, rows);
    }
}

// ------------------------------------------------------------------------------------------
// K6  ewa_project backward
// ------------------------------------------------------------------------------------------
template <bool CAM>
__global__ void __launch_bounds__(NT) ewa_bwd_kernel(int P, const float* __restrict__ xyz,
                                                     const float* __restrict__ cov3d,
                                                     const float* __restrict__ intr,
                                                     const float* __restrict__ extr,
                                                     const int* __restrict__ radius,
                                                     const float* __restrict__ dL_dconic,
                                                     float* __restrict__ dL_dxyz,
                                                     float* __restrict__ dL_dcov3d,
                                                     float* __restrict__ dL_dintr,
                                                     float* __restrict__ dL_dextr) {
    __shared__ __align__(16) float s[NT * 12];
    __shared__ float s_red[8 * 16];
    float* s_xyz = s;
    float* s_cov = s + NT * 3;
    float* s_gc = s + NT * 9;
    const long long row0 = (long long)blockIdx.x * NT;
    const int rows = (int)min((long long)NT, P - row0);
    const Cam c = load_cam(intr, extr);
    __shared__ unsigned long long s_bar;
    {
        const SlabIn in[] = {{s_xyz, xyz, 3}, {s_cov, cov3d, 6}, {s_gc, dL_dconic, 3}};
        slabs_load<NT>(&s_bar, in, row0, rows);
    }
    const int t = threadIdx.x;
    float cam[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) cam[i] = 0.f;
    float dx = 0.f, dy = 0.f, dz = 0.f;
    float dcv[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (t < rows && radius[row0 + t] > 0) {  // ewa_project.cu:97
        const float2* cp = reinterpret_cast<const float2*>(s_cov) + 3 * t;
        const float2 c0 = cp[0], c1 = cp[1], c2 = cp[2];
        const float cv[6] = {c0.x, c0.y, c1.x, c1.y, c2.x, c2.y};
        if (!ewa_bwd<CAM>(c, s_xyz[3 * t], s_xyz[3 * t + 1], s_xyz[3 * t + 2], cv, s_gc[3 * t], s_gc[3 * t + 1],
                          s_gc[3 * t + 2], dx, dy, dz, dcv, cam)) {
            dx = dy = dz = 0.f;
#pragma unroll
            for (int i = 0; i < 6; ++i) dcv[i] = 0.f;
        }
    }
    __syncthreads();
    if (t < rows) {
        s_xyz[3 * t] = dx;
        s_xyz[3 * t + 1] = dy;
        s_xyz[3 * t + 2] = dz;
        float2* o = reinterpret_cast<float2*>(s_cov) + 3 * t;
        o[0] = make_float2(dcv[0], dcv[1]);
        o[1] = make_float2(dcv[2], dcv[3]);
        o[2] = make_float2(dcv[4], dcv[5]);
    }
    {
        const SlabOut out[] = {{s_xyz, dL_dxyz, 3}, {s_cov, dL_dcov3d, 6}};
        slabs_store<NT>(out, row0, rows);
    }
    if (CAM) cam_reduce_atomic(cam, dL_dintr, dL_dextr, s_red);
}

// ------------------------------------------------------------------------------------------
// Fused forward: project + (visible = depth != 0) + cov3d + ewa  (msplat/__init__.py:70-81)
// 40 B in, 32 B out per Gaussian; cov3d is never materialised.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) preprocess_fwd_kernel(
    int P, const float* __restrict__ xyz, const float* __restrict__ scale, const float* __restrict__ quat,
    const float* __restrict__ intr, const float* __restrict__ extr, int W, int H, float nearest, float extent,
    float* __restrict__ uv, float* __restrict__ depth, float* __restrict__ conic, int* __restrict__ radius,
    int* __restrict__ tiles) {
    __shared__ __align__(16) float s[NT * 10];
    float* s_xyz = s;
    float* s_scale = s + NT * 3;
    float* s_quat = s + NT * 6;
    const long long row0 = (long long)blockIdx.x * NT;
    const int rows = (int)min((long long)NT, P - row0);
    const int gx = (W + MSB_TILE - 1) / MSB_TILE, gy = (H + MSB_TILE - 1) / MSB_TILE;
    const Cam c = load_cam(intr, extr);
    __shared__ unsigned long long s_bar;
    {
        const SlabIn in[] = {{s_xyz, xyz, 3}, {s_scale, scale, 3}, {s_quat, quat, 4}};
        slabs_load<NT>(&s_bar, in, row0, rows);
    }
    const int t = threadIdx.x;
    float u = 0.f, v = 0.f, d = 0.f, cx = 0.f, cy = 0.f, cz = 0.f;
    int rad = 0, til = 0;
    if (t < rows) {
        const float px = s_xyz[3 * t], py = s_xyz[3 * t + 1], pz = s_xyz[3 * t + 2];
        if (!project_fwd(c, px, py, pz, W, H, nearest, extent, u, v, d)) u = v = d = 0.f;
        if (d != 0.f) {  // visible = depth != 0
            const float4 q = reinterpret_cast<const float4*>(s_quat)[t];
            float cv[6];
            cov3d_fwd(s_scale[3 * t], s_scale[3 * t + 1], s_scale[3 * t + 2], q.x, q.y, q.z, q.w, cv);
            if (!ewa_fwd(c, px, py, pz, cv, u, v, gx, gy, cx, cy, cz, rad, til)) {
                cx = cy = cz = 0.f;
                rad = til = 0;
            }
        }
    }
    __syncthreads();
    if (t < rows) {
        s[2 * t] = u;  // uv slab at s[0 .. 2*NT)
        s[2 * t + 1] = v;
        float* sc = s + NT * 2;  // conic slab at s[2*NT .. 5*NT)
        sc[3 * t] = cx;
        sc[3 * t + 1] = cy;
        sc[3 * t + 2] = cz;
        depth[row0 + t] = d;
        radius[row0 + t] = rad;
        tiles[row0 + t] = til;
    }
    {
        const SlabOut out[] = {{s, uv, 2}, {s + NT * 2, conic, 3}};
        slabs_store<NT>(out, row0, rows);
    }
}

// ------------------------------------------------------------------------------------------
// Fused backward: ewa bwd -> cov3d bwd, + project bwd; xyz gets both contributions.
// ------------------------------------------------------------------------------------------
template <bool CAM>
__global__ void __launch_bounds__(NT) preprocess_bwd_kernel(
    int P, const float* __restrict__ xyz, const float* __restrict__ scale, const float* __restrict__ quat,
    const float* __restrict__ intr, const float* __restrict__ extr, const float* __restrict__ depth,
    const int* __restrict__ radius, const float* __restrict__ dL_duv, const float* __restrict__ dL_ddepth,
    const float* __restrict__ dL_dconic, float* __restrict__ dL_dxyz, float* __restrict__ dL_dscale,
    float* __restrict__ dL_dquat, float* __restrict__ dL_dintr, float* __restrict__ dL_dextr) {
    __shared__ __align__(16) float s[NT * 15];
    __shared__ float s_red[8 * 16];
    float* s_xyz = s;
    float* s_scale = s + NT * 3;
    float* s_quat = s + NT * 6;
    float* s_guv = s + NT * 10;
    float* s_gc = s + NT * 12;
    const long long row0 = (long long)blockIdx.x * NT;
    const int rows = (int)min((long long)NT, P - row0);
    const Cam c = load_cam(intr, extr);
    __shared__ unsigned long long s_bar;
    {
        const SlabIn in[] = {{s_xyz, xyz, 3}, {s_scale, scale, 3}, {s_quat, quat, 4}, {s_guv, dL_duv, 2}, {s_gc, dL_dconic, 3}};
        slabs_load<NT>(&s_bar, in, row0, rows);
    }
    const int t = threadIdx.x;
    float cam[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) cam[i] = 0.f;
    float dx = 0.f, dy = 0.f, dz = 0.f;
    float ds[3] = {0.f, 0.f, 0.f}, dq[4] = {0.f, 0.f, 0.f, 0.f};
    if (t < rows) {
        const float d = depth[row0 + t];
        if (d != 0.f) {
            const float px = s_xyz[3 * t], py = s_xyz[3 * t + 1], pz = s_xyz[3 * t + 2];
            const float gd = dL_ddepth != nullptr ? dL_ddepth[row0 + t] : 0.f;
            project_bwd<CAM>(c, px, py, pz, s_guv[2 * t], s_guv[2 * t + 1], gd, dx, dy, dz, cam);
            if (radius[row0 + t] > 0) {
                const float4 q = reinterpret_cast<const float4*>(s_quat)[t];
                const float sx = s_scale[3 * t], sy = s_scale[3 * t + 1], sz = s_scale[3 * t + 2];
                float cv[6], dcv[6], ex, ey, ez;
                cov3d_fwd(sx, sy, sz, q.x, q.y, q.z, q.w, cv);
                if (ewa_bwd<CAM>(c, px, py, pz, cv, s_gc[3 * t], s_gc[3 * t + 1], s_gc[3 * t + 2], ex, ey, ez, dcv,
                                 cam)) {
                    dx += ex;
                    dy += ey;
                    dz += ez;
                    cov3d_bwd(sx, sy, sz, q.x, q.y, q.z, q.w, dcv, ds, dq);
                }
            }
        }
    }
    __syncthreads();
    if (t < rows) {
        s_xyz[3 * t] = dx;
        s_xyz[3 * t + 1] = dy;
        s_xyz[3 * t + 2] = dz;
        s_scale[3 * t] = ds[0];
        s_scale[3 * t + 1] = ds[1];
        s_scale[3 * t + 2] = ds[2];
        reinterpret_cast<float4*>(s_quat)[t] = make_float4(dq[0], dq[1], dq[2], dq[3]);
    }
    {
        const SlabOut out[] = {{s_xyz, dL_dxyz, 3}, {s_scale, dL_dscale, 3}, {s_quat, dL_dquat, 4}};
        slabs_store<NT>(out, row0, rows);
    }
    if (CAM) cam_reduce_atomic(cam, dL_dintr, dL_dextr, s_red);
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline unsigned nblocks(int P) { return (unsigned)((P + NT - 1) / NT); }

}  // namespace msb

using namespace msb;

#define MSB_REQUIRE(cond, msg) \
    if (!(cond)) return set_error(MSB_ERR_ARG, msg)

extern "C" {

int msb_project_point_fwd(const float* xyz, const float* intr, const float* extr, int P, int W, int H,
                          float nearest, float extent, float* uv, float* depth, void* stream) {
    if (P == 0) return MSB_OK;
    MSB_REQUIRE(P > 0 && xyz && intr && extr && uv && depth, "project_point_fwd: null pointer or negative P");
    MSB_REQUIRE(aligned16(xyz) && aligned16(uv), "project_point_fwd: xyz/uv must be 16-byte aligned");
    project_point_fwd_kernel<<<nblocks(P), NT, 0, (cudaStream_t)stream>>>(P, xyz, intr, extr, W, H, nearest, extent,
                                                                          uv, depth);
    return check_launch("project_point_fwd");
}

int msb_project_point_bwd(const float* xyz, const float* intr, const float* extr, const float* depth,
                          const float* dL_duv, const float* dL_ddepth, int P, float* dL_dxyz, float* dL_dintr,
                          float* dL_dextr, void* stream) {
    if (P == 0) return MSB_OK;
    MSB_REQUIRE(P > 0 && xyz && intr && extr && depth && dL_duv && dL_ddepth && dL_dxyz,
                "project_point_bwd: null pointer or negative P");
    MSB_REQUIRE(aligned16(xyz) && aligned16(dL_duv) && aligned16(dL_dxyz), "project_point_bwd: 16-byte alignment");
    if (dL_dintr || dL_dextr)
        project_point_bwd_kernel<true><<<nblocks(P), NT, 0, (cudaStream_t)stream>>>(
            P, xyz, intr, extr, depth, dL_duv, dL_ddepth, dL_dxyz, dL_dintr, dL_dextr);
    else
        project_point_bwd_kernel<false><<<nblocks(P), NT, 0, (cudaStream_t)stream>>>(
            P, xyz, intr, extr, depth, dL_duv, dL_ddepth, dL_dxyz, nullptr, nullptr);
    return check_launch("project_point_bwd");
}

int msb_compute_cov3d_fwd(const float* scale, const float* quat, const uint8_t* visible, int P, float* cov3d,
                          void* stream) {
    if (P == 0) return MSB_OK;
    MSB_REQUIRE(P > 0 && scale && quat && cov3d, "compute_cov3d_fwd: null pointer or negative P");
    MSB_REQUIRE(aligned16(scale) && aligned16(quat) && aligned16(cov3d), "compute_cov3d_fwd: 16-byte alignment");
    cov3d_fwd_kernel<<<nblocks(P), NT, 0, (cudaStream_t)stream>>>(P, scale, quat, visible, cov3d);
    return check_launch("compute_cov3d_fwd");
}

int msb_compute_cov3d_bwd(const float* scale, const float* quat, const uint8_t* visible, const float* dL_dcov3d,
                          int P, float* dL_dscale, float* dL_dquat, void* stream) {
    if (P == 0) return MSB_OK;
    MSB_REQUIRE(P > 0 && scale && quat && dL_dcov3d && dL_dscale && dL_dquat,
                "compute_cov3d_bwd: null pointer or negative P");
    MSB_REQUIRE(aligned16(scale) && aligned16(quat) && aligned16(dL_dcov3d) && aligned16(dL_dscale) &&
                    aligned16(dL_dquat),
                "compute_cov3d_bwd: 16-byte alignment");
    cov3d_bwd_kernel<<<nblocks(P), NT, 0, (cudaStream_t)stream>>>(P, scale, quat, visible, dL_dcov3d, dL_dscale,
                                                                  dL_dquat);
    return check_launch("compute_cov3d_bwd");
}

int msb_ewa_project_fwd(const float* xyz, const float* cov3d, const float* intr, const float* extr,
                        const float* uv, const uint8_t* visible, int P, int W, int H, float* conic,
                        int32_t* radius, int32_t* tiles, void* stream) {
    if (P == 0) return MSB_OK;
    MSB_REQUIRE(P > 0 && xyz && cov3d && intr && extr && uv && conic && radius && tiles,
                "ewa_project_fwd: null pointer or negative P");
    MSB_REQUIRE(aligned16(xyz) && aligned16(cov3d) && aligned16(uv) && aligned16(conic),
                "ewa_project_fwd: 16-byte alignment");
    const int gx = (W + MSB_TILE - 1) / MSB_TILE, gy = (H + MSB_TILE - 1) / MSB_TILE;
    ewa_fwd_kernel<<<nblocks(P), NT, 0, (cudaStream_t)stream>>>(P, xyz, cov3d, intr, extr, uv, visible, gx, gy,
                                                                conic, radius, tiles);
    return check_launch("ewa_project_fwd");
}

int msb_ewa_project_bwd(const float* xyz, const float* cov3d, const float* intr, const float* extr,
                        const int32_t* radius, const float* dL_dconic, int P, float* dL_dxyz, float* dL_dcov3d,
                        float* dL_dintr, float* dL_dextr, void* stream) {
    if (P == 0) return MSB_OK;
    MSB_REQUIRE(P > 0 && xyz && cov3d && intr && extr && radius && dL_dconic && dL_dxyz && dL_dcov3d,
                "ewa_project_bwd: null pointer or negative P");
    MSB_REQUIRE(aligned16(xyz) && aligned16(cov3d) && aligned16(dL_dconic) && aligned16(dL_dxyz) &&
                    aligned16(dL_dcov3d),
                "ewa_project_bwd: 16-byte alignment");
    if (dL_dintr || dL_dextr)
        ewa_bwd_kernel<true><<<nblocks(P), NT, 0, (cudaStream_t)stream>>>(P, xyz, cov3d, intr, extr, radius,
                                                                          dL_dconic, dL_dxyz, dL_dcov3d, dL_dintr,
                                                                          dL_dextr);
    else
        ewa_bwd_kernel<false><<<nblocks(P), NT, 0, (cudaStream_t)stream>>>(P, xyz, cov3d, intr, extr, radius,
                                                                           dL_dconic, dL_dxyz, dL_dcov3d, nullptr,
                                                                           nullptr);
    return check_launch("ewa_project_bwd");
}

int msb_preprocess_fwd(const float* xyz, const float* scale, const float* quat, const float* intr,
                       const float* extr, int P, int W, int H, float nearest, float extent, float* uv,
                       float* depth, float* conic, int32_t* radius, int32_t* tiles, void* stream) {
    if (P == 0) return MSB_OK;
    MSB_REQUIRE(P > 0 && xyz && scale && quat && intr && extr && uv && depth && conic && radius && tiles,
                "preprocess_fwd: null pointer or negative P");
    MSB_REQUIRE(aligned16(xyz) && aligned16(scale) && aligned16(quat) && aligned16(uv) && aligned16(conic),
                "preprocess_fwd: 16-byte alignment");
    preprocess_fwd_kernel<<<nblocks(P), NT, 0, (cudaStream_t)stream>>>(P, xyz, scale, quat, intr, extr, W, H,
                                                                       nearest, extent, uv, depth, conic, radius,
                                                                       tiles);
    return check_launch("preprocess_fwd");
}

int msb_preprocess_bwd(const float* xyz, const float* scale, const float* quat, const float* intr,
                       const float* extr, const float* depth, const int32_t* radius, const float* dL_duv,
                       const float* dL_ddepth, const float* dL_dconic, int P, float* dL_dxyz, float* dL_dscale,
                       float* dL_dquat, float* dL_dintr, float* dL_dextr, void* stream) {
    if (P == 0) return MSB_OK;
    MSB_REQUIRE(P > 0 && xyz && scale && quat && intr && extr && depth && radius && dL_duv && dL_dconic &&
                    dL_dxyz && dL_dscale && dL_dquat,
                "preprocess_bwd: null pointer or negative P");
    MSB_REQUIRE(aligned16(xyz) && aligned16(scale) && aligned16(quat) && aligned16(dL_duv) &&
                    aligned16(dL_dconic) && aligned16(dL_dxyz) && aligned16(dL_dscale) && aligned16(dL_dquat),
                "preprocess_bwd: 16-byte alignment");
    if (dL_dintr || dL_dextr)
        preprocess_bwd_kernel<true><<<nblocks(P), NT, 0, (cudaStream_t)stream>>>(
            P, xyz, scale, quat, intr, extr, depth, radius, dL_duv, dL_ddepth, dL_dconic, dL_dxyz, dL_dscale,
            dL_dquat, dL_dintr, dL_dextr);
    else
        preprocess_bwd_kernel<false><<<nblocks(P), NT, 0, (cudaStream_t)stream>>>(
            P, xyz, scale, quat, intr, extr, depth, radius, dL_duv, dL_ddepth, dL_dconic, dL_dxyz, dL_dscale,
            dL_dquat, nullptr, nullptr);
    return check_launch("preprocess_bwd");
}

}  // extern "C"
