// msplat_b200/csrc/render.cu -- fused per-Gaussian preprocess for the SH render path.
//
// One forward kernel does what the steps pipeline does in seven launches plus torch glue
//   project_point  (/root/reference/msplat/src/project_point.cu:13-57)
//   visible = depth != 0, compute_cov3d (src/compute_cov3d.cu:14-58)
//   ewa_project    (src/ewa_project.cu:16-83)
//   view dir = normalize(xyz - camera centre), compute_sh (src/compute_sh.cu:17-503,1600-1637)
//   colour = clamp_min(sh + 0.5, 0), feature = cat(colour, depth)   (tutorial glue)
//   + the blend kernels' record packing (blend.cu: blend_pack_kernel)
// and one backward kernel does the reverse (K8 + glue + K6 + K4 + K2), optionally ACCUMULATING
// into the gradient tensors so that a batch of views needs no separate add passes.
//
// The geometry uses the same device functions as the step kernels (geom.cuh), so uv, depth,
// conic, radius and tiles are bit-identical to the steps API and to the reference build.
//
// Layout / roofline (HBM-bound): a block owns G consecutive Gaussians (G = 256 for degree <= 4).
//   phase 1  one thread per Gaussian: [G,K] input slabs arrive as TMA bulk copies (cp.async.bulk +
//            mbarrier; per-thread 16-byte loads for the ragged last block), geometry + packed cull extents
//            + SH basis (once per Gaussian, into shared memory [D][G+1]); Gaussians that touch no tile
//            are dropped here: their SH rows are never read; the forward also sums the tile counts (M);
//   phase 2  groups of LPR lanes stream the surviving Gaussians' [Cs, D] coefficient rows with
//            16-byte loads (sh_layout.cuh), shuffle-reduce, write colours into a shared slab;
//   phase 3  (backward) one thread per Gaussian: dL_ddir -> dL_dxyz through the normalisation,
//            projection / EWA / cov3d backward; outputs leave through shared slabs as bulk stores, or as
//            bulk FP32 reduce-adds (cp.reduce.async.bulk) / red.global.v4 when accumulating over a view batch.
// Algorithmic bytes per Gaussian at SH3 RGB+depth: forward 44 + 192*v in, 64 out;
// backward 44 + 48 + 4 + 192*v in, 48 + 192 out (+ the same again when accumulating),
// v = fraction of Gaussians that touch a tile.
#include <stdlib.h>

#include "blend_math.cuh"
#include "geom.cuh"
#include "sh_eval.cuh"
#include "sh_layout.cuh"

namespace msb {

constexpr int RP_NT = 256;
// SH rows (channels) fetched together per lane group: 4 x 16 B in flight per lane (2 when a lane already holds >8 floats per row)
__host__ __device__ constexpr int rp_cu(int deg) { return sh_iters(deg) * (sh_vec(deg) ? 4 : 1) > 8 ? 2 : 4; }
__host__ __device__ constexpr int rp_gpb(int deg) { return deg <= 4 ? 256 : deg <= 6 ? 128 : 64; }
__host__ __device__ constexpr int rp_bs(int deg) { return (sh_dim(deg) * (rp_gpb(deg) + 1) + 3) / 4 * 4; }

constexpr unsigned RP_INVISIBLE = 0x80000000u;  // list flag: Gaussian touches no tile
constexpr unsigned RP_DEAD = 0x40000000u;       // list flag: no colour gradient arrived

struct CamCenter {
    float x, y, z;
};
// camera centre in world space: -R^T t
MSB_HD CamCenter cam_center(const Cam& c) {
    const float* e = c.e;
    CamCenter o;
    o.x = -(e[0] * e[3] + e[4] * e[7] + e[8] * e[11]);
    o.y = -(e[1] * e[3] + e[5] * e[7] + e[9] * e[11]);
    o.z = -(e[2] * e[3] + e[6] * e[7] + e[10] * e[11]);
    return o;
}

// block-wide append of the threads with `flag` set to s_list (order irrelevant)
__device__ __forceinline__ void list_append(bool flag, unsigned value, unsigned* s_list, int* s_cnt) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    int base = 0;
    if (lane == 0 && bal) base = atomicAdd(s_cnt, __popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (flag) s_list[base + __popc(bal & ((1u << lane) - 1u))] = value;
}

template <int NT>
__device__ __forceinline__ void slab_store_acc(float* __restrict__ g, const float* __restrict__ smem,
                                               long long first, int count) {
    float* dst = g + first;
    const int nvec = count >> 2;
    float4* v = reinterpret_cast<float4*>(dst);
    const float4* s = reinterpret_cast<const float4*>(smem);
    for (int i = threadIdx.x; i < nvec; i += NT) {
        float4 a = v[i];
        const float4 b = s[i];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        v[i] = a;
    }
    for (int i = 4 * nvec + threadIdx.x; i < count; i += NT) dst[i] += smem[i];
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
// One launch serves a whole VIEW BATCH: a block loads its parameter slabs once and then loops over the
// `views` cameras (intr [views,4], extr [views,estride]); view b writes its outputs at Gaussian offset
// b * vstride (rec / featp / uv / depth / radius / tiles are [views, vstride, ...], view-major -- the
// layout the batched sort and blend kernels index with view * P + index).  The block's SH rows come from
// HBM once; views 1.. find them in L2 (the block walks its views back to back).
template <int DEG>
__global__ void __launch_bounds__(RP_NT, DEG <= 4 ? 5 : 2) render_pre_fwd_kernel(
    int P, int Cs, int Cpad, int with_depth, int views, long long vstride, const float* __restrict__ xyz,
    const float* __restrict__ scale, const float* __restrict__ quat, const float* __restrict__ opacity,
    const float* __restrict__ shs, const float* __restrict__ intr, const float* __restrict__ extr, int estride, int W,
    int H, float nearest, float extent, float sh_bias, int clamp, float* __restrict__ rec, float* __restrict__ featp,
    float* __restrict__ uv, float* __restrict__ depth, int* __restrict__ radius, int* __restrict__ tiles,
    unsigned long long* __restrict__ total_tiles) {
    constexpr int D = sh_dim(DEG);
    constexpr int G = rp_gpb(DEG);
    constexpr int GS = G + 1;
    constexpr bool VEC = sh_vec(DEG);
    constexpr int LPR = sh_lpr(DEG);
    constexpr int IT = sh_iters(DEG);
    constexpr int UNITS = sh_units(DEG);
    constexpr int WD = VEC ? 4 : 1;
    constexpr int GROUPS = RP_NT / LPR;
    constexpr int CU = rp_cu(DEG);
    extern __shared__ __align__(16) float sm[];
    float* s_xyz = sm;             // [G,3]   inputs: loaded once, read by every view
    float* s_scale = sm + 3 * G;   // [G,3]
    float* s_quat = sm + 6 * G;    // [G,4]
    float* s_op = sm + 10 * G;     // [G]
    float* s_rec = sm + 11 * G;    // [G,8]   outputs of the current view
    float* s_uv = sm + 19 * G;     // [G,2]
    float* s_B = sm + 21 * G;      // [D][GS]
    float* s_feat = s_B + rp_bs(DEG);                              // [G,Cpad]
    unsigned* s_list = reinterpret_cast<unsigned*>(s_feat + (size_t)Cpad * G);  // [G]
    __shared__ int s_cnt;
    __shared__ unsigned long long s_bar;
    __shared__ unsigned int s_tsum;

    const int tid = threadIdx.x;
    const long long g0 = (long long)blockIdx.x * G;
    const int rows = (int)min((long long)G, (long long)P - g0);
    const int gx = (W + MSB_TILE - 1) / MSB_TILE, gy = (H + MSB_TILE - 1) / MSB_TILE;
    // full blocks: the four input slabs arrive as TMA bulk copies issued by one thread (3 + 3 + 4 + 1 KB
    // at G = 256), everyone waits on the mbarrier; the ragged last block uses the per-thread path
    const bool full = rows == G;
    if (tid == 0 && full) mbar_init(&s_bar, 1);
    __syncthreads();
    if (full) {
        if (tid == 0) {
            mbar_expect_tx(&s_bar, (unsigned)(11 * G * sizeof(float)));
            bulk_g2s(s_xyz, xyz + g0 * 3, 3 * G * sizeof(float), &s_bar);
            bulk_g2s(s_scale, scale + g0 * 3, 3 * G * sizeof(float), &s_bar);
            bulk_g2s(s_quat, quat + g0 * 4, 4 * G * sizeof(float), &s_bar);
            bulk_g2s(s_op, opacity + g0, G * sizeof(float), &s_bar);
        }
        mbar_wait(&s_bar, 0);
    } else {
        slab_load<RP_NT>(s_xyz, xyz, g0 * 3, rows * 3);
        slab_load<RP_NT>(s_scale, scale, g0 * 3, rows * 3);
        slab_load<RP_NT>(s_quat, quat, g0 * 4, rows * 4);
        slab_load<RP_NT>(s_op, opacity, g0, rows);
        __syncthreads();
    }
    const int t = tid;
    float px = 0.f, py = 0.f, pz = 0.f, op = 0.f;
    float cv[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // cov3d does not depend on the camera: once per Gaussian
    if (t < rows) {
        px = s_xyz[3 * t];
        py = s_xyz[3 * t + 1];
        pz = s_xyz[3 * t + 2];
        op = s_op[t];
        const float4 q = reinterpret_cast<const float4*>(s_quat)[t];
        cov3d_fwd(s_scale[3 * t], s_scale[3 * t + 1], s_scale[3 * t + 2], q.x, q.y, q.z, q.w, cv);
    }

    for (int b = 0; b < views; ++b) {
        const Cam c = load_cam(intr + 4 * b, extr + (size_t)estride * b);
        const CamCenter cc = cam_center(c);
        const long long v0 = (long long)b * vstride + g0;  // first output row of this block in view b
        if (tid == 0) {
            s_cnt = 0;
            s_tsum = 0;
        }
        __syncthreads();  // also: the previous view's bulk stores have read the output slabs (thread 0 waited)

        // ---- phase 1: geometry + basis, one thread per Gaussian -------------------------------------
        float u = 0.f, v = 0.f, d = 0.f, cx = 0.f, cy = 0.f, cz = 0.f, hx = 0.f, hy = 0.f;
        int rad = 0, til = 0;
        if (t < rows) {
            if (!project_fwd(c, px, py, pz, W, H, nearest, extent, u, v, d)) u = v = d = 0.f;
            if (d != 0.f) {  // visible = depth != 0 (msplat/__init__.py:73)
                if (!ewa_fwd(c, px, py, pz, cv, u, v, gx, gy, cx, cy, cz, rad, til)) {
                    cx = cy = cz = 0.f;
                    rad = til = 0;
                }
            }
            if (til > 0) {
                float ex, ey, es, et;
                cull_extent(cx, cy, cz, op, ex, ey, es, et);
                cull_pack(ex, ey, es, et, hx, hy);  // hx, hy now hold the packed FP16 pairs of the blend record
                const float rx = px - cc.x, ry = py - cc.y, rz = pz - cc.z;
                const float inv = 1.0f / sqrtf(rx * rx + ry * ry + rz * rz);
                sh_basis<DEG>(rx * inv, ry * inv, rz * inv, s_B + t, GS);
            }
        }
        list_append(til > 0, (unsigned)t, s_list, &s_cnt);
        if (total_tiles != nullptr) {  // M = sum(tiles): sizes the sort output (replaces the separate count pass)
            unsigned wsum = (unsigned)til;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
            if ((tid & 31) == 0 && wsum) atomicAdd(&s_tsum, wsum);
        }
        if (t < rows) {
            float4* r = reinterpret_cast<float4*>(s_rec) + 2 * t;
            r[0] = make_float4(u, v, cx, cy);
            r[1] = make_float4(cz, op, hx, hy);
            s_uv[2 * t] = u;
            s_uv[2 * t + 1] = v;
            float* f = s_feat + (size_t)t * Cpad;
            for (int k = 0; k < Cpad; ++k) f[k] = 0.f;
            if (with_depth) f[Cs] = d;
            depth[v0 + t] = d;
            radius[v0 + t] = rad;
            tiles[v0 + t] = til;
        }
        __syncthreads();
        if (total_tiles != nullptr && tid == 0 && s_tsum) atomicAdd(total_tiles + b, (unsigned long long)s_tsum);  // one per CTA

        // ---- phase 2: SH rows of the surviving Gaussians, LPR lanes per Gaussian ---------------------
        const int cnt = s_cnt;
        const int grp = tid / LPR, s = tid % LPR;
        for (int k0 = 0; k0 < cnt; k0 += GROUPS) {  // uniform trip count across the block
            const int k = k0 + grp;
            const bool act = k < cnt;
            const int gl = act ? (int)s_list[k] : 0;
            float bs[IT * WD];
#pragma unroll
            for (int it = 0; it < IT; ++it) {
                const int un = s + it * LPR;
#pragma unroll
                for (int j = 0; j < WD; ++j) bs[it * WD + j] = (act && un < UNITS) ? s_B[(un * WD + j) * GS + gl] : 0.f;
            }
            const long long row0 = (g0 + gl) * Cs;
            // rows are fetched CU channels at a time so that every lane keeps CU * IT 16-byte loads
            // in flight (a load -> fma -> shuffle chain per channel is latency-bound)
            for (int c0 = 0; c0 < Cs; c0 += CU) {
                float xv[CU][IT * WD];
#pragma unroll
                for (int cc2 = 0; cc2 < CU; ++cc2) {
                    const bool chv = act && (c0 + cc2 < Cs);
                    const float* rp = shs + (row0 + c0 + cc2) * D;
#pragma unroll
                    for (int it = 0; it < IT; ++it) {
                        const int un = s + it * LPR;
                        if (VEC) {
                            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (chv && un < UNITS) x = ldg_stream4(reinterpret_cast<const float4*>(rp) + un);
                            xv[cc2][it * WD + 0] = x.x;
                            xv[cc2][it * WD + (WD > 1 ? 1 : 0)] = x.y;
                            xv[cc2][it * WD + (WD > 2 ? 2 : 0)] = x.z;
                            xv[cc2][it * WD + (WD > 3 ? 3 : 0)] = x.w;
                        } else {
                            xv[cc2][it * WD] = (chv && un < UNITS) ? __ldg(rp + un) : 0.f;
                        }
                    }
                }
#pragma unroll
                for (int cc2 = 0; cc2 < CU; ++cc2) {
                    float acc = 0.f;
#pragma unroll
                    for (int i = 0; i < IT * WD; ++i) acc = fmaf(xv[cc2][i], bs[i], acc);
                    acc = group_sum<LPR>(acc);
                    if (act && s == 0 && c0 + cc2 < Cs) {
                        float val = acc + sh_bias;
                        if (clamp) val = fmaxf(val, 0.f);
                        s_feat[(size_t)gl * Cpad + c0 + cc2] = val;
                    }
                }
            }
        }
        if (full) {
            fence_async_smem();  // this thread's slab writes -> visible to the bulk stores below
            __syncthreads();
            if (tid == 0) {
                bulk_s2g(rec + v0 * 8, s_rec, 8 * G * sizeof(float));
                bulk_s2g(uv + v0 * 2, s_uv, 2 * G * sizeof(float));
                bulk_s2g(featp + v0 * Cpad, s_feat, (unsigned)((size_t)Cpad * G * sizeof(float)));
                bulk_commit();
                bulk_wait_read();  // shared memory must stay valid until the copy engine has read it
            }
        } else {
            __syncthreads();
            slab_store<RP_NT>(rec, s_rec, v0 * 8, rows * 8);
            slab_store<RP_NT>(uv, s_uv, v0 * 2, rows * 2);
            slab_store<RP_NT>(featp, s_feat, v0 * Cpad, rows * Cpad);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
// One launch serves a whole view batch (see the forward kernel): the parameter slabs are loaded once, the
// packed blend gradients of view b + 1 (grec / gfeat [views, vstride, ...]) are prefetched by TMA while
// view b is processed, the geometry gradients of all views are summed in registers and leave once, and
// dL_dshs rows are written by view 0 and reduced into (red.global.add.v4 at the L2, where the row still
// sits) by the later views.
template <int DEG, bool CAM>
__global__ void __launch_bounds__(RP_NT, DEG <= 4 ? 4 : 2) render_pre_bwd_kernel(
    int P, int Cs, int Cpad, int with_depth, int accumulate, int views, long long vstride,
    const float* __restrict__ xyz, const float* __restrict__ scale, const float* __restrict__ quat,
    const float* __restrict__ shs, const float* __restrict__ intr, const float* __restrict__ extr, int estride,
    float sh_bias, int clamp, const int* __restrict__ tiles, const float* __restrict__ grec,
    const float* __restrict__ gfeat, float* __restrict__ dL_dxyz,
    float* __restrict__ dL_dscale, float* __restrict__ dL_dquat, float* __restrict__ dL_dopacity,
    float* __restrict__ dL_dshs, float* __restrict__ dL_dintr, float* __restrict__ dL_dextr) {
    constexpr int D = sh_dim(DEG);
    constexpr int G = rp_gpb(DEG);
    constexpr int GS = G + 1;
    constexpr bool VEC = sh_vec(DEG);
    constexpr int LPR = sh_lpr(DEG);
    constexpr int IT = sh_iters(DEG);
    constexpr int UNITS = sh_units(DEG);
    constexpr int WD = VEC ? 4 : 1;
    constexpr int GROUPS = RP_NT / LPR;
    constexpr int CU = rp_cu(DEG);
    extern __shared__ __align__(16) float sm[];
    float* s_xyz = sm;             // [G,3]  -> dL_dxyz slab after the last view
    float* s_scale = sm + 3 * G;   // [G,3]  -> dL_dscale slab
    float* s_quat = sm + 6 * G;    // [G,4]  -> dL_dquat slab
    float* s_gin = sm + 10 * G;    // [2][G, 8 + Cpad]  packed gradients of the current / next view
    const int gin_stride = (8 + Cpad) * G;
    float* s_B = s_gin + 2 * gin_stride;  // [D][GS]
    float* s_W = s_B;              // [D][GS]  aliases s_B: column gl is loaded into registers by the lane group
                                   //          that owns Gaussian gl before the same group overwrites it with w
    unsigned* s_list = reinterpret_cast<unsigned*>(s_B + rp_bs(DEG));   // [G]
    __shared__ int s_cnt;
    __shared__ float s_red[8 * 16];
    __shared__ unsigned long long s_bar[3];  // inputs | gradient buffer 0 | gradient buffer 1

    const int tid = threadIdx.x;
    const long long g0 = (long long)blockIdx.x * G;
    const int rows = (int)min((long long)G, (long long)P - g0);
    // write mode: the all-zero dL_dshs rows of Gaussians without a colour gradient are filled by bulk stores
    // from this zero block (one instruction per 2 KB instead of a lane group walking the row)
    // (compiled in for degree >= 5 only: for short rows the lane groups are as fast)
    constexpr bool ZF = DEG >= 5;
    constexpr int ZB = ZF ? 512 : 4;
    __shared__ __align__(16) float s_zero[ZB];
    const size_t row_bytes = (size_t)Cs * D * sizeof(float);
    const bool zero_fill = ZF && !accumulate && (row_bytes % 16) == 0 && row_bytes >= 1024;
    if constexpr (ZF) {
        for (int i = tid; i < ZB; i += RP_NT) s_zero[i] = 0.f;
        fence_async_smem();
    }
    const bool full = rows == G;  // full blocks move their slabs with TMA bulk copies (see the forward kernel)
    if (tid == 0 && full) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_init(&s_bar[2], 1);
    }
    __syncthreads();
    const unsigned gin_bytes = (unsigned)((size_t)(8 + Cpad) * G * sizeof(float));
    auto prefetch = [&](int b) {  // thread 0 of a full block: packed gradients of view b -> buffer b & 1
        float* dst = s_gin + (b & 1) * gin_stride;
        const long long v0 = (long long)b * vstride + g0;
        mbar_expect_tx(&s_bar[1 + (b & 1)], gin_bytes);
        bulk_g2s(dst, grec + v0 * 8, 8 * G * sizeof(float), &s_bar[1 + (b & 1)]);
        bulk_g2s(dst + 8 * G, gfeat + v0 * Cpad, (unsigned)((size_t)Cpad * G * sizeof(float)), &s_bar[1 + (b & 1)]);
    };
    if (full) {
        if (tid == 0) {
            mbar_expect_tx(&s_bar[0], (unsigned)(10 * G * sizeof(float)));
            bulk_g2s(s_xyz, xyz + g0 * 3, 3 * G * sizeof(float), &s_bar[0]);
            bulk_g2s(s_scale, scale + g0 * 3, 3 * G * sizeof(float), &s_bar[0]);
            bulk_g2s(s_quat, quat + g0 * 4, 4 * G * sizeof(float), &s_bar[0]);
            prefetch(0);
        }
        mbar_wait(&s_bar[0], 0);
    } else {
        slab_load<RP_NT>(s_xyz, xyz, g0 * 3, rows * 3);
        slab_load<RP_NT>(s_scale, scale, g0 * 3, rows * 3);
        slab_load<RP_NT>(s_quat, quat, g0 * 4, rows * 4);
        __syncthreads();
    }

    const int t = tid;
    // geometry gradients, summed over the views in registers
    float dx = 0.f, dy = 0.f, dz = 0.f, dop = 0.f;
    float ds[3] = {0.f, 0.f, 0.f}, dq[4] = {0.f, 0.f, 0.f, 0.f};
    bool filled = false;

    for (int b = 0; b < views; ++b) {
        const Cam c = load_cam(intr + 4 * b, extr + (size_t)estride * b);
        const CamCenter cc = cam_center(c);
        const long long v0 = (long long)b * vstride + g0;
        float* s_grec = s_gin + (b & 1) * gin_stride;   // [G,8]
        float* s_gfeat = s_grec + 8 * G;                // [G,Cpad]
        const bool acc_b = accumulate || b > 0;
        if (tid == 0) s_cnt = 0;
        if (full) {
            // every thread is past its reads of the other buffer (barrier at the end of the previous view)
            if (tid == 0 && b + 1 < views) prefetch(b + 1);
            mbar_wait(&s_bar[1 + (b & 1)], (unsigned)((b >> 1) & 1));
        } else {
            slab_load<RP_NT>(s_grec, grec, v0 * 8, rows * 8);
            slab_load<RP_NT>(s_gfeat, gfeat, v0 * Cpad, rows * Cpad);
        }
        __syncthreads();

        // ---- phase 1: basis of the Gaussians that received a colour gradient -------------------------
        bool vis = false, live = false;
        if (t < rows) {
            vis = tiles[v0 + t] > 0;
            if (vis) {
                const float* gf = s_gfeat + (size_t)t * Cpad;
                for (int k = 0; k < Cs; ++k) live = live || (gf[k] != 0.f);
                if (live) {
                    const float rx = s_xyz[3 * t] - cc.x, ry = s_xyz[3 * t + 1] - cc.y, rz = s_xyz[3 * t + 2] - cc.z;
                    const float inv = 1.0f / sqrtf(rx * rx + ry * ry + rz * rz);
                    sh_basis<DEG>(rx * inv, ry * inv, rz * inv, s_B + t, GS);
                }
            }
        }
        // write mode (first view of a launch that does not accumulate): every row of dL_dshs must be produced
        // (zeros for untouched Gaussians); otherwise only rows that actually change are touched
        const bool zf_b = zero_fill && !acc_b;
        const bool listed = (acc_b || zf_b) ? live : (t < rows);
        list_append(listed, (unsigned)t | (vis ? 0u : RP_INVISIBLE) | (live ? 0u : RP_DEAD), s_list, &s_cnt);
        if (ZF && zf_b && t < rows && !live) {
            char* row = reinterpret_cast<char*>(dL_dshs) + (size_t)(g0 + t) * row_bytes;
            for (size_t off = 0; off < row_bytes; off += ZB * sizeof(float))
                bulk_s2g(row + off, s_zero, (unsigned)min((size_t)(ZB * sizeof(float)), row_bytes - off));
            bulk_commit();
            filled = true;
        }
        __syncthreads();

        // ---- phase 2: dL_dshs rows + w_d = sum_c dL_dvalue_c * shs[c, d] ------------------------------
        const int cnt = s_cnt;
        const int grp = tid / LPR, s = tid % LPR;
        for (int k0 = 0; k0 < cnt; k0 += GROUPS) {
            const int k = k0 + grp;
            const bool act = k < cnt;
            const unsigned ent = act ? s_list[k] : (RP_INVISIBLE | RP_DEAD);
            const int gl = (int)(ent & 0x3fffffffu);
            const bool lv = act && !(ent & RP_DEAD);
            float bs[IT * WD], wacc[IT * WD];
#pragma unroll
            for (int it = 0; it < IT; ++it) {
                const int un = s + it * LPR;
#pragma unroll
                for (int j = 0; j < WD; ++j) {
                    bs[it * WD + j] = (lv && un < UNITS) ? s_B[(un * WD + j) * GS + gl] : 0.f;
                    wacc[it * WD + j] = 0.f;
                }
            }
            const long long row0 = (g0 + gl) * Cs;
            for (int c0 = 0; c0 < Cs; c0 += CU) {
                // fetch CU coefficient rows at once (memory-level parallelism), then consume them
                float sv[CU][IT * WD];
#pragma unroll
                for (int cc2 = 0; cc2 < CU; ++cc2) {
                    const bool chv = lv && (c0 + cc2 < Cs);
                    const float* rp = shs + (row0 + c0 + cc2) * D;
#pragma unroll
                    for (int it = 0; it < IT; ++it) {
                        const int un = s + it * LPR;
                        if (VEC) {
                            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (chv && un < UNITS) x = ldg_stream4(reinterpret_cast<const float4*>(rp) + un);
                            sv[cc2][it * WD + 0] = x.x;
                            sv[cc2][it * WD + (WD > 1 ? 1 : 0)] = x.y;
                            sv[cc2][it * WD + (WD > 2 ? 2 : 0)] = x.z;
                            sv[cc2][it * WD + (WD > 3 ? 3 : 0)] = x.w;
                        } else {
                            sv[cc2][it * WD] = (chv && un < UNITS) ? __ldg(rp + un) : 0.f;
                        }
                    }
                }
#pragma unroll
                for (int cc2 = 0; cc2 < CU; ++cc2) {
                    const int ch = c0 + cc2;
                    if (ch >= Cs) break;  // uniform across the block
                    float* op = dL_dshs + (row0 + ch) * D;
                    const float gv = lv ? s_gfeat[(size_t)gl * Cpad + ch] : 0.f;
                    float acc = 0.f;
#pragma unroll
                    for (int i = 0; i < IT * WD; ++i) acc = fmaf(sv[cc2][i], bs[i], acc);
                    // same arithmetic as the forward pass -> same clamp decision (clamp_min passes x >= 0)
                    acc = group_sum<LPR>(acc);
                    const float dv = (clamp && !(acc + sh_bias >= 0.f)) ? 0.f : gv;
#pragma unroll
                    for (int it = 0; it < IT; ++it) {
                        const int un = s + it * LPR;
                        if (!(act && un < UNITS)) continue;
                        if (VEC) {
                            float4 o = make_float4(bs[it * WD] * dv, bs[it * WD + (WD > 1 ? 1 : 0)] * dv,
                                                   bs[it * WD + (WD > 2 ? 2 : 0)] * dv, bs[it * WD + (WD > 3 ? 3 : 0)] * dv);
                            float4* q = reinterpret_cast<float4*>(op) + un;
                            if (acc_b) {
                                // fire-and-forget vector reduction: no read of the old row, no load latency
                                if (dv != 0.f) red_add_v4(reinterpret_cast<float*>(q), o.x, o.y, o.z, o.w);
                            } else {
                                *q = o;
                            }
                        } else {
                            const float o = bs[it * WD] * dv;
                            if (acc_b) {
                                if (dv != 0.f) atomicAdd(op + un, o);  // result unused -> RED
                            } else {
                                op[un] = o;
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < IT * WD; ++i) wacc[i] = fmaf(sv[cc2][i], dv, wacc[i]);
                }
            }
            if (lv) {
#pragma unroll
                for (int it = 0; it < IT; ++it) {
                    const int un = s + it * LPR;
#pragma unroll
                    for (int j = 0; j < WD; ++j)
                        if (un < UNITS) s_W[(un * WD + j) * GS + gl] = wacc[it * WD + j];
                }
            }
        }
        __syncthreads();

        // ---- phase 3: geometry backward, one thread per Gaussian ------------------------------------
        float cam[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) cam[i] = 0.f;
        if (t < rows && vis) {
            const float px = s_xyz[3 * t], py = s_xyz[3 * t + 1], pz = s_xyz[3 * t + 2];
            const float4 ga = reinterpret_cast<const float4*>(s_grec)[2 * t];      // dL_duv, dL_dconic.xy
            const float4 gb = reinterpret_cast<const float4*>(s_grec)[2 * t + 1];  // dL_dconic.z, dL_dopacity
            dop += gb.y;
            const float gd = with_depth ? s_gfeat[(size_t)t * Cpad + Cs] : 0.f;
            float ax, ay, az;
            project_bwd<CAM>(c, px, py, pz, ga.x, ga.y, gd, ax, ay, az, cam);
            const float4 q = reinterpret_cast<const float4*>(s_quat)[t];
            const float sx = s_scale[3 * t], sy = s_scale[3 * t + 1], sz = s_scale[3 * t + 2];
            float cv[6], dcv[6], ex, ey, ez;
            cov3d_fwd(sx, sy, sz, q.x, q.y, q.z, q.w, cv);
            if (ewa_bwd<CAM>(c, px, py, pz, cv, ga.z, ga.w, gb.x, ex, ey, ez, dcv, cam)) {
                ax += ex;
                ay += ey;
                az += ez;
                float vs[3] = {0.f, 0.f, 0.f}, vq[4] = {0.f, 0.f, 0.f, 0.f};
                cov3d_bwd(sx, sy, sz, q.x, q.y, q.z, q.w, dcv, vs, vq);
#pragma unroll
                for (int i = 0; i < 3; ++i) ds[i] += vs[i];
#pragma unroll
                for (int i = 0; i < 4; ++i) dq[i] += vq[i];
            }
            if (live) {
                const float rx = px - cc.x, ry = py - cc.y, rz = pz - cc.z;
                const float inv = 1.0f / sqrtf(rx * rx + ry * ry + rz * rz);
                const float dirx = rx * inv, diry = ry * inv, dirz = rz * inv;
                float hx, hy, hz;  // dL_ddir
                sh_basis_grad<DEG>(dirx, diry, dirz, s_W + t, GS, hx, hy, hz);
                // dir = r / |r|  =>  dL_dr = (g - dir (dir . g)) / |r|
                const float dt = dirx * hx + diry * hy + dirz * hz;
                const float grx = (hx - dirx * dt) * inv, gry = (hy - diry * dt) * inv, grz = (hz - dirz * dt) * inv;
                ax += grx;
                ay += gry;
                az += grz;
                if (CAM) {
                    // r = p - centre, centre = -R^T t  =>  dL_dR[i][j] += dL_dr[j] t[i], dL_dt[i] += R[i][:] . dL_dr
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const float ti = c.e[4 * i + 3];
                        cam[4 + 4 * i + 0] += grx * ti;
                        cam[4 + 4 * i + 1] += gry * ti;
                        cam[4 + 4 * i + 2] += grz * ti;
                        cam[4 + 4 * i + 3] += c.e[4 * i] * grx + c.e[4 * i + 1] * gry + c.e[4 * i + 2] * grz;
                    }
                }
            }
            dx += ax;
            dy += ay;
            dz += az;
        }
        if (CAM) cam_reduce_atomic<RP_NT>(cam, dL_dintr ? dL_dintr + 4 * b : nullptr,
                                          dL_dextr ? dL_dextr + (size_t)estride * b : nullptr, s_red);
        __syncthreads();  // this view's gradient buffer, basis and list may now be overwritten
    }

    if (t < rows) {
        s_xyz[3 * t] = dx;
        s_xyz[3 * t + 1] = dy;
        s_xyz[3 * t + 2] = dz;
        s_scale[3 * t] = ds[0];
        s_scale[3 * t + 1] = ds[1];
        s_scale[3 * t + 2] = ds[2];
        reinterpret_cast<float4*>(s_quat)[t] = make_float4(dq[0], dq[1], dq[2], dq[3]);
        if (accumulate) {
            if (dop != 0.f) atomicAdd(dL_dopacity + g0 + t, dop);  // result unused -> RED
        } else {
            dL_dopacity[g0 + t] = dop;
        }
    }
    if (full) {
        // gradient slabs leave as bulk stores, or as bulk FP32 reduce-adds when accumulating into the sums of
        // earlier launches (cp.reduce.async.bulk: the read-modify-write happens at the L2, not in this SM)
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            if (accumulate) {
                bulk_s2g_add_f32(dL_dxyz + g0 * 3, s_xyz, 3 * G * sizeof(float));
                bulk_s2g_add_f32(dL_dscale + g0 * 3, s_scale, 3 * G * sizeof(float));
                bulk_s2g_add_f32(dL_dquat + g0 * 4, s_quat, 4 * G * sizeof(float));
            } else {
                bulk_s2g(dL_dxyz + g0 * 3, s_xyz, 3 * G * sizeof(float));
                bulk_s2g(dL_dscale + g0 * 3, s_scale, 3 * G * sizeof(float));
                bulk_s2g(dL_dquat + g0 * 4, s_quat, 4 * G * sizeof(float));
            }
            bulk_commit();
        }
    } else {
        __syncthreads();
        if (accumulate) {
            slab_store_acc<RP_NT>(dL_dxyz, s_xyz, g0 * 3, rows * 3);
            slab_store_acc<RP_NT>(dL_dscale, s_scale, g0 * 3, rows * 3);
            slab_store_acc<RP_NT>(dL_dquat, s_quat, g0 * 4, rows * 4);
        } else {
            slab_store<RP_NT>(dL_dxyz, s_xyz, g0 * 3, rows * 3);
            slab_store<RP_NT>(dL_dscale, s_scale, g0 * 3, rows * 3);
            slab_store<RP_NT>(dL_dquat, s_quat, g0 * 4, rows * 4);
        }
    }
    if ((full && tid == 0) || filled) bulk_wait_read();  // shared memory stays valid until the copy engine has read it
}

// ------------------------------------------------------------------------------------------------
// "PT" variants for short coefficient rows (degree 1 and 3, Cs * D <= 64 floats): the block's SH rows are
// staged ONCE into shared memory (padded to an odd number of 16-byte units per row, so that 32 threads
// reading their own rows with LDS.128 hit distinct banks) and reused by all views; the thread that owns a
// Gaussian keeps the basis in registers and evaluates its colours itself.  Compared with the lane-group path
// above this removes the per-view global row loads (the latency the block stalled on), the basis round trip
// through shared memory, the survivor list and the shuffles: ~25 % fewer instructions per Gaussian and view, and
// the rows come from HBM once per block instead of once per view.
// ------------------------------------------------------------------------------------------------
__host__ __device__ constexpr int rp_pt_rs(int row_floats) { return ((row_floats / 4) | 1) * 4; }  // padded row stride
constexpr int RP_PT_MAX_ROW = 64;  // floats per Gaussian (Cs * D) the PT kernels stage

// stage `rows` coefficient rows of `rowf` floats (multiple of 4) into s_sh with row stride RS
__device__ __forceinline__ void pt_stage_rows(float* __restrict__ s_sh, const float* __restrict__ shs, long long g0,
                                              int rows, int rowf, int RS) {
    const int units = rowf >> 2;
    const float* src = shs + g0 * rowf;
    for (int i = threadIdx.x; i < rows * units; i += RP_NT) {
        const int r = i / units, u = i - r * units;
        cp_async16(s_sh + r * RS + 4 * u, src + (long long)r * rowf + 4 * u);
    }
    cp_async_commit();
}

// shared-memory floats of the forward PT kernel in front of the staged rows: the input slabs (11 G) are dead once
// every thread holds its Gaussian in registers, so the per-view output slabs (8 + 2 + Cpad floats per Gaussian)
// overlay them
__host__ __device__ constexpr int rp_pt_fwd_io(int Cpad) { return (10 + Cpad) > 11 ? (10 + Cpad) : 11; }

// Thread 0 moves the CTA's tile count of one view into the per-view total and re-arms the counter.  An atomic exchange
// (not a plain load + store): the compiler may not hoist it to the other threads, so no thread but 0 ever touches the
// word between the warps' atomicAdds of two views (racecheck-clean; a speculated plain load was flagged).
__device__ __forceinline__ void pt_drain_tiles(unsigned* s_tsum, unsigned long long* total_tiles, int b) {
    const unsigned s = atomicExch(s_tsum, 0u);
    if (total_tiles != nullptr && s) atomicAdd(total_tiles + b, (unsigned long long)s);
}

template <int DEG>
__global__ void __launch_bounds__(RP_NT, 3) render_pre_fwd_pt_kernel(
    int P, int Cs, int Cpad, int with_depth, int views, long long vstride, const float* __restrict__ xyz,
    const float* __restrict__ scale, const float* __restrict__ quat, const float* __restrict__ opacity,
    const float* __restrict__ shs, const float* __restrict__ intr, const float* __restrict__ extr, int estride, int W,
    int H, float nearest, float extent, float sh_bias, int clamp, float* __restrict__ rec, float* __restrict__ featp,
    float* __restrict__ uv, float* __restrict__ depth, int* __restrict__ radius, int* __restrict__ tiles,
    unsigned long long* __restrict__ total_tiles) {
    constexpr int D = sh_dim(DEG);
    constexpr int G = RP_NT;
    static_assert(D % 4 == 0, "PT kernels need 16-byte coefficient units");
    extern __shared__ __align__(16) float sm[];
    float* s_xyz = sm;             // [G,3]   inputs: read once, into registers
    float* s_scale = sm + 3 * G;   // [G,3]
    float* s_quat = sm + 6 * G;    // [G,4]
    float* s_op = sm + 10 * G;     // [G]
    float* s_rec = sm;             // [G,8]   outputs of the current view (overlay the dead inputs)
    float* s_uv = sm + 8 * G;      // [G,2]
    float* s_feat = sm + 10 * G;   // [G,Cpad]
    float* s_sh = sm + (size_t)rp_pt_fwd_io(Cpad) * G;  // [G, RS] coefficient rows, staged once
    __shared__ unsigned long long s_bar;
    __shared__ unsigned int s_tsum;
    const int rowf = Cs * D, RS = rp_pt_rs(rowf);

    const int tid = threadIdx.x;
    const long long g0 = (long long)blockIdx.x * G;
    const int rows = (int)min((long long)G, (long long)P - g0);
    const int gx = (W + MSB_TILE - 1) / MSB_TILE, gy = (H + MSB_TILE - 1) / MSB_TILE;
    const bool full = rows == G;
    if (tid == 0 && full) mbar_init(&s_bar, 1);
    if (tid == 0) s_tsum = 0;
    __syncthreads();
    pt_stage_rows(s_sh, shs, g0, rows, rowf, RS);
    if (full) {
        if (tid == 0) {
            mbar_expect_tx(&s_bar, (unsigned)(11 * G * sizeof(float)));
            bulk_g2s(s_xyz, xyz + g0 * 3, 3 * G * sizeof(float), &s_bar);
            bulk_g2s(s_scale, scale + g0 * 3, 3 * G * sizeof(float), &s_bar);
            bulk_g2s(s_quat, quat + g0 * 4, 4 * G * sizeof(float), &s_bar);
            bulk_g2s(s_op, opacity + g0, G * sizeof(float), &s_bar);
        }
        mbar_wait(&s_bar, 0);
    } else {
        slab_load<RP_NT>(s_xyz, xyz, g0 * 3, rows * 3);
        slab_load<RP_NT>(s_scale, scale, g0 * 3, rows * 3);
        slab_load<RP_NT>(s_quat, quat, g0 * 4, rows * 4);
        slab_load<RP_NT>(s_op, opacity, g0, rows);
    }
    cp_async_wait<0>();
    __syncthreads();
    const int t = tid;
    float px = 0.f, py = 0.f, pz = 0.f, op = 0.f;
    float cv[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // cov3d does not depend on the camera: once per Gaussian
    if (t < rows) {
        px = s_xyz[3 * t];
        py = s_xyz[3 * t + 1];
        pz = s_xyz[3 * t + 2];
        op = s_op[t];
        const float4 q = reinterpret_cast<const float4*>(s_quat)[t];
        cov3d_fwd(s_scale[3 * t], s_scale[3 * t + 1], s_scale[3 * t + 2], q.x, q.y, q.z, q.w, cv);
    }
    const float4* my_sh = reinterpret_cast<const float4*>(s_sh + (size_t)t * RS);

    for (int b = 0; b < views; ++b) {
        const Cam c = load_cam(intr + 4 * b, extr + (size_t)estride * b);
        const CamCenter cc = cam_center(c);
        const long long v0 = (long long)b * vstride + g0;  // first output row of this block in view b
        __syncthreads();  // the previous view's stores have read the output slabs (thread 0 waited); s_tsum drained

        float u = 0.f, v = 0.f, d = 0.f, cx = 0.f, cy = 0.f, cz = 0.f, hx = 0.f, hy = 0.f;
        int rad = 0, til = 0;
        if (t < rows) {
            if (!project_fwd(c, px, py, pz, W, H, nearest, extent, u, v, d)) u = v = d = 0.f;
            if (d != 0.f) {  // visible = depth != 0 (msplat/__init__.py:73)
                if (!ewa_fwd(c, px, py, pz, cv, u, v, gx, gy, cx, cy, cz, rad, til)) {
                    cx = cy = cz = 0.f;
                    rad = til = 0;
                }
            }
            float* f = s_feat + (size_t)t * Cpad;
            for (int k = 0; k < Cpad; ++k) f[k] = 0.f;
            if (with_depth) f[Cs] = d;
            if (til > 0) {
                float ex, ey, es, et;
                cull_extent(cx, cy, cz, op, ex, ey, es, et);
                cull_pack(ex, ey, es, et, hx, hy);  // hx, hy now hold the packed FP16 pairs of the blend record
                const float rx = px - cc.x, ry = py - cc.y, rz = pz - cc.z;
                const float inv = 1.0f / sqrtf(rx * rx + ry * ry + rz * rz);
                float bs[D];
                sh_basis<DEG>(rx * inv, ry * inv, rz * inv, bs, 1);
                for (int ch = 0; ch < Cs; ++ch) {
                    float acc = 0.f;
#pragma unroll
                    for (int k = 0; k < D / 4; ++k) {
                        const float4 x = my_sh[ch * (D / 4) + k];
                        acc = fmaf(x.x, bs[4 * k], acc);
                        acc = fmaf(x.y, bs[4 * k + 1], acc);
                        acc = fmaf(x.z, bs[4 * k + 2], acc);
                        acc = fmaf(x.w, bs[4 * k + 3], acc);
                    }
                    float val = acc + sh_bias;
                    if (clamp) val = fmaxf(val, 0.f);
                    f[ch] = val;
                }
            }
            float4* r = reinterpret_cast<float4*>(s_rec) + 2 * t;
            r[0] = make_float4(u, v, cx, cy);
            r[1] = make_float4(cz, op, hx, hy);
            s_uv[2 * t] = u;
            s_uv[2 * t + 1] = v;
            depth[v0 + t] = d;
            radius[v0 + t] = rad;
            tiles[v0 + t] = til;
        }
        if (total_tiles != nullptr) {  // M = sum(tiles): sizes the sort output (replaces the separate count pass)
            unsigned wsum = (unsigned)til;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
            if ((tid & 31) == 0 && wsum) atomicAdd(&s_tsum, wsum);
        }
        if (full) {
            fence_async_smem();  // this thread's slab writes -> visible to the bulk stores below
            __syncthreads();
            if (tid == 0) {
                pt_drain_tiles(&s_tsum, total_tiles, b);
                bulk_s2g(rec + v0 * 8, s_rec, 8 * G * sizeof(float));
                bulk_s2g(uv + v0 * 2, s_uv, 2 * G * sizeof(float));
                bulk_s2g(featp + v0 * Cpad, s_feat, (unsigned)((size_t)Cpad * G * sizeof(float)));
                bulk_commit();
                bulk_wait_read();  // shared memory must stay valid until the copy engine has read it
            }
        } else {
            __syncthreads();
            if (tid == 0) pt_drain_tiles(&s_tsum, total_tiles, b);
            slab_store<RP_NT>(rec, s_rec, v0 * 8, rows * 8);
            slab_store<RP_NT>(uv, s_uv, v0 * 2, rows * 2);
            slab_store<RP_NT>(featp, s_feat, v0 * Cpad, rows * Cpad);
        }
    }
}

template <int DEG, bool CAM>
__global__ void __launch_bounds__(RP_NT, 2) render_pre_bwd_pt_kernel(
    int P, int Cs, int Cpad, int with_depth, int accumulate, int views, long long vstride,
    const float* __restrict__ xyz, const float* __restrict__ scale, const float* __restrict__ quat,
    const float* __restrict__ shs, const float* __restrict__ intr, const float* __restrict__ extr, int estride,
    float sh_bias, int clamp, const int* __restrict__ tiles, const float* __restrict__ grec,
    const float* __restrict__ gfeat, float* __restrict__ dL_dxyz, float* __restrict__ dL_dscale,
    float* __restrict__ dL_dquat, float* __restrict__ dL_dopacity, float* __restrict__ dL_dshs,
    float* __restrict__ dL_dintr, float* __restrict__ dL_dextr) {
    constexpr int D = sh_dim(DEG);
    constexpr int G = RP_NT;
    static_assert(D % 4 == 0, "PT kernels need 16-byte coefficient units");
    extern __shared__ __align__(16) float sm[];
    float* s_xyz = sm;             // [G,3]  -> dL_dxyz slab after the last view
    float* s_scale = sm + 3 * G;   // [G,3]  -> dL_dscale slab
    float* s_quat = sm + 6 * G;    // [G,4]  -> dL_dquat slab
    float* s_gin = sm + 10 * G;    // [2][G, 8 + Cpad]  packed gradients of the current / next view
    const int gin_stride = (8 + Cpad) * G;
    float* s_sh = s_gin + 2 * gin_stride;  // [G, RS] coefficient rows, staged once
    __shared__ float s_red[8 * 16];
    __shared__ unsigned long long s_bar[3];  // inputs | gradient buffer 0 | gradient buffer 1
    const int rowf = Cs * D, RS = rp_pt_rs(rowf);

    const int tid = threadIdx.x;
    const long long g0 = (long long)blockIdx.x * G;
    const int rows = (int)min((long long)G, (long long)P - g0);
    const bool full = rows == G;
    if (tid == 0 && full) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_init(&s_bar[2], 1);
    }
    __syncthreads();
    pt_stage_rows(s_sh, shs, g0, rows, rowf, RS);
    const unsigned gin_bytes = (unsigned)((size_t)(8 + Cpad) * G * sizeof(float));
    auto prefetch = [&](int b) {  // thread 0 of a full block: packed gradients of view b -> buffer b & 1
        float* dst = s_gin + (b & 1) * gin_stride;
        const long long v0 = (long long)b * vstride + g0;
        mbar_expect_tx(&s_bar[1 + (b & 1)], gin_bytes);
        bulk_g2s(dst, grec + v0 * 8, 8 * G * sizeof(float), &s_bar[1 + (b & 1)]);
        bulk_g2s(dst + 8 * G, gfeat + v0 * Cpad, (unsigned)((size_t)Cpad * G * sizeof(float)), &s_bar[1 + (b & 1)]);
    };
    if (full) {
        if (tid == 0) {
            mbar_expect_tx(&s_bar[0], (unsigned)(10 * G * sizeof(float)));
            bulk_g2s(s_xyz, xyz + g0 * 3, 3 * G * sizeof(float), &s_bar[0]);
            bulk_g2s(s_scale, scale + g0 * 3, 3 * G * sizeof(float), &s_bar[0]);
            bulk_g2s(s_quat, quat + g0 * 4, 4 * G * sizeof(float), &s_bar[0]);
            prefetch(0);
        }
        mbar_wait(&s_bar[0], 0);
    } else {
        slab_load<RP_NT>(s_xyz, xyz, g0 * 3, rows * 3);
        slab_load<RP_NT>(s_scale, scale, g0 * 3, rows * 3);
        slab_load<RP_NT>(s_quat, quat, g0 * 4, rows * 4);
    }
    cp_async_wait<0>();
    __syncthreads();

    const int t = tid;
    const float4* my_sh = reinterpret_cast<const float4*>(s_sh + (size_t)t * RS);
    float* my_out = dL_dshs + (g0 + t) * (long long)rowf;  // this Gaussian's dL_dshs row
    // geometry gradients, summed over the views in registers
    float dx = 0.f, dy = 0.f, dz = 0.f, dop = 0.f;
    float ds[3] = {0.f, 0.f, 0.f}, dq[4] = {0.f, 0.f, 0.f, 0.f};

    // tiles_touched of the next view is fetched one view ahead (a dependent global load right after the wait for the
    // gradient slabs was 11 % of the kernel's stall samples)
    int til_next = t < rows ? tiles[g0 + t] : 0;
    for (int b = 0; b < views; ++b) {
        const Cam c = load_cam(intr + 4 * b, extr + (size_t)estride * b);
        const CamCenter cc = cam_center(c);
        const long long v0 = (long long)b * vstride + g0;
        const int til = til_next;
        if (b + 1 < views && t < rows) til_next = tiles[v0 + vstride + t];
        float* s_grec = s_gin + (b & 1) * gin_stride;   // [G,8]
        float* s_gfeat = s_grec + 8 * G;                // [G,Cpad]
        const bool acc_b = accumulate || b > 0;
        if (full) {
            // every thread is past its reads of the other buffer (barrier at the end of the previous view)
            if (tid == 0 && b + 1 < views) prefetch(b + 1);
            mbar_wait(&s_bar[1 + (b & 1)], (unsigned)((b >> 1) & 1));
        } else {
            slab_load<RP_NT>(s_grec, grec, v0 * 8, rows * 8);
            slab_load<RP_NT>(s_gfeat, gfeat, v0 * Cpad, rows * Cpad);
            __syncthreads();
        }
        float cam[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) cam[i] = 0.f;
        if (t < rows) {
            const bool vis = til > 0;
            const float px = s_xyz[3 * t], py = s_xyz[3 * t + 1], pz = s_xyz[3 * t + 2];
            const float* gf = s_gfeat + (size_t)t * Cpad;
            bool live = false;
            if (vis)
                for (int k = 0; k < Cs; ++k) live = live || (gf[k] != 0.f);
            float ax = 0.f, ay = 0.f, az = 0.f;
            if (live) {
                const float rx = px - cc.x, ry = py - cc.y, rz = pz - cc.z;
                const float inv = 1.0f / sqrtf(rx * rx + ry * ry + rz * rz);
                const float dirx = rx * inv, diry = ry * inv, dirz = rz * inv;
                float bs[D], w[D];
                sh_basis<DEG>(dirx, diry, dirz, bs, 1);
#pragma unroll
                for (int i = 0; i < D; ++i) w[i] = 0.f;
                for (int ch = 0; ch < Cs; ++ch) {
                    float sv[D];
                    float acc = 0.f;
#pragma unroll
                    for (int k = 0; k < D / 4; ++k) {
                        const float4 x = my_sh[ch * (D / 4) + k];
                        sv[4 * k] = x.x; sv[4 * k + 1] = x.y; sv[4 * k + 2] = x.z; sv[4 * k + 3] = x.w;
                    }
#pragma unroll
                    for (int i = 0; i < D; ++i) acc = fmaf(sv[i], bs[i], acc);
                    // same arithmetic as the forward pass -> same clamp decision (clamp_min passes x >= 0)
                    const float dv = (clamp && !(acc + sh_bias >= 0.f)) ? 0.f : gf[ch];
                    float* op = my_out + ch * D;
#pragma unroll
                    for (int k = 0; k < D / 4; ++k) {
                        const float o0 = bs[4 * k] * dv, o1 = bs[4 * k + 1] * dv, o2 = bs[4 * k + 2] * dv, o3 = bs[4 * k + 3] * dv;
                        if (acc_b) {
                            if (dv != 0.f) red_add_v4(op + 4 * k, o0, o1, o2, o3);  // no read of the old row
                        } else {
                            *reinterpret_cast<float4*>(op + 4 * k) = make_float4(o0, o1, o2, o3);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < D; ++i) w[i] = fmaf(sv[i], dv, w[i]);
                }
                float hx, hy, hz;  // dL_ddir
                sh_basis_grad<DEG>(dirx, diry, dirz, w, 1, hx, hy, hz);
                // dir = r / |r|  =>  dL_dr = (g - dir (dir . g)) / |r|
                const float dt = dirx * hx + diry * hy + dirz * hz;
                ax = (hx - dirx * dt) * inv;
                ay = (hy - diry * dt) * inv;
                az = (hz - dirz * dt) * inv;
                if (CAM) {
                    // r = p - centre, centre = -R^T t  =>  dL_dR[i][j] += dL_dr[j] t[i], dL_dt[i] += R[i][:] . dL_dr
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const float ti = c.e[4 * i + 3];
                        cam[4 + 4 * i + 0] += ax * ti;
                        cam[4 + 4 * i + 1] += ay * ti;
                        cam[4 + 4 * i + 2] += az * ti;
                        cam[4 + 4 * i + 3] += c.e[4 * i] * ax + c.e[4 * i + 1] * ay + c.e[4 * i + 2] * az;
                    }
                }
            } else if (!acc_b) {
                // write mode: every row of dL_dshs must be produced (zeros for untouched Gaussians)
                for (int k = 0; k < rowf / 4; ++k) reinterpret_cast<float4*>(my_out)[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (vis) {
                const float4 ga = reinterpret_cast<const float4*>(s_grec)[2 * t];      // dL_duv, dL_dconic.xy
                const float4 gb = reinterpret_cast<const float4*>(s_grec)[2 * t + 1];  // dL_dconic.z, dL_dopacity
                dop += gb.y;
                const float gd = with_depth ? gf[Cs] : 0.f;
                float bx, by, bz;
                project_bwd<CAM>(c, px, py, pz, ga.x, ga.y, gd, bx, by, bz, cam);
                ax += bx;
                ay += by;
                az += bz;
                const float4 q = reinterpret_cast<const float4*>(s_quat)[t];
                const float sx = s_scale[3 * t], sy = s_scale[3 * t + 1], sz = s_scale[3 * t + 2];
                float cv[6], dcv[6], ex, ey, ez;
                cov3d_fwd(sx, sy, sz, q.x, q.y, q.z, q.w, cv);
                if (ewa_bwd<CAM>(c, px, py, pz, cv, ga.z, ga.w, gb.x, ex, ey, ez, dcv, cam)) {
                    ax += ex;
                    ay += ey;
                    az += ez;
                    float vs[3] = {0.f, 0.f, 0.f}, vq[4] = {0.f, 0.f, 0.f, 0.f};
                    cov3d_bwd(sx, sy, sz, q.x, q.y, q.z, q.w, dcv, vs, vq);
#pragma unroll
                    for (int i = 0; i < 3; ++i) ds[i] += vs[i];
#pragma unroll
                    for (int i = 0; i < 4; ++i) dq[i] += vq[i];
                }
                dx += ax;
                dy += ay;
                dz += az;
            }
        }
        if (CAM) cam_reduce_atomic<RP_NT>(cam, dL_dintr ? dL_dintr + 4 * b : nullptr,
                                          dL_dextr ? dL_dextr + (size_t)estride * b : nullptr, s_red);
        __syncthreads();  // this view's gradient buffer may now be overwritten
    }

    if (t < rows) {
        s_xyz[3 * t] = dx;
        s_xyz[3 * t + 1] = dy;
        s_xyz[3 * t + 2] = dz;
        s_scale[3 * t] = ds[0];
        s_scale[3 * t + 1] = ds[1];
        s_scale[3 * t + 2] = ds[2];
        reinterpret_cast<float4*>(s_quat)[t] = make_float4(dq[0], dq[1], dq[2], dq[3]);
        if (accumulate) {
            if (dop != 0.f) atomicAdd(dL_dopacity + g0 + t, dop);  // result unused -> RED
        } else {
            dL_dopacity[g0 + t] = dop;
        }
    }
    if (full) {
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            if (accumulate) {
                bulk_s2g_add_f32(dL_dxyz + g0 * 3, s_xyz, 3 * G * sizeof(float));
                bulk_s2g_add_f32(dL_dscale + g0 * 3, s_scale, 3 * G * sizeof(float));
                bulk_s2g_add_f32(dL_dquat + g0 * 4, s_quat, 4 * G * sizeof(float));
            } else {
                bulk_s2g(dL_dxyz + g0 * 3, s_xyz, 3 * G * sizeof(float));
                bulk_s2g(dL_dscale + g0 * 3, s_scale, 3 * G * sizeof(float));
                bulk_s2g(dL_dquat + g0 * 4, s_quat, 4 * G * sizeof(float));
            }
            bulk_commit();
            bulk_wait_read();
        }
    } else {
        __syncthreads();
        if (accumulate) {
            slab_store_acc<RP_NT>(dL_dxyz, s_xyz, g0 * 3, rows * 3);
            slab_store_acc<RP_NT>(dL_dscale, s_scale, g0 * 3, rows * 3);
            slab_store_acc<RP_NT>(dL_dquat, s_quat, g0 * 4, rows * 4);
        } else {
            slab_store<RP_NT>(dL_dxyz, s_xyz, g0 * 3, rows * 3);
            slab_store<RP_NT>(dL_dscale, s_scale, g0 * 3, rows * 3);
            slab_store<RP_NT>(dL_dquat, s_quat, g0 * 4, rows * 4);
        }
    }
}

static size_t rp_pt_smem_fwd(int Cs, int D, int Cpad) {
    return ((size_t)rp_pt_fwd_io(Cpad) * RP_NT + (size_t)rp_pt_rs(Cs * D) * RP_NT) * sizeof(float);
}
static size_t rp_pt_smem_bwd(int Cs, int D, int Cpad) {
    return ((size_t)10 * RP_NT + 2 * (size_t)(8 + Cpad) * RP_NT + (size_t)rp_pt_rs(Cs * D) * RP_NT) * sizeof(float);
}
// PT path: degree 1 or 3, rows of at most RP_PT_MAX_ROW floats, 16-byte aligned rows; MSB_RP_PT=0 disables (A/B)
static bool rp_use_pt(int deg, int Cs) {
    static const int on = [] { const char* e = getenv("MSB_RP_PT"); return e ? atoi(e) : 1; }();
    return on && (deg == 1 || deg == 3) && Cs > 0 && Cs * sh_dim(deg) <= RP_PT_MAX_ROW;
}

static size_t rp_smem_fwd(int deg, int Cpad) {
    const int G = rp_gpb(deg);
    return ((size_t)21 * G + rp_bs(deg) + (size_t)Cpad * G + G) * sizeof(float);
}
static size_t rp_smem_bwd(int deg, int Cpad) {
    const int G = rp_gpb(deg);
    return ((size_t)10 * G + 2 * (size_t)(8 + Cpad) * G + (size_t)rp_bs(deg) + G) * sizeof(float);
}

struct RpFwdArgs {
    int P, Cs, Cpad, with_depth, views;
    long long vstride;
    const float *xyz, *scale, *quat, *opacity, *shs, *intr, *extr;
    int estride, W, H;
    float nearest, extent, sh_bias;
    int clamp;
    float *rec, *featp, *uv, *depth;
    int *radius, *tiles;
    unsigned long long* total_tiles;
};

template <int DEG>
static int rp_launch_fwd_pt(const RpFwdArgs& a, cudaStream_t st) {
    const size_t smem = rp_pt_smem_fwd(a.Cs, sh_dim(DEG), a.Cpad);
    cudaError_t e = cudaFuncSetAttribute(render_pre_fwd_pt_kernel<DEG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return set_error((int)e, "render_preprocess_fwd: cudaFuncSetAttribute failed");
    const unsigned grid = (unsigned)(((long long)a.P + RP_NT - 1) / RP_NT);
    render_pre_fwd_pt_kernel<DEG><<<grid, RP_NT, smem, st>>>(a.P, a.Cs, a.Cpad, a.with_depth, a.views, a.vstride, a.xyz,
                                                             a.scale, a.quat, a.opacity, a.shs, a.intr, a.extr,
                                                             a.estride, a.W, a.H, a.nearest, a.extent, a.sh_bias, a.clamp,
                                                             a.rec, a.featp, a.uv, a.depth, a.radius, a.tiles,
                                                             a.total_tiles);
    return check_launch("render_preprocess_fwd");
}

template <int DEG>
static int rp_launch_fwd(const RpFwdArgs& a, cudaStream_t st) {
    if constexpr (DEG == 1 || DEG == 3) {
        if (rp_use_pt(DEG, a.Cs) && rp_pt_smem_fwd(a.Cs, sh_dim(DEG), a.Cpad) <= 100 * 1024) return rp_launch_fwd_pt<DEG>(a, st);
    }
    const size_t smem = rp_smem_fwd(DEG, a.Cpad);
    if (smem > 200 * 1024) return set_error(MSB_ERR_RANGE, "render_preprocess_fwd: too many channels for shared memory");
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(render_pre_fwd_kernel<DEG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return set_error((int)e, "render_preprocess_fwd: cudaFuncSetAttribute failed");
    }
    const int G = rp_gpb(DEG);
    const unsigned grid = (unsigned)(((long long)a.P + G - 1) / G);
    render_pre_fwd_kernel<DEG><<<grid, RP_NT, smem, st>>>(a.P, a.Cs, a.Cpad, a.with_depth, a.views, a.vstride, a.xyz,
                                                          a.scale, a.quat, a.opacity, a.shs, a.intr, a.extr, a.estride,
                                                          a.W, a.H, a.nearest, a.extent, a.sh_bias, a.clamp, a.rec,
                                                          a.featp, a.uv, a.depth, a.radius, a.tiles, a.total_tiles);
    return check_launch("render_preprocess_fwd");
}

struct RpBwdArgs {
    int P, Cs, Cpad, with_depth, accumulate, views;
    long long vstride;
    const float *xyz, *scale, *quat, *shs, *intr, *extr;
    int estride;
    float sh_bias;
    int clamp;
    const int* tiles;
    const float *grec, *gfeat;
    float *dxyz, *dscale, *dquat, *dopacity, *dshs, *dintr, *dextr;
};

template <int DEG, bool CAM>
static int rp_launch_bwd_pt(const RpBwdArgs& a, cudaStream_t st) {
    const size_t smem = rp_pt_smem_bwd(a.Cs, sh_dim(DEG), a.Cpad);
    cudaError_t e = cudaFuncSetAttribute(render_pre_bwd_pt_kernel<DEG, CAM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return set_error((int)e, "render_preprocess_bwd: cudaFuncSetAttribute failed");
    const unsigned grid = (unsigned)(((long long)a.P + RP_NT - 1) / RP_NT);
    render_pre_bwd_pt_kernel<DEG, CAM><<<grid, RP_NT, smem, st>>>(
        a.P, a.Cs, a.Cpad, a.with_depth, a.accumulate, a.views, a.vstride, a.xyz, a.scale, a.quat, a.shs, a.intr,
        a.extr, a.estride, a.sh_bias, a.clamp, a.tiles, a.grec, a.gfeat, a.dxyz, a.dscale, a.dquat, a.dopacity, a.dshs,
        a.dintr, a.dextr);
    return check_launch("render_preprocess_bwd");
}

template <int DEG, bool CAM>
static int rp_launch_bwd(const RpBwdArgs& a, cudaStream_t st) {
    if constexpr (DEG == 1 || DEG == 3) {
        if (rp_use_pt(DEG, a.Cs) && rp_pt_smem_bwd(a.Cs, sh_dim(DEG), a.Cpad) <= 110 * 1024) return rp_launch_bwd_pt<DEG, CAM>(a, st);
    }
    const size_t smem = rp_smem_bwd(DEG, a.Cpad);
    if (smem > 200 * 1024) return set_error(MSB_ERR_RANGE, "render_preprocess_bwd: too many channels for shared memory");
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(render_pre_bwd_kernel<DEG, CAM>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error((int)e, "render_preprocess_bwd: cudaFuncSetAttribute failed");
    }
    const int G = rp_gpb(DEG);
    const unsigned grid = (unsigned)(((long long)a.P + G - 1) / G);
    render_pre_bwd_kernel<DEG, CAM><<<grid, RP_NT, smem, st>>>(
        a.P, a.Cs, a.Cpad, a.with_depth, a.accumulate, a.views, a.vstride, a.xyz, a.scale, a.quat, a.shs, a.intr,
        a.extr, a.estride, a.sh_bias, a.clamp, a.tiles, a.grec, a.gfeat, a.dxyz, a.dscale, a.dquat, a.dopacity, a.dshs,
        a.dintr, a.dextr);
    return check_launch("render_preprocess_bwd");
}

static inline bool rp_al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace msb

using namespace msb;

extern "C" {

int msb_blend_cpad(int C);

// Fused forward preprocess of the SH render path for a batch of `views` cameras over the same Gaussians
// (ONE launch; parameters and SH rows are read once).  intr [views,4]; extr [views,estride] with
// estride = 12 ([3,4]) or 16 ([4,4]).  Outputs are view-major with `vstride` rows per view (vstride >= P,
// a multiple of 4 when views > 1): rec [views,vstride,8] and featp [views,vstride,Cpad]
// (Cpad = msb_blend_cpad(Cs + with_depth)) in the blend kernels' packed layout, uv [views,vstride,2],
// depth / radius / tiles [views,vstride] (bit-identical to project_point / ewa_project); rows >= P of a view
// are not written.  total_dev [views] (device int64, optional) receives M = sum(tiles) per view (zeroed first).
int msb_render_preprocess_fwd_views(const float* xyz, const float* scale, const float* quat, const float* opacity,
                                    const float* shs, const float* intr, const float* extr, int estride, int P,
                                    int views, long long vstride, int Cs, int D, int with_depth, int W, int H,
                                    float nearest, float extent, float sh_bias, int clamp, float* rec, float* featp,
                                    float* uv, float* depth, int32_t* radius, int32_t* tiles, long long* total_dev,
                                    void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (views <= 0 || P < 0) return set_error(MSB_ERR_ARG, "render_preprocess_fwd: bad size");
    if (total_dev) {
        cudaError_t e = cudaMemsetAsync(total_dev, 0, (size_t)views * sizeof(long long), st);
        if (e != cudaSuccess) return set_error((int)e, "render_preprocess_fwd: memset failed");
    }
    if (P == 0) return MSB_OK;
    const int deg = sh_degree_of(D);
    if (Cs < 0 || deg < 0 || W <= 0 || H <= 0 || (estride != 12 && estride != 16) || vstride < P ||
        (views > 1 && (vstride & 3)))
        return set_error(MSB_ERR_ARG, "render_preprocess_fwd: bad size (D must be (deg+1)^2, deg <= 10; estride 12|16; "
                                      "vstride >= P and a multiple of 4)");
    if (!xyz || !scale || !quat || !opacity || (Cs > 0 && !shs) || !intr || !extr || !rec || !featp || !uv || !depth ||
        !radius || !tiles)
        return set_error(MSB_ERR_ARG, "render_preprocess_fwd: null pointer");
    if (!(rp_al16(xyz) && rp_al16(scale) && rp_al16(quat) && rp_al16(opacity) && rp_al16(shs) && rp_al16(rec) &&
          rp_al16(featp) && rp_al16(uv)))
        return set_error(MSB_ERR_ARG, "render_preprocess_fwd: 16-byte alignment");
    RpFwdArgs a{P, Cs, msb_blend_cpad(Cs + (with_depth ? 1 : 0)), with_depth ? 1 : 0, views, vstride, xyz, scale, quat,
                opacity, shs, intr, extr, estride, W, H, nearest, extent, sh_bias, clamp ? 1 : 0, rec, featp, uv, depth,
                radius, tiles, reinterpret_cast<unsigned long long*>(total_dev)};
    switch (deg) {
#define MSB_RP_CASE(d) \
    case d:            \
        return rp_launch_fwd<d>(a, st);
        MSB_RP_CASE(0) MSB_RP_CASE(1) MSB_RP_CASE(2) MSB_RP_CASE(3) MSB_RP_CASE(4) MSB_RP_CASE(5)
        MSB_RP_CASE(6) MSB_RP_CASE(7) MSB_RP_CASE(8) MSB_RP_CASE(9) MSB_RP_CASE(10)
#undef MSB_RP_CASE
    }
    return set_error(MSB_ERR_ARG, "render_preprocess_fwd: unsupported degree");
}

// One camera (views = 1).  total_dev (8 bytes of device memory) / total_host (pinned) are optional: with
// total_dev, M = sum(tiles) is accumulated by the kernel into *total_dev (zeroed first); with total_host it is
// also copied asynchronously to *total_host.  Replaces msb_sort_scan for this path.
int msb_render_preprocess_fwd(const float* xyz, const float* scale, const float* quat, const float* opacity,
                              const float* shs, const float* intr, const float* extr, int P, int Cs, int D,
                              int with_depth, int W, int H, float nearest, float extent, float sh_bias, int clamp,
                              float* rec, float* featp, float* uv, float* depth, int32_t* radius, int32_t* tiles,
                              long long* total_dev, long long* total_host, void* stream) {
    if (total_host != nullptr && total_dev == nullptr)
        return set_error(MSB_ERR_ARG, "render_preprocess_fwd: total_host needs total_dev");
    int rc = msb_render_preprocess_fwd_views(xyz, scale, quat, opacity, shs, intr, extr, 12, P, 1, P, Cs, D, with_depth,
                                             W, H, nearest, extent, sh_bias, clamp, rec, featp, uv, depth, radius, tiles,
                                             total_dev, stream);
    if (rc) return rc;
    if (total_host) {  // M = sum(tiles) -> pinned host memory; the caller synchronises the stream before reading it
        if (P == 0) {
            *total_host = 0;
            return MSB_OK;
        }
        cudaError_t e = cudaMemcpyAsync(total_host, total_dev, sizeof(long long), cudaMemcpyDeviceToHost,
                                        (cudaStream_t)stream);
        if (e != cudaSuccess) return set_error((int)e, "render_preprocess_fwd: cudaMemcpyAsync failed");
    }
    return MSB_OK;
}

// Fused backward for a view batch (ONE launch).  tiles [views,vstride], grec [views,vstride,8] =
// {dL_duv.xy, dL_dconic.xyz, dL_dopacity, -, -} and gfeat [views,vstride,Cpad] are the forward's tile
// counts and the packed gradients written by msb_blend_packed_bwd_views; per-Gaussian pointers (xyz ...
// dL_dshs, tiles, grec, gfeat) may be offset to a slab of P Gaussians.  The gradients are summed over the
// views.  accumulate != 0: added to the outputs; otherwise every output element is written.
// dL_dintr [views,4] / dL_dextr [views,estride] may be NULL; otherwise they are accumulated into.
int msb_render_preprocess_bwd_views(const float* xyz, const float* scale, const float* quat, const float* shs,
                                    const float* intr, const float* extr, int estride, const int32_t* tiles,
                                    const float* grec, const float* gfeat, int P, int views, long long vstride, int Cs, int D, int with_depth, float sh_bias, int clamp,
                                    int accumulate, float* dL_dxyz, float* dL_dscale, float* dL_dquat,
                                    float* dL_dopacity, float* dL_dshs, float* dL_dintr, float* dL_dextr,
                                    void* stream) {
    if (P == 0) return MSB_OK;
    const int deg = sh_degree_of(D);
    if (P < 0 || Cs < 0 || deg < 0 || views <= 0 || (estride != 12 && estride != 16) || vstride < P ||
        (views > 1 && (vstride & 3)))
        return set_error(MSB_ERR_ARG, "render_preprocess_bwd: bad size");
    if (!xyz || !scale || !quat || (Cs > 0 && (!shs || !dL_dshs)) || !intr || !extr || !tiles || !grec || !gfeat ||
        !dL_dxyz || !dL_dscale || !dL_dquat || !dL_dopacity)
        return set_error(MSB_ERR_ARG, "render_preprocess_bwd: null pointer");
    if (!(rp_al16(xyz) && rp_al16(scale) && rp_al16(quat) && rp_al16(shs) && rp_al16(grec) && rp_al16(gfeat) &&
          rp_al16(dL_dxyz) && rp_al16(dL_dscale) && rp_al16(dL_dquat) && rp_al16(dL_dshs)))
        return set_error(MSB_ERR_ARG, "render_preprocess_bwd: 16-byte alignment");
    RpBwdArgs a{P, Cs, msb_blend_cpad(Cs + (with_depth ? 1 : 0)), with_depth ? 1 : 0, accumulate ? 1 : 0, views,
                vstride, xyz, scale, quat, shs, intr, extr, estride, sh_bias, clamp ? 1 : 0, tiles, grec, gfeat,
                dL_dxyz, dL_dscale, dL_dquat, dL_dopacity, dL_dshs, dL_dintr, dL_dextr};
    cudaStream_t st = (cudaStream_t)stream;
    const bool camg = dL_dintr || dL_dextr;
    switch (deg) {
#define MSB_RP_CASE(d) \
    case d:            \
        return camg ? rp_launch_bwd<d, true>(a, st) : rp_launch_bwd<d, false>(a, st);
        MSB_RP_CASE(0) MSB_RP_CASE(1) MSB_RP_CASE(2) MSB_RP_CASE(3) MSB_RP_CASE(4) MSB_RP_CASE(5)
        MSB_RP_CASE(6) MSB_RP_CASE(7) MSB_RP_CASE(8) MSB_RP_CASE(9) MSB_RP_CASE(10)
#undef MSB_RP_CASE
    }
    return set_error(MSB_ERR_ARG, "render_preprocess_bwd: unsupported degree");
}

// One camera (views = 1).  dL_dintr [4] / dL_dextr [12] may be NULL; otherwise they are accumulated into.
int msb_render_preprocess_bwd(const float* xyz, const float* scale, const float* quat, const float* shs,
                              const float* intr, const float* extr, const int32_t* tiles, const float* grec,
                              const float* gfeat, int P, int Cs, int D, int with_depth, float sh_bias, int clamp,
                              int accumulate, float* dL_dxyz, float* dL_dscale, float* dL_dquat, float* dL_dopacity,
                              float* dL_dshs, float* dL_dintr, float* dL_dextr, void* stream) {
    return msb_render_preprocess_bwd_views(xyz, scale, quat, shs, intr, extr, 12, tiles, grec, gfeat, P, 1, P,
                                           Cs, D, with_depth, sh_bias, clamp, accumulate, dL_dxyz, dL_dscale, dL_dquat,
                                           dL_dopacity, dL_dshs, dL_dintr, dL_dextr, stream);
}

}  // extern "C"
