// msplat_b200/csrc/blend.cu -- tile-based alpha blending, forward and backward.
//
// Replaces alphaBlendingForward/Backward and their channel-chunk dispatcher
//   /root/reference/msplat/src/alpha_blending.cu:16-110 (K11), :112-246 (K12), :248-573 (D1).
// Per-pixel semantics are exactly those of the reference: front-to-back over the
// tile's depth-sorted list, skip on power > 0 / alpha < 1/255, terminate before the entry that
// would push T below 1e-4, out = F + T*bg; backward replays T back-to-front from final_T and
// ncontrib.
//
// Design (one CTA of 256 threads per 16x16 tile, like the reference, but):
//  * a pack pass turns the per-Gaussian inputs into one 32-byte record
//      {u, v, conic.x, conic.y | conic.z, opacity, fp16x2(hx, hy), fp16x2(hs, ht)}
//    plus a 16-byte-aligned feature row, so a batch of 256 list entries is staged into shared
//    memory with 16-byte cp.async (LDGSTS) copies, double-buffered against the blend loop;
//  * each warp owns an 8x4 pixel block; lane l tests staged Gaussian 32k+l against the warp's
//    block (conservative alpha-footprint octagon, blend_math.cuh) and a ballot yields the
//    Gaussians worth visiting -- most of a tile's list never touches a given 8x4 block, so the
//    per-pair work drops by ~3-4x without changing any pixel's result;
//  * warp-vote early termination (a warp stops when its 32 pixels are done, the CTA when all are);
//  * features live in shared memory (the reference re-reads them from global per pair);
//  * backward: every pair parks two scalars (X, w) in shared memory; a Gaussian-parallel phase turns
//    them into the gradient moments in registers and leaves the SM as three vector reductions
//    (red.global.add.v4/.v2) per Gaussian and warp, into a packed 32-byte gradient record; the
//    reference issues (6+C) atomics per lane;
//  * the backward walks only list positions below the tile's max ncontrib.
#include <stdlib.h>

#include "blend_math.cuh"

namespace msb {

constexpr int BL_NT = 256;
constexpr int BL_BATCH = 256;

// ------------------------------------------------------------------------------------------------
// pack / unpack
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) blend_pack_kernel(int P, int C, int Cpad, const float2* __restrict__ uv,
                                                         const float* __restrict__ conic,
                                                         const float* __restrict__ opacity,
                                                         const float* __restrict__ feature,
                                                         float4* __restrict__ rec, float* __restrict__ featp) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float2 p = uv[i];
    const float cx = conic[3 * i], cy = conic[3 * i + 1], cz = conic[3 * i + 2];
    const float op = opacity[i];
    float hx, hy, hs, ht, p0, p1;
    cull_extent(cx, cy, cz, op, hx, hy, hs, ht);
    cull_pack(hx, hy, hs, ht, p0, p1);
    rec[2 * i] = make_float4(p.x, p.y, cx, cy);
    rec[2 * i + 1] = make_float4(cz, op, p0, p1);
    if (featp != nullptr) {
        for (int k = 0; k < Cpad; ++k) featp[i * Cpad + k] = k < C ? feature[i * C + k] : 0.f;
    }
}

__global__ void __launch_bounds__(256) blend_unpack_kernel(int P, int C, int Cpad, const float* __restrict__ grec,
                                                           const float* __restrict__ gfeat,
                                                           float* __restrict__ dL_duv,
                                                           float* __restrict__ dL_dconic,
                                                           float* __restrict__ dL_dopacity,
                                                           float* __restrict__ dL_dfeature) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float4 a = reinterpret_cast<const float4*>(grec)[2 * i];
    const float4 b = reinterpret_cast<const float4*>(grec)[2 * i + 1];
    dL_duv[2 * i] = a.x;
    dL_duv[2 * i + 1] = a.y;
    dL_dconic[3 * i] = a.z;
    dL_dconic[3 * i + 1] = a.w;
    dL_dconic[3 * i + 2] = b.x;
    dL_dopacity[i] = b.y;
    if (dL_dfeature != nullptr)
        for (int k = 0; k < C; ++k) dL_dfeature[i * C + k] = gfeat[i * Cpad + k];
}

// ------------------------------------------------------------------------------------------------
// shared-memory stage: 256 records + 256 feature rows
// ------------------------------------------------------------------------------------------------
template <int CH, int B = BL_BATCH>
struct Stage {
    float4 rec[B * 2];
    float feat[B * CH];
};

template <int CH, int B>
__device__ __forceinline__ void stage_issue(Stage<CH, B>& st, int slot, int id, const float4* __restrict__ rec,
                                            const float* __restrict__ featp, int fstride, int foff) {
    const float4* r = rec + 2 * (long long)id;
    cp_async16(&st.rec[2 * slot], r);
    cp_async16(&st.rec[2 * slot + 1], r + 1);
    const float* f = featp + (long long)id * fstride + foff;
#pragma unroll
    for (int k = 0; k < CH; k += 4) cp_async16(&st.feat[slot * CH + k], f + k);
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
// The per-pair body has no divergent branches: a pair that fails a test blends with weight 0
// (ffma(T, 0 * f, F) == F), so the image is bit-identical to a branching formulation and to the
// reference, and a warp-visit costs ~40 issue slots.
// COUNT (diagnostic, bench.py): instead of colours the kernel writes, per pixel, the number of list entries
// that actually blend (pass the power / alpha tests before termination) into `image` reinterpreted as int32
// [views,H,W] -- the "blended pairs" the roofline of the blend kernels is computed on.
template <int CH, bool COUNT = false>
__global__ void __launch_bounds__(BL_NT) blend_fwd_kernel(const float4* __restrict__ rec,
                                                          const float* __restrict__ featp, int fstride, int foff,
                                                          const int* __restrict__ ids,
                                                          const int2* __restrict__ tile_range, float bg,
                                                          int c_valid, int W, int H, int write_aux,
                                                          float* __restrict__ final_T, int* __restrict__ ncontrib,
                                                          float* __restrict__ image, long long img_vstride) {
    extern __shared__ __align__(16) unsigned char bl_raw[];
    Stage<CH>* stages = reinterpret_cast<Stage<CH>*>(bl_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gxt = (W + MSB_TILE - 1) / MSB_TILE;
    // blockIdx.z = view of a view batch: tile ids, Gaussian ids and tile ranges are those of the batch's
    // single sort (view * T + tile, view * P + index); images / final_T / ncontrib are [views, ...]
    const int view = blockIdx.z;
    const int tile = (view * (int)gridDim.y + blockIdx.y) * gxt + blockIdx.x;
    const int bx0 = blockIdx.x * MSB_TILE + (warp & 1) * 8, by0 = blockIdx.y * MSB_TILE + (warp >> 1) * 4;
    const int px = bx0 + (lane & 7), py = by0 + (lane >> 3);
    const float pxf = (float)px, pyf = (float)py;
    const float wx0 = (float)bx0, wy0 = (float)by0;
    const bool inside = px < W && py < H;
    bool done = !inside;

    const int2 range = tile_range[tile];
    const int n = range.y - range.x;
    const int nb = (n + BL_BATCH - 1) / BL_BATCH;

    float T = 1.0f;
    int last = 0;
    int nblend = 0;
    float F[CH];  // accumulated colour
#pragma unroll
    for (int k = 0; k < CH; ++k) F[k] = 0.f;

    int id_next = 0;
    if (nb > 0) {
        if (tid < n) stage_issue(stages[0], tid, ids[range.x + tid], rec, featp, fstride, foff);
        cp_async_commit();
        if (BL_BATCH + tid < n) id_next = ids[range.x + BL_BATCH + tid];
    }
    for (int b = 0; b < nb; ++b) {
        cp_async_wait<0>();
        if (__syncthreads_and(done)) break;
        if (b + 1 < nb) {
            if ((b + 1) * BL_BATCH + tid < n)
                stage_issue(stages[(b + 1) & 1], tid, id_next, rec, featp, fstride, foff);
            cp_async_commit();
            if ((b + 2) * BL_BATCH + tid < n) id_next = ids[range.x + (b + 2) * BL_BATCH + tid];
        }
        const Stage<CH>& st = stages[b & 1];
        const int cnt = min(BL_BATCH, n - b * BL_BATCH);
        if (__all_sync(0xffffffffu, done)) continue;
        int lastj = 0;  // 1 + slot of this pixel's last blended entry within the batch (0: none yet)
        for (int k0 = 0; k0 < cnt; k0 += 32) {
            bool hit = false;
            if (k0 + lane < cnt) {
                const float4 r0 = st.rec[2 * (k0 + lane)];
                const float4 r1 = st.rec[2 * (k0 + lane) + 1];
                hit = !cull_miss(r0.x, r0.y, r1.z, r1.w, wx0, wy0, 8.0f, 4.0f);
            }
            unsigned m = __ballot_sync(0xffffffffu, hit);
            while (m) {
                const int j = k0 + __ffs(m) - 1;
                m &= m - 1;
                const float4 r0 = st.rec[2 * j];      // u, v, cx, cy   (broadcast LDS.128)
                const float4 r1 = st.rec[2 * j + 1];  // cz, opacity, packed cull extents
                const float dx = fadd(r0.x, -pxf), dy = fadd(r0.y, -pyf);
                const float power = pair_power(dx, dy, r0.z, r0.w, r1.x);
                const float G = ex2_approx(fmul(power, kLog2e));
                const float alpha = fmin_ftz(fmul(r1.y, G), kAlphaMax);
                const bool ok = !done && !(power > 0.0f) && !(alpha < kAlphaMin);
                if (!__any_sync(0xffffffffu, ok)) continue;  // ~1 visit in 5: box hit, but no pixel of the warp blends
                const float nT = fmul(T, fadd(-alpha, 1.0f));
                const bool term = ok && (nT < kTmin);  // alpha_blending.cu:90-94: entry not blended
                const bool blend = ok && !term;
                done = done || term;
                const float a = blend ? alpha : 0.0f;
                const float* f = &st.feat[j * CH];
                const f32x2 a2 = pk2(a, a);
#pragma unroll
                for (int k = 0; k < CH; k += 4) {  // F_k = fma(T, alpha * f_k, F_k), same rounding as the scalar ops
                    const float4 fv = *reinterpret_cast<const float4*>(f + k);
                    float p0, p1, p2, p3;
                    upk2(mul2(a2, pk2(fv.x, fv.y)), p0, p1);  // packed products, scalar accumulation: the
                    upk2(mul2(a2, pk2(fv.z, fv.w)), p2, p3);  // loop-carried FFMA2 pairs cost a MOV per register
                    F[k] = ffma(T, p0, F[k]);
                    F[k + 1] = ffma(T, p1, F[k + 1]);
                    F[k + 2] = ffma(T, p2, F[k + 2]);
                    F[k + 3] = ffma(T, p3, F[k + 3]);
                }
                T = blend ? nT : T;
                lastj = blend ? j + 1 : lastj;
                if (COUNT) nblend += blend ? 1 : 0;
            }
            if (__all_sync(0xffffffffu, done)) break;
        }
        if (lastj) last = b * BL_BATCH + lastj;  // list position + 1 (alpha_blending.cu: n_contrib)
    }
    cp_async_wait<0>();
    if (inside) {
        const long long hw = (long long)H * W;
        const long long pix = (long long)py * W + px;
        if (write_aux) {
            final_T[view * hw + pix] = T;
            ncontrib[view * hw + pix] = last;
        }
        if (COUNT) {
            reinterpret_cast<int*>(image)[view * hw + pix] = nblend;
        } else {
            float* img = image + view * img_vstride;
#pragma unroll
            for (int k = 0; k < CH; ++k)
                if (k < c_valid) img[k * hw + pix] = ffma(T, bg, F[k]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward: pixel-parallel replay + Gaussian-parallel gradient accumulation
// ------------------------------------------------------------------------------------------------
// Every gradient of a (pixel p, Gaussian g) pair is a product of two scalars that need the
// sequential per-pixel replay,
//     X(g,p) = G dL/dalpha            (alpha_blending.cu:222-231, 243)
//     w(g,p) = alpha T                (:218-219)
// with factors that depend only on g and on the pixel position:
//     dL/dopacity = sum_p X                        dL/dfeature_k = sum_p w dpix_k(p)
//     dL/du = -o (cx sum X dx + cy sum X dy)       dL/dv = -o (cz sum X dy + cy sum X dx)
//     dL/dconic = -o (0.5 sum X dx^2, sum X dx dy, 0.5 sum X dy^2)         (:232-242)
// Per staged batch of B list entries and per warp (8x4 pixels):
// Pass 1 (lane = list slot): footprint test of the whole batch, the slots that pass are compacted in list
// order into the warp's u16 hit list in shared memory.
// Pass 2 / phase 1 (lane = pixel, as in the forward): walk the hit list two slots per iteration (their alpha
// evaluations interleave), replay the visits that blend back to front and park (X, w) of each visit in a
// per-warp shared-memory matrix [GQ Gaussians][32 pixels + pad] (one STS.64 per visit; lane 0 puts the
// visit's staged slot into the pad cell).
// Phase 2 (lane = Gaussian, every GQ = 16 visits): lane (g, h) walks 16 of the 32 pixels of Gaussian g's row
// and accumulates raw moments about the block origin and CH feature sums IN REGISTERS, shifts the moments to
// the Gaussian's centre, combines the two halves with one shuffle per value; the sums leave the SM as three
// vector reductions (red.global.add.v4/.v2) per Gaussian.  No cross-lane reduction per visit; ~88 issue slots
// per replayed warp-visit over the whole kernel (CH = 4; ncu, BASELINE config #3).
constexpr int BW_GQ = 16;  // Gaussians per phase-2 group
constexpr int BW_PS = 33;  // row stride of the (X, w) matrix in float2: conflict-free LDS.64 in both phases


template <int CH, int B>
struct Bwd3 {
    static constexpr size_t STAGE_BYTES = 2 * sizeof(Stage<CH, B>);
    static constexpr size_t XW_BYTES = (size_t)(BL_NT / 32) * BW_GQ * BW_PS * sizeof(float2);
    static constexpr size_t DPIX_BYTES = (size_t)BL_NT * CH * sizeof(float);
    static constexpr int HL_STRIDE = B + 2;  // hit list of one warp: batch slots that pass the footprint test (+ pad)
    static constexpr size_t HL_BYTES = (size_t)(BL_NT / 32) * HL_STRIDE * sizeof(unsigned short);
    static constexpr size_t SMEM = STAGE_BYTES + XW_BYTES + DPIX_BYTES + HL_BYTES;
};

// phase 2 for the first n (<= BW_GQ) parked visits of this warp
template <int CH, int B>
__device__ __forceinline__ void bwd_reduce_group(int n, int lane, const Stage<CH, B>& st, const int* __restrict__ sid,
                                                 const float2* __restrict__ xw,
                                                 const float* __restrict__ dpw, float wx0, float wy0,
                                                 float* __restrict__ grec, float* __restrict__ gfeat, int fstride,
                                                 int foff, int geom_grads) {
    __syncwarp();
    const int g = lane & (BW_GQ - 1), h = lane >> 4;
    // Raw moments of X over this half's 2 x 8 pixels about the block's column origin, one set per pixel row:
    // a = sum X, ax = sum X x, axx = sum X x^2 with x = 0..7 a compile-time constant of the unrolled loop (FFMA with
    // an immediate; no per-pixel dx / dy arithmetic).  They are shifted to the Gaussian's centre once per group below.
    float a0 = 0.f, ax0 = 0.f, axx0 = 0.f, a1 = 0.f, ax1 = 0.f, axx1 = 0.f;
    float m0 = 0.f, mx = 0.f, my = 0.f, mxx = 0.f, mxy = 0.f, myy = 0.f;
    f32x2 fs[CH / 2];
#pragma unroll
    for (int k = 0; k < CH / 2; ++k) fs[k] = pk2(0.f, 0.f);
    float cx = 0.f, cy = 0.f, cz = 0.f, op = 0.f;
    int id = 0;
    if (g < n) {
        const int j = __float_as_int(xw[g * BW_PS + 32].x);  // the row's pad cell holds the staged slot
        const float4 r0 = st.rec[2 * j];
        const float4 r1 = st.rec[2 * j + 1];
        id = sid[j];
        cx = r0.z; cy = r0.w; cz = r1.x; op = r1.y;
        const float2* row = xw + g * BW_PS + h * 16;
        const float4* dp = reinterpret_cast<const float4*>(dpw + h * 16 * CH);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const float2 v = row[q];
            const float x = (float)(q & 7);
            if (q < 8) {
                a0 += v.x;
                ax0 = fmaf(v.x, x, ax0);
                axx0 = fmaf(v.x, x * x, axx0);
            } else {
                a1 += v.x;
                ax1 = fmaf(v.x, x, ax1);
                axx1 = fmaf(v.x, x * x, axx1);
            }
            const f32x2 w2 = pk2(v.y, v.y);
#pragma unroll
            for (int k = 0; k < CH; k += 4) {
                const float4 d = dp[q * (CH / 4) + k / 4];
                fs[k / 2] = fma2(w2, pk2(d.x, d.y), fs[k / 2]);
                fs[k / 2 + 1] = fma2(w2, pk2(d.z, d.w), fs[k / 2 + 1]);
            }
        }
        // dx = ux - x, dy = dy0 (row 2h) or dy1 (row 2h + 1)
        const float ux = r0.x - wx0;
        const float dy0 = r0.y - (wy0 + (float)(2 * h));
        const float dy1 = dy0 - 1.0f;
        const float s0 = a0 + a1, sx = ax0 + ax1, sxx = axx0 + axx1;
        m0 = s0;
        mx = fmaf(ux, s0, -sx);                             // sum X dx
        mxx = fmaf(ux, fmaf(ux, s0, -2.0f * sx), sxx);      // sum X dx^2
        my = fmaf(dy0, a0, dy1 * a1);                       // sum X dy
        myy = fmaf(dy0 * dy0, a0, dy1 * dy1 * a1);          // sum X dy^2
        mxy = fmaf(dy0, fmaf(ux, a0, -ax0), dy1 * fmaf(ux, a1, -ax1));  // sum X dx dy
    }
    // combine the two pixel halves (every lane takes part; idle lanes hold zeros)
    m0 += __shfl_xor_sync(0xffffffffu, m0, 16);
    mx += __shfl_xor_sync(0xffffffffu, mx, 16);
    my += __shfl_xor_sync(0xffffffffu, my, 16);
    mxx += __shfl_xor_sync(0xffffffffu, mxx, 16);
    mxy += __shfl_xor_sync(0xffffffffu, mxy, 16);
    myy += __shfl_xor_sync(0xffffffffu, myy, 16);
    float f[CH];
#pragma unroll
    for (int k = 0; k < CH; k += 2) {
        upk2(fs[k / 2], f[k], f[k + 1]);
        f[k] += __shfl_xor_sync(0xffffffffu, f[k], 16);
        f[k + 1] += __shfl_xor_sync(0xffffffffu, f[k + 1], 16);
    }
    if (g < n) {
        if (h == 0) {
            if (geom_grads) {
                const float nop = -op;
                float* gp = grec + (long long)id * 8;
                red_add_v4(gp, nop * fmaf(cx, mx, cy * my), nop * fmaf(cz, my, cy * mx), 0.5f * nop * mxx, nop * mxy);
                red_add_v2(gp + 4, 0.5f * nop * myy, m0);
            }
        } else {
            float* fp = gfeat + (long long)id * fstride + foff;
#pragma unroll
            for (int k = 0; k < CH; k += 4) red_add_v4(fp + k, f[k], f[k + 1], f[k + 2], f[k + 3]);
        }
    }
    __syncwarp();
}

template <int CH, int B, int MINB>
__global__ void __launch_bounds__(BL_NT, MINB) blend_bwd_kernel(const float4* __restrict__ rec,
                                                          const float* __restrict__ featp, int fstride, int foff,
                                                          const int* __restrict__ ids,
                                                          const int2* __restrict__ tile_range, float bg,
                                                          int c_valid, int W, int H,
                                                          const float* __restrict__ final_T,
                                                          const int* __restrict__ ncontrib,
                                                          const float* __restrict__ dL_dimage,
                                                          long long img_vstride,
                                                          float* __restrict__ grec, float* __restrict__ gfeat,
                                                          int geom_grads) {
    extern __shared__ __align__(16) unsigned char bl_raw[];
    using L = Bwd3<CH, B>;
    static_assert(B <= BL_NT, "one staged slot per thread");
    Stage<CH, B>* stages = reinterpret_cast<Stage<CH, B>*>(bl_raw);
    float2* s_xw = reinterpret_cast<float2*>(bl_raw + L::STAGE_BYTES);
    float* s_dpix = reinterpret_cast<float*>(bl_raw + L::STAGE_BYTES + L::XW_BYTES);
    unsigned short* s_hl =
        reinterpret_cast<unsigned short*>(bl_raw + L::STAGE_BYTES + L::XW_BYTES + L::DPIX_BYTES);
    __shared__ int s_id[2 * B];  // [2][B] Gaussian ids of the staged slots
    __shared__ int s_max[BL_NT / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gxt = (W + MSB_TILE - 1) / MSB_TILE;
    const int view = blockIdx.z;  // see blend_fwd_kernel
    const int tile = (view * (int)gridDim.y + blockIdx.y) * gxt + blockIdx.x;
    const int bx0 = blockIdx.x * MSB_TILE + (warp & 1) * 8, by0 = blockIdx.y * MSB_TILE + (warp >> 1) * 4;
    const int px = bx0 + (lane & 7), py = by0 + (lane >> 3);
    const float pxf = (float)px, pyf = (float)py;
    const float wx0 = (float)bx0, wy0 = (float)by0;
    const bool inside = px < W && py < H;
    const long long hw = (long long)H * W;
    const long long pix = (long long)py * W + px;
    final_T += view * hw;
    ncontrib += view * hw;
    dL_dimage += view * img_vstride;

    const int2 range = tile_range[tile];
    const int lc = inside ? min(ncontrib[pix], range.y - range.x) : 0;  // this pixel's last contributor
    int wmax = lc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if (lane == 0) s_max[warp] = wmax;

    const float T_final = inside ? final_T[pix] : 0.f;
    float Tn = -T_final;  // the replay keeps the transmittance negated (see replay)
    f32x2 dpix2[CH / 2], S2[CH / 2];  // cotangent and suffix colour, two channels per 64-bit register pair
    float bgdot = 0.f;
    float* dpw = s_dpix + warp * 32 * CH;  // this warp's cotangents, [pixel][channel]
    {
        float dpix[CH];
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            dpix[k] = (inside && k < c_valid) ? dL_dimage[k * hw + pix] : 0.f;
            bgdot = fmaf(bg, dpix[k], bgdot);
        }
#pragma unroll
        for (int k = 0; k < CH; k += 2) {
            dpix2[k / 2] = pk2(dpix[k], dpix[k + 1]);
            S2[k / 2] = pk2(0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < CH; k += 4)
            *reinterpret_cast<float4*>(dpw + lane * CH + k) = make_float4(dpix[k], dpix[k + 1], dpix[k + 2], dpix[k + 3]);
    }
    __syncthreads();
    int maxc = 0;
#pragma unroll
    for (int w = 0; w < BL_NT / 32; ++w) maxc = max(maxc, s_max[w]);
    const int nb = (maxc + B - 1) / B;
    const float nbg = -T_final * bgdot;  // background term of dL_dalpha, still to be divided by (1 - alpha)

    float2* xw = s_xw + warp * (BW_GQ * BW_PS);
    float2* xw_wr = xw + lane;  // this lane's cell in the row of the next parked visit
    // a parked visit occupies one row; lane 0 also stores the visit's staged slot in the row's pad cell (column 32)
    int cnt = 0;                // parked visits (warp-uniform)

    // batches walk the list back to front: batch b, slot j <-> list position maxc-1-(b*256+j)
    int id_next = 0;
    if (nb > 0) {
        if (tid < B && tid < maxc) {
            const int id = ids[range.x + maxc - 1 - tid];
            s_id[tid] = id;
            stage_issue(stages[0], tid, id, rec, featp, fstride, foff);
        }
        cp_async_commit();
        if (tid < B && B + tid < maxc) id_next = ids[range.x + maxc - 1 - (B + tid)];
    }
    for (int b = 0; b < nb; ++b) {
        cp_async_wait<0>();
        __syncthreads();
        if (b + 1 < nb) {
            if (tid < B && (b + 1) * B + tid < maxc) {
                s_id[((b + 1) & 1) * B + tid] = id_next;
                stage_issue(stages[(b + 1) & 1], tid, id_next, rec, featp, fstride, foff);
            }
            cp_async_commit();
            if (tid < B && (b + 2) * B + tid < maxc) id_next = ids[range.x + maxc - 1 - ((b + 2) * B + tid)];
        }
        const Stage<CH, B>& st = stages[b & 1];
        const int* sid = s_id + (b & 1) * B;
        const int bcnt = min(B, maxc - b * B);
        const int pos0 = maxc - 1 - b * B;  // list position of slot 0 of this batch
        if (pos0 - (bcnt - 1) >= wmax) continue;   // whole batch lies beyond every pixel of this warp
        // Pass 1: footprint test of the whole batch (lane <-> slot k0 + lane of a 32-slot window); the slots that
        // pass are compacted, in list order, into the warp's hit list.  Pass 2 then walks that list two slots per
        // iteration: no find-first-set chain per hit, and an odd hit is left over once per batch, not once per window.
        unsigned short* hl = s_hl + warp * L::HL_STRIDE;
        int nh = 0;  // hits of this warp in the batch
        __syncwarp();
        for (int k0 = 0; k0 < bcnt; k0 += 32) {
            bool hit = false;
            if (k0 + lane < bcnt) {
                const float4 r0 = st.rec[2 * (k0 + lane)];
                const float4 r1 = st.rec[2 * (k0 + lane) + 1];
                hit = !cull_miss(r0.x, r0.y, r1.z, r1.w, wx0, wy0, 8.0f, 4.0f) && (k0 + lane > pos0 - wmax);
            }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (hit) hl[nh + __popc(m & ((1u << lane) - 1u))] = (unsigned short)(k0 + lane);
            nh += __popc(m);
        }
        __syncwarp();
        if ((nh & 1) && lane == 0) hl[nh] = hl[nh - 1];  // pad to whole pairs (the copy is never replayed)
        __syncwarp();
        {
            // alpha of a pair (alpha_blending.cu:190-203); the pos < lc test (:185-187) is slot j > jmin
            const int jmin = pos0 - lc;
            auto pair_eval = [&](int j, float& Graw, float& araw) -> bool {
                const float4 r0 = st.rec[2 * j];
                const float4 r1 = st.rec[2 * j + 1];
                const float dx = fadd(r0.x, -pxf), dy = fadd(r0.y, -pyf);
                const float power = pair_power(dx, dy, r0.z, r0.w, r1.x);
                Graw = ex2_approx(fmul(power, kLog2e));
                araw = fmin_ftz(fmul(r1.y, Graw), kAlphaMax);
                return (j > jmin) && !(power > 0.0f) && !(araw < kAlphaMin);
            };
            // sequential replay step of one visit; parks (X, w) for phase 2
            auto replay = [&](int j, bool valid, float Graw, float araw) {
                const float alpha = valid ? araw : 0.0f;
                const float G = valid ? Graw : 0.0f;
                const float rinv = rcp_approx(1.0f - alpha);  // == 1 for a failing pair
                Tn = Tn * rinv;                                // :205, kept negated: Tn = -T
                const float wgt = fmul(-alpha, Tn);            // alpha T
                // channel math on packed pairs (FFMA2/FMUL2): per pair of channels
                //   e = f T - S / (1 - alpha);  dL_dalpha += e . dpix;  S += f alpha T
                // computed as -e = S rinv + f Tn so that no operand has to be negated (the packed ops take no
                // negation modifier); the sign returns for free in the scalar FFMA that adds the background term
                const f32x2 TN2 = pk2(Tn, Tn), W2 = pk2(wgt, wgt), R2 = pk2(rinv, rinv);
                f32x2 dacc;
#pragma unroll
                for (int k = 0; k < CH; k += 4) {
                    const float4 fv = *reinterpret_cast<const float4*>(&st.feat[j * CH + k]);
                    const f32x2 fa = pk2(fv.x, fv.y), fb = pk2(fv.z, fv.w);
                    const f32x2 ea = fma2(fa, TN2, mul2(S2[k / 2], R2));           // :213-217
                    const f32x2 eb = fma2(fb, TN2, mul2(S2[k / 2 + 1], R2));
                    dacc = k == 0 ? mul2(ea, dpix2[0]) : fma2(ea, dpix2[k / 2], dacc);
                    dacc = fma2(eb, dpix2[k / 2 + 1], dacc);
                    S2[k / 2] = fma2(fa, W2, S2[k / 2]);
                    S2[k / 2 + 1] = fma2(fb, W2, S2[k / 2 + 1]);
                }
                float dlo, dhi;
                upk2(dacc, dlo, dhi);
                // X = G dL/dalpha, dL/dalpha = e . dpix + nbg / (1 - alpha)   (:222-229: background term)
                *xw_wr = make_float2(G * fmaf(nbg, rinv, -(dlo + dhi)), wgt);  // (X, w) of this pair
                if (lane == 0) xw_wr[32].x = __int_as_float(j);
                xw_wr += BW_PS;
                if (++cnt == BW_GQ) {
                    bwd_reduce_group<CH, B>(BW_GQ, lane, st, sid, xw, dpw, wx0, wy0, grec, gfeat, fstride, foff,
                                            geom_grads);
                    cnt = 0;
                    xw_wr = xw + lane;
                }
            };
            // Two hits per iteration: their alpha evaluations (LDS -> 7 dependent FP32 ops -> MUFU.EX2 -> min ->
            // compare -> vote) are independent and interleave; only the replay steps are sequential.
            // The list entry is loaded one pair ahead (the kernel is latency-sensitive: 7.65 -> 7.50 ms per 8-view step);
            // the read past the end stays inside the warp's list (HL_STRIDE = B + 2).
            const unsigned* hp = reinterpret_cast<const unsigned*>(hl);
            unsigned jnext = *hp;
            for (int rem = nh; rem > 0; rem -= 2) {
                const unsigned jj = jnext;
                jnext = *++hp;
                const int j1 = (int)(jj & 0xffffu), j2 = (int)(jj >> 16);
                const bool two = rem > 1;
                float G1, a1, G2, a2;
                const bool v1 = pair_eval(j1, G1, a1);
                const bool v2 = pair_eval(j2, G2, a2) && two;
                const bool any1 = __any_sync(0xffffffffu, v1);
                const bool any2 = __any_sync(0xffffffffu, v2);
                if (any1) replay(j1, v1, G1, a1);
                if (any2) replay(j2, v2, G2, a2);
            }
        }
        if (cnt > 0) {  // the stage buffer is recycled after this batch
            bwd_reduce_group<CH, B>(cnt, lane, st, sid, xw, dpw, wx0, wy0, grec, gfeat, fstride, foff, geom_grads);
            cnt = 0;
            xw_wr = xw + lane;
        }
    }
    cp_async_wait<0>();
}

static inline int pick_fwd_ch(int rem) { return rem >= 32 ? 32 : rem > 8 ? 16 : rem > 4 ? 8 : 4; }
static inline int pick_bwd_ch(int rem) { return rem >= 16 ? 16 : rem > 4 ? 8 : 4; }

template <int CH>
static int launch_fwd(dim3 grid, cudaStream_t st, const float4* rec, const float* featp, int fstride, int foff,
                      const int* ids, const int2* tr, float bg, int c_valid, int W, int H, int write_aux,
                      float* final_T, int* ncontrib, float* image, long long img_vstride) {
    const size_t smem = 2 * sizeof(Stage<CH>);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(blend_fwd_kernel<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return set_error((int)e, "alpha_blending_fwd: cudaFuncSetAttribute failed");
    }
    blend_fwd_kernel<CH><<<grid, BL_NT, smem, st>>>(rec, featp, fstride, foff, ids, tr, bg, c_valid, W, H, write_aux,
                                                    final_T, ncontrib, image, img_vstride);
    return check_launch("alpha_blending_fwd");
}

// A/B switch for profiling runs (CH = 4), MSB_BWD_CFG.  Measured on BASELINE config #3, ms per view
// (profiles/r2_ab_experiments.md; round 1 in brackets):
//   default  256-entry batches, registers capped for 3 CTAs/SM (80)            0.84  (1.054)
//   1        128-entry batches, 3 CTAs/SM                                             (1.11)
//   2        256-entry batches, uncapped (74 registers, 3 CTAs/SM)             0.91  (1.19 at 93 registers)
//   3        128-entry batches, 64 registers (4 CTAs/SM, spills)               0.91  (1.15)
// Fewer barriers per list entry beat the extra resident CTA.
static int blend_bwd_cfg() {
    static const int v = [] { const char* e = getenv("MSB_BWD_CFG"); return e ? atoi(e) : 0; }();
    return v;
}

template <int CH, int B, int MINB>
static int launch_bwd_cfg(dim3 grid, cudaStream_t st, const float4* rec, const float* featp, int fstride, int foff,
                          const int* ids, const int2* tr, float bg, int c_valid, int W, int H, const float* final_T,
                          const int* ncontrib, const float* dL_dimage, long long img_vstride, float* grec, float* gfeat,
                          int geom) {
    const size_t smem = Bwd3<CH, B>::SMEM;
    if (smem > 40 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(blend_bwd_kernel<CH, B, MINB>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error((int)e, "alpha_blending_bwd: cudaFuncSetAttribute failed");
    }
    blend_bwd_kernel<CH, B, MINB><<<grid, BL_NT, smem, st>>>(rec, featp, fstride, foff, ids, tr, bg, c_valid, W,
                                                                  H, final_T, ncontrib, dL_dimage, img_vstride, grec,
                                                                  gfeat, geom);
    return check_launch("alpha_blending_bwd");
}

template <int CH>
static int launch_bwd(dim3 grid, cudaStream_t st, const float4* rec, const float* featp, int fstride, int foff,
                      const int* ids, const int2* tr, float bg, int c_valid, int W, int H, const float* final_T,
                      const int* ncontrib, const float* dL_dimage, long long img_vstride, float* grec, float* gfeat,
                      int geom) {
#define MSB_BWD_ARGS grid, st, rec, featp, fstride, foff, ids, tr, bg, c_valid, W, H, final_T, ncontrib, dL_dimage, img_vstride, grec, gfeat, geom
    if constexpr (CH == 4) {
        switch (blend_bwd_cfg()) {
            case 1: return launch_bwd_cfg<CH, 128, 3>(MSB_BWD_ARGS);
            case 2: return launch_bwd_cfg<CH, 256, 0>(MSB_BWD_ARGS);
            case 3: return launch_bwd_cfg<CH, 128, 4>(MSB_BWD_ARGS);
            default: return launch_bwd_cfg<CH, 256, 3>(MSB_BWD_ARGS);
        }
    } else {
        // CH = 16: registers capped at 128 so that two CTAs fit an SM (143 uncapped -> one); CH = 8 is limited to
        // two CTAs by shared memory either way
        return launch_bwd_cfg<CH, 256, (CH == 16 ? 2 : 0)>(MSB_BWD_ARGS);
    }
#undef MSB_BWD_ARGS
}

// channel-chunk dispatcher of the forward pass (reference D1: alpha_blending.cu:248-394)
// views > 1: one grid for a whole view batch (blockIdx.z = view), image [views, C, H, W]
static int run_fwd_passes(cudaStream_t st, const float4* rec, const float* fsrc, int Cpad, int C,
                          const int32_t* idx_sorted, const int32_t* tile_range, float bg, int W, int H, float* image,
                          float* final_T, int32_t* ncontrib, int views = 1) {
    const dim3 grid((W + MSB_TILE - 1) / MSB_TILE, (H + MSB_TILE - 1) / MSB_TILE, views);
    const long long vs = (long long)C * H * W;
    const int2* tr = reinterpret_cast<const int2*>(tile_range);
    int c0 = 0, first = 1;
    do {  // at least one pass so that final_T / ncontrib exist even for C == 0
        const int rem = Cpad - c0;
        const int ch = C == 0 ? 4 : pick_fwd_ch(rem);
        const int c_valid = max(0, min(ch, C - c0));
        float* img = image ? image + (size_t)c0 * H * W : nullptr;
        int rc;
        if (C == 0) {  // geometry-only pass: stage rec twice (no feature rows exist)
            rc = launch_fwd<4>(grid, st, rec, reinterpret_cast<const float*>(rec), 8, 0, idx_sorted, tr, bg, 0, W, H, 1,
                               final_T, ncontrib, img, vs);
        } else if (ch == 32) {
            rc = launch_fwd<32>(grid, st, rec, fsrc, Cpad, c0, idx_sorted, tr, bg, c_valid, W, H, first, final_T,
                                ncontrib, img, vs);
        } else if (ch == 16) {
            rc = launch_fwd<16>(grid, st, rec, fsrc, Cpad, c0, idx_sorted, tr, bg, c_valid, W, H, first, final_T,
                                ncontrib, img, vs);
        } else if (ch == 8) {
            rc = launch_fwd<8>(grid, st, rec, fsrc, Cpad, c0, idx_sorted, tr, bg, c_valid, W, H, first, final_T,
                               ncontrib, img, vs);
        } else {
            rc = launch_fwd<4>(grid, st, rec, fsrc, Cpad, c0, idx_sorted, tr, bg, c_valid, W, H, first, final_T,
                               ncontrib, img, vs);
        }
        if (rc) return rc;
        c0 += ch;
        first = 0;
    } while (c0 < C);
    return MSB_OK;
}

// channel-chunk dispatcher of the backward pass; grec/gfeat must be zero on entry
static int run_bwd_passes(cudaStream_t st, const float4* rec, const float* fsrc, int Cpad, int C,
                          const int32_t* idx_sorted, const int32_t* tile_range, float bg, int W, int H,
                          const float* final_T, const int32_t* ncontrib, const float* dL_dimage, float* grec,
                          float* gfeat, int views = 1) {
    const dim3 grid((W + MSB_TILE - 1) / MSB_TILE, (H + MSB_TILE - 1) / MSB_TILE, views);
    const long long vs = (long long)C * H * W;
    const int2* tr = reinterpret_cast<const int2*>(tile_range);
    for (int c0 = 0; c0 < C;) {
        const int rem = Cpad - c0;
        const int ch = pick_bwd_ch(rem);
        const int c_valid = min(ch, C - c0);
        const float* dimg = dL_dimage + (size_t)c0 * H * W;
        int rc;
        // geometric gradients are linear in the channels: every chunk adds its share (D1 in the
        // reference does the same, alpha_blending.cu:436-567)
        if (ch == 16)
            rc = launch_bwd<16>(grid, st, rec, fsrc, Cpad, c0, idx_sorted, tr, bg, c_valid, W, H, final_T, ncontrib,
                                dimg, vs, grec, gfeat, 1);
        else if (ch == 8)
            rc = launch_bwd<8>(grid, st, rec, fsrc, Cpad, c0, idx_sorted, tr, bg, c_valid, W, H, final_T, ncontrib,
                               dimg, vs, grec, gfeat, 1);
        else
            rc = launch_bwd<4>(grid, st, rec, fsrc, Cpad, c0, idx_sorted, tr, bg, c_valid, W, H, final_T, ncontrib,
                               dimg, vs, grec, gfeat, 1);
        if (rc) return rc;
        c0 += ch;
    }
    return MSB_OK;
}

}  // namespace msb

using namespace msb;

extern "C" {

// Padded feature-row width used by the packed layout.
int msb_blend_cpad(int C) { return C <= 4 ? 4 : C <= 8 ? 8 : (C + 15) / 16 * 16; }

// Workspace: rec [P][8] floats, then (if C != Cpad) featp [P][Cpad] floats.
size_t msb_blend_fwd_workspace_bytes(int P, int C) {
    const int Cpad = msb_blend_cpad(C);
    size_t b = (size_t)P * 8 * sizeof(float);
    if (C != Cpad) b += (size_t)P * Cpad * sizeof(float);
    return b + 256;
}

// Backward workspace: grec [P][8] + gfeat [P][Cpad] (zeroed by the call).
size_t msb_blend_bwd_workspace_bytes(int P, int C) {
    return (size_t)P * (8 + msb_blend_cpad(C)) * sizeof(float) + 256;
}

// Forward.  packed (workspace, msb_blend_fwd_workspace_bytes) is an OUTPUT that the backward
// pass reuses (the caller keeps it alive with the autograd context).
int msb_alpha_blending_fwd(const float* uv, const float* conic, const float* opacity, const float* feature,
                           const int32_t* idx_sorted, const int32_t* tile_range, float bg, int P, int C, int W,
                           int H, float* image, float* final_T, int32_t* ncontrib, void* packed,
                           size_t packed_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (P < 0 || C < 0 || W <= 0 || H <= 0 || !tile_range || !final_T || !ncontrib)
        return set_error(MSB_ERR_ARG, "alpha_blending_fwd: bad argument");
    if (C > 0 && !image) return set_error(MSB_ERR_ARG, "alpha_blending_fwd: null image");
    if (P > 0 && (!uv || !conic || !opacity || !packed || (C > 0 && !feature)))
        return set_error(MSB_ERR_ARG, "alpha_blending_fwd: null pointer");
    if (packed_bytes < msb_blend_fwd_workspace_bytes(P, C))
        return set_error(MSB_ERR_WORKSPACE, "alpha_blending_fwd: workspace too small");
    const int Cpad = msb_blend_cpad(C);
    float4* rec = reinterpret_cast<float4*>(packed);
    float* featp = (C != Cpad) ? reinterpret_cast<float*>(packed) + (size_t)P * 8 : nullptr;
    if ((reinterpret_cast<uintptr_t>(packed) & 15u) || (C == Cpad && (reinterpret_cast<uintptr_t>(feature) & 15u)))
        return set_error(MSB_ERR_ARG, "alpha_blending_fwd: 16-byte alignment");
    if (P > 0) {
        blend_pack_kernel<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(P, C, Cpad, reinterpret_cast<const float2*>(uv),
                                                                      conic, opacity, feature, rec, featp);
        int rc = check_launch("alpha_blending_fwd/pack");
        if (rc) return rc;
    }
    const float* fsrc = (C != Cpad) ? featp : feature;
    return run_fwd_passes(st, rec, fsrc, Cpad, C, idx_sorted, tile_range, bg, W, H, image, final_T, ncontrib);
}

// Pack only: rebuilds the `packed` buffer of msb_alpha_blending_fwd from the same inputs.  For callers
// whose backward entry does not receive the forward's workspace (msplat._C.alpha_blending_backward takes
// uv / conic / opacity / feature again: integration/_C.py).
int msb_blend_pack(const float* uv, const float* conic, const float* opacity, const float* feature, int P, int C,
                   void* packed, size_t packed_bytes, void* stream) {
    if (P < 0 || C < 0) return set_error(MSB_ERR_ARG, "blend_pack: bad argument");
    if (P == 0) return MSB_OK;
    if (!uv || !conic || !opacity || !packed || (C > 0 && !feature)) return set_error(MSB_ERR_ARG, "blend_pack: null pointer");
    if (packed_bytes < msb_blend_fwd_workspace_bytes(P, C)) return set_error(MSB_ERR_WORKSPACE, "blend_pack: workspace too small");
    if (reinterpret_cast<uintptr_t>(packed) & 15u) return set_error(MSB_ERR_ARG, "blend_pack: 16-byte alignment");
    const int Cpad = msb_blend_cpad(C);
    float4* rec = reinterpret_cast<float4*>(packed);
    float* featp = (C != Cpad) ? reinterpret_cast<float*>(packed) + (size_t)P * 8 : nullptr;
    blend_pack_kernel<<<(unsigned)((P + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        P, C, Cpad, reinterpret_cast<const float2*>(uv), conic, opacity, feature, rec, featp);
    return check_launch("blend_pack");
}

// Forward on inputs that are already in the packed layout (rec [P,8], featp [P,Cpad] with
// Cpad = msb_blend_cpad(C)): what msb_render_preprocess_fwd produces.
// views > 1: a view batch in one grid.  rec [views*P,8], featp [views*P,Cpad], idx_sorted and
// tile_range [views*T,2] from msb_sort_gaussian_views; image [views,C,H,W], final_T / ncontrib [views,H,W].
int msb_blend_packed_fwd_views(const float* rec, const float* featp, const int32_t* idx_sorted,
                               const int32_t* tile_range, float bg, int C, int W, int H, int views, float* image,
                               float* final_T, int32_t* ncontrib, void* stream) {
    // rec / featp may be NULL for an empty cloud (every tile range is then (0, 0))
    if (C < 0 || W <= 0 || H <= 0 || views <= 0 || views > 65535 || !tile_range || !final_T || !ncontrib ||
        (C > 0 && !image))
        return set_error(MSB_ERR_ARG, "blend_packed_fwd: bad argument");
    if ((reinterpret_cast<uintptr_t>(rec) | reinterpret_cast<uintptr_t>(featp)) & 15u)
        return set_error(MSB_ERR_ARG, "blend_packed_fwd: 16-byte alignment");
    return run_fwd_passes((cudaStream_t)stream, reinterpret_cast<const float4*>(rec), featp, msb_blend_cpad(C), C,
                          idx_sorted, tile_range, bg, W, H, image, final_T, ncontrib, views);
}
// Diagnostic (bench.py): blended [views,H,W] int32 = list entries that blend at each pixel.
int msb_blend_packed_count(const float* rec, const int32_t* idx_sorted, const int32_t* tile_range, int W, int H,
                           int views, int32_t* blended, void* stream) {
    if (W <= 0 || H <= 0 || views <= 0 || views > 65535 || !tile_range || !blended)
        return set_error(MSB_ERR_ARG, "blend_packed_count: bad argument");
    const dim3 grid((W + MSB_TILE - 1) / MSB_TILE, (H + MSB_TILE - 1) / MSB_TILE, views);
    // features are not needed: the records are staged twice (as in the C == 0 pass of run_fwd_passes)
    blend_fwd_kernel<4, true><<<grid, BL_NT, 2 * sizeof(Stage<4>), (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(rec), rec, 8, 0, idx_sorted, reinterpret_cast<const int2*>(tile_range), 0.f, 0,
        W, H, 0, nullptr, nullptr, reinterpret_cast<float*>(blended), 0);
    return check_launch("blend_packed_count");
}

int msb_blend_packed_fwd(const float* rec, const float* featp, const int32_t* idx_sorted, const int32_t* tile_range,
                         float bg, int C, int W, int H, float* image, float* final_T, int32_t* ncontrib,
                         void* stream) {
    return msb_blend_packed_fwd_views(rec, featp, idx_sorted, tile_range, bg, C, W, H, 1, image, final_T, ncontrib,
                                      stream);
}

// Backward.  `packed` is the buffer produced by the forward call on the same inputs.
int msb_alpha_blending_bwd(const float* feature, const int32_t* idx_sorted, const int32_t* tile_range, float bg,
                           int P, int C, int W, int H, const float* final_T, const int32_t* ncontrib,
                           const float* dL_dimage, const void* packed, float* dL_duv, float* dL_dconic,
                           float* dL_dopacity, float* dL_dfeature, void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (P < 0 || C < 0 || W <= 0 || H <= 0) return set_error(MSB_ERR_ARG, "alpha_blending_bwd: bad argument");
    if (P == 0) return MSB_OK;
    if (!tile_range || !final_T || !ncontrib || !packed || !dL_duv || !dL_dconic || !dL_dopacity || !ws ||
        (C > 0 && (!dL_dimage || !dL_dfeature || !feature)))
        return set_error(MSB_ERR_ARG, "alpha_blending_bwd: null pointer");
    if (ws_bytes < msb_blend_bwd_workspace_bytes(P, C))
        return set_error(MSB_ERR_WORKSPACE, "alpha_blending_bwd: workspace too small");
    const int Cpad = msb_blend_cpad(C);
    const float4* rec = reinterpret_cast<const float4*>(packed);
    const float* fsrc = (C != Cpad) ? reinterpret_cast<const float*>(packed) + (size_t)P * 8 : feature;
    float* grec = reinterpret_cast<float*>(ws);
    float* gfeat = grec + (size_t)P * 8;
    cudaError_t e = cudaMemsetAsync(ws, 0, (size_t)P * (8 + Cpad) * sizeof(float), st);
    if (e != cudaSuccess) return set_error((int)e, "alpha_blending_bwd: memset failed");
    int rc = run_bwd_passes(st, rec, fsrc, Cpad, C, idx_sorted, tile_range, bg, W, H, final_T, ncontrib, dL_dimage, grec,
                            gfeat);
    if (rc) return rc;
    blend_unpack_kernel<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(P, C, Cpad, grec, gfeat, dL_duv, dL_dconic,
                                                                    dL_dopacity, dL_dfeature);
    return check_launch("alpha_blending_bwd/unpack");
}

// Backward on packed inputs; the gradients stay packed: grec [P,8] = {dL_duv.xy, dL_dconic.xyz,
// dL_dopacity, 0, 0}, gfeat [P,Cpad].  Both are zeroed by this call and consumed by
// msb_render_preprocess_bwd.
int msb_blend_packed_bwd_views(const float* rec, const float* featp, const int32_t* idx_sorted,
                               const int32_t* tile_range, float bg, int P, int C, int W, int H, int views,
                               const float* final_T, const int32_t* ncontrib, const float* dL_dimage, float* grec,
                               float* gfeat, int already_zero, void* stream);
int msb_blend_packed_bwd(const float* rec, const float* featp, const int32_t* idx_sorted, const int32_t* tile_range,
                         float bg, int P, int C, int W, int H, const float* final_T, const int32_t* ncontrib,
                         const float* dL_dimage, float* grec, float* gfeat, int already_zero, void* stream) {
    return msb_blend_packed_bwd_views(rec, featp, idx_sorted, tile_range, bg, P, C, W, H, 1, final_T, ncontrib,
                                      dL_dimage, grec, gfeat, already_zero, stream);
}

// View batch (see msb_blend_packed_fwd_views): P = Gaussians per view; grec [views*P,8], gfeat [views*P,Cpad],
// dL_dimage [views,C,H,W].
int msb_blend_packed_bwd_views(const float* rec, const float* featp, const int32_t* idx_sorted,
                               const int32_t* tile_range, float bg, int P, int C, int W, int H, int views,
                               const float* final_T, const int32_t* ncontrib, const float* dL_dimage, float* grec,
                               float* gfeat, int already_zero, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (P < 0 || C < 0 || W <= 0 || H <= 0 || views <= 0 || views > 65535)
        return set_error(MSB_ERR_ARG, "blend_packed_bwd: bad argument");
    if (P == 0) return MSB_OK;
    if (!rec || !featp || !tile_range || !final_T || !ncontrib || !grec || !gfeat || (C > 0 && !dL_dimage))
        return set_error(MSB_ERR_ARG, "blend_packed_bwd: null pointer");
    const int Cpad = msb_blend_cpad(C);
    if (!already_zero) {  // the kernels accumulate with reductions: the packed gradients start from zero
        cudaError_t e = cudaMemsetAsync(grec, 0, (size_t)views * P * 8 * sizeof(float), st);
        if (e == cudaSuccess) e = cudaMemsetAsync(gfeat, 0, (size_t)views * P * Cpad * sizeof(float), st);
        if (e != cudaSuccess) return set_error((int)e, "blend_packed_bwd: memset failed");
    }
    return run_bwd_passes(st, reinterpret_cast<const float4*>(rec), featp, Cpad, C, idx_sorted, tile_range, bg, W, H,
                          final_T, ncontrib, dL_dimage, grec, gfeat, views);
}

}  // extern "C"
