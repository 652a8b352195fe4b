// msplat_b200/csrc/blend_math.cuh -- per (pixel, Gaussian) pair arithmetic of the tile blender.
//
// The skip / terminate decisions of a pair (power > 0, alpha < 1/255, T' < 1e-4) are
// discontinuities: a 1-ulp difference in `power` or `alpha` flips a pair in or out and moves a
// pixel by up to alpha*feature ~ 4e-3.  To stay inside the 1e-4 image tolerance of the contract
// the forward/backward pair math therefore mirrors, operation for operation, what the
// reference's `-O3 --use_fast_math` sm_100 build executes for
//   /root/reference/msplat/src/alpha_blending.cu:76-100 (forward) and :190-205 (backward):
//     power = fma(fma(dx, dx*cx, dy*(dy*cz)), -0.5, -(dy*(dx*cy)))
//     G     = ex2.approx(power * 1.4426950216)          (expf under fast-math)
//     alpha = min(opacity * G, 0.99)
//     skip if alpha < float(1/255) = 0x3B808081        (the reference's double compare against
//                                                        1.0/255.0f selects exactly the same floats)
//     T'    = T * (1 - alpha);  terminate if T' < 1e-4
//     F_k   = fma(T, alpha * f_k, F_k)
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace msb {

constexpr float kLog2e = 1.4426950216293334961f;   // 0x3FB8AA3B, the constant in the reference SASS
constexpr float kAlphaMin = 0.0039215688593685627f;  // 0x3B808081
constexpr float kAlphaMax = 0.99f;
constexpr float kTmin = 0.0001f;

MSB_HD float inf_f() {
#ifdef __CUDA_ARCH__
    return __int_as_float(0x7f800000);
#else
    return INFINITY;
#endif
}

MSB_HD float pair_power(float dx, float dy, float cx, float cy, float cz) {
    const float q = ffma(dx, fmul(dx, cx), fmul(dy, fmul(dy, cz)));
    return ffma(q, -0.5f, -fmul(dy, fmul(dx, cy)));
}

// Returns true if the pair passes the power/alpha tests; G and alpha are outputs.
MSB_HD bool pair_alpha(float power, float opacity, float& G, float& alpha) {
    if (power > 0.0f) return false;
    G = ex2_approx(fmul(power, kLog2e));
    alpha = fmin_ftz(fmul(opacity, G), kAlphaMax);
    return !(alpha < kAlphaMin);
}

// Conservative extents (in pixels) of the region where a Gaussian can pass the alpha test:
// alpha >= 1/255  =>  q(d) = cx dx^2 + 2 cy dx dy + cz dy^2 <= 2 ln(255 opacity) =: tau.  The ellipse
// q <= tau has the support function sqrt(tau a^T Q^-1 a) along a; four directions give an octagon:
//   |dx| <= hx = sqrt(tau cz / det)            |dy| <= hy = sqrt(tau cx / det)
//   |dx + dy| <= hs = sqrt(tau (cx + cz - 2 cy) / det)     |dx - dy| <= ht = sqrt(tau (cx + cz + 2 cy) / det)
// (the two diagonals cut the corners of the box: ~12 % fewer warp-visits on BASELINE config #3).
// Slack: tau is inflated by 0.05 + 0.1 % (FP32 error of `power` incl. cancellation for
// anisotropic conics, ex2.approx error), the box extents by 0.01 px + 0.01 %, the diagonal ones by
// 0.02 px + 0.3 % (cx + cz -+ 2 cy cancels down to 5e-5 (cx + cz) at the conditioning limit below).
// Ill-conditioned or non-positive-definite conics, and NaNs, disable culling (+inf).  Opacities
// that can never reach 1/255 give -inf (always culled).  Purely an optimisation: every surviving
// pair still runs the exact tests above.
MSB_HD void cull_extent(float cx, float cy, float cz, float opacity, float& hx, float& hy, float& hs, float& ht) {
    const float inf = inf_f();
    if (!(opacity == opacity)) { hx = hy = hs = ht = inf; return; }
    if (opacity * 255.0f * 1.001f < 1.0f) { hx = hy = hs = ht = -inf; return; }
    const float det = cx * cz - cy * cy;
    if (!(cx > 0.0f) || !(cz > 0.0f) || !(det > 1e-4f * cx * cz)) { hx = hy = hs = ht = inf; return; }
    const float tau = 2.0f * logf(opacity * 255.0f) * 1.001f + 0.05f;
    const float k = tau / det;
    hx = sqrtf(fmaxf(k * cz, 0.0f)) * 1.0001f + 0.01f;
    hy = sqrtf(fmaxf(k * cx, 0.0f)) * 1.0001f + 0.01f;
    hs = sqrtf(fmaxf(k * (cx + cz - 2.0f * cy), 0.0f)) * 1.003f + 0.02f;
    ht = sqrtf(fmaxf(k * (cx + cz + 2.0f * cy), 0.0f)) * 1.003f + 0.02f;
}

// The four extents travel in the two spare floats of the 32-byte blend record as FP16 pairs,
// rounded UP (values beyond the FP16 range become +inf = "no culling").
MSB_HD void cull_pack(float hx, float hy, float hs, float ht, float& p0, float& p1) {
    const __half2 a = __halves2half2(__float2half_ru(hx), __float2half_ru(hy));
    const __half2 b = __halves2half2(__float2half_ru(hs), __float2half_ru(ht));
    p0 = *reinterpret_cast<const float*>(&a);
    p1 = *reinterpret_cast<const float*>(&b);
}
MSB_HD void cull_unpack(float p0, float p1, float& hx, float& hy, float& hs, float& ht) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&p0));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&p1));
    hx = a.x; hy = a.y; hs = b.x; ht = b.y;
}

// Footprint test of one Gaussian (centre u, v, packed extents p0, p1) against the pixel block
// [x0, x0 + w - 1] x [y0, y0 + h - 1] (integer pixel coordinates as floats): true = cannot touch it.
MSB_HD bool cull_miss(float u, float v, float p0, float p1, float x0, float y0, float w, float h) {
    float hx, hy, hs, ht;
    cull_unpack(p0, p1, hx, hy, hs, ht);
    const float rx = 0.5f * (w - 1.0f), ry = 0.5f * (h - 1.0f);
    const float du = u - (x0 + rx), dv = v - (y0 + ry);   // relative to the block centre
    // interval tests |c - centre| > extent + half range, for x, y, x + y and x - y
    return (fabsf(du) > hx + rx) || (fabsf(dv) > hy + ry) || (fabsf(du + dv) > hs + rx + ry) ||
           (fabsf(du - dv) > ht + rx + ry);
}

}  // namespace msb
