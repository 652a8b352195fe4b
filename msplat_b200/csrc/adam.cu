// msplat_b200/csrc/adam.cu -- fused multi-tensor Adam step.
//
// The reference's training loop (/root/reference/tutorials/gs_2d.py:32-36,66-87) updates its five parameter
// tensors (xyz, scale, rotate, opacity, rgb) with torch.optim.Adam, i.e. a handful of elementwise kernels per
// tensor and step.  Here ALL tensors of a parameter group are updated by ONE launch: a block looks up which
// tensor its element range belongs to (<= 8 tensors per launch, their pointers and sizes travel as kernel
// arguments) and applies torch.optim.Adam's update (no weight decay, no amsgrad) in the same operation order:
//     m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2
//     p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// HBM-bound: 16 B read + 12 B written per element, 16-byte vector accesses.
#include "common.cuh"

namespace msb {

constexpr int ADAM_MAX_TENSORS = 8;
constexpr int ADAM_NT = 256;
constexpr int ADAM_VEC_PER_THREAD = 4;                              // float4 per thread
constexpr int ADAM_CHUNK = ADAM_NT * ADAM_VEC_PER_THREAD * 4;       // elements per block

struct AdamArgs {
    float* p[ADAM_MAX_TENSORS];
    const float* g[ADAM_MAX_TENSORS];
    float* m[ADAM_MAX_TENSORS];
    float* v[ADAM_MAX_TENSORS];
    long long n[ADAM_MAX_TENSORS];
    long long first_block[ADAM_MAX_TENSORS + 1];  // blocks are assigned tensor by tensor
    int count;
};

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, float b1, float b2, float step_size,
                                          float inv_bc2_sqrt, float eps) {
    m = fmaf(b1, m, (1.0f - b1) * g);
    v = fmaf(b2, v, (1.0f - b2) * g * g);
    const float denom = sqrtf(v) * inv_bc2_sqrt + eps;
    p -= step_size * (m / denom);
}

__global__ void __launch_bounds__(ADAM_NT) adam_kernel(AdamArgs a, float b1, float b2, float step_size,
                                                       float inv_bc2_sqrt, float eps) {
    int t = 0;
#pragma unroll
    for (int k = 1; k < ADAM_MAX_TENSORS; ++k)
        if (k < a.count && (long long)blockIdx.x >= a.first_block[k]) t = k;
    const long long base = ((long long)blockIdx.x - a.first_block[t]) * ADAM_CHUNK;
    float* __restrict__ p = a.p[t];
    const float* __restrict__ g = a.g[t];
    float* __restrict__ m = a.m[t];
    float* __restrict__ v = a.v[t];
    const long long n = a.n[t];
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 15u) == 0;
#pragma unroll
    for (int k = 0; k < ADAM_VEC_PER_THREAD; ++k) {
        const long long i = base + ((long long)k * ADAM_NT + threadIdx.x) * 4;
        if (i >= n) break;
        if (vec && i + 3 < n) {
            float4 pp = *reinterpret_cast<float4*>(p + i);
            const float4 gg = ldg_stream4(reinterpret_cast<const float4*>(g + i));
            float4 mm = *reinterpret_cast<float4*>(m + i);
            float4 vv = *reinterpret_cast<float4*>(v + i);
            adam_elem(pp.x, gg.x, mm.x, vv.x, b1, b2, step_size, inv_bc2_sqrt, eps);
            adam_elem(pp.y, gg.y, mm.y, vv.y, b1, b2, step_size, inv_bc2_sqrt, eps);
            adam_elem(pp.z, gg.z, mm.z, vv.z, b1, b2, step_size, inv_bc2_sqrt, eps);
            adam_elem(pp.w, gg.w, mm.w, vv.w, b1, b2, step_size, inv_bc2_sqrt, eps);
            *reinterpret_cast<float4*>(p + i) = pp;
            *reinterpret_cast<float4*>(m + i) = mm;
            *reinterpret_cast<float4*>(v + i) = vv;
        } else {
            for (long long j = i; j < min(i + 4, n); ++j) {
                float pp = p[j], mm = m[j], vv = v[j];
                adam_elem(pp, g[j], mm, vv, b1, b2, step_size, inv_bc2_sqrt, eps);
                p[j] = pp;
                m[j] = mm;
                v[j] = vv;
            }
        }
    }
}

}  // namespace msb

using namespace msb;

extern "C" {

// One Adam step (torch.optim.Adam semantics, no weight decay / amsgrad) for `ntensors` float32 tensors in one
// launch per 8 tensors.  params / grads / exp_avg / exp_avg_sq are HOST arrays of device pointers, numel a host
// array of element counts; `step` is the 1-based step count used for the bias corrections.
int msb_adam_step(int ntensors, float* const* params, const float* const* grads, float* const* exp_avg,
                  float* const* exp_avg_sq, const long long* numel, float lr, float beta1, float beta2, float eps,
                  int step, void* stream) {
    if (ntensors < 0 || step < 1 || (ntensors > 0 && (!params || !grads || !exp_avg || !exp_avg_sq || !numel)))
        return set_error(MSB_ERR_ARG, "adam_step: bad argument");
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    const float step_size = (float)((double)lr / bc1);
    const float inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
    for (int t0 = 0; t0 < ntensors; t0 += ADAM_MAX_TENSORS) {
        AdamArgs a;
        a.count = 0;
        long long blocks = 0;
        for (int t = t0; t < ntensors && a.count < ADAM_MAX_TENSORS; ++t) {
            if (numel[t] < 0) return set_error(MSB_ERR_ARG, "adam_step: negative size");
            if (numel[t] == 0) continue;
            if (!params[t] || !grads[t] || !exp_avg[t] || !exp_avg_sq[t])
                return set_error(MSB_ERR_ARG, "adam_step: null pointer");
            const int k = a.count++;
            a.p[k] = params[t];
            a.g[k] = grads[t];
            a.m[k] = exp_avg[t];
            a.v[k] = exp_avg_sq[t];
            a.n[k] = numel[t];
            a.first_block[k] = blocks;
            blocks += (numel[t] + ADAM_CHUNK - 1) / ADAM_CHUNK;
        }
        if (a.count == 0) continue;
        a.first_block[a.count] = blocks;
        if (blocks > 0x7fffffffll) return set_error(MSB_ERR_RANGE, "adam_step: too many elements for one launch");
        adam_kernel<<<(unsigned)blocks, ADAM_NT, 0, (cudaStream_t)stream>>>(a, beta1, beta2, step_size, inv_bc2_sqrt, eps);
        int rc = check_launch("adam_step");
        if (rc) return rc;
    }
    return MSB_OK;
}

}  // extern "C"
