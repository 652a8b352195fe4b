// tools/host_check.cu -- TEST INFRASTRUCTURE: compiles the per-Gaussian / per-pair math headers
// of msplat_b200 (geom.cuh, sh_eval.cuh, blend_math.cuh) for the HOST so that the CPU-only test
// suite (-m "not gpu") can check the very same source against the oracle without a GPU.
// MUFU approximations become IEEE operations on the host, so these are tolerance-level checks.
// Never linked into libmsplat_b200.so; the product has no CPU path.
//
// Build: nvcc -std=c++17 -Xcompiler -fPIC,-ffp-contract=off -shared -I msplat_b200/csrc \
//             tools/host_check.cu -o tools/_build/libhost_check.so
#include "blend_math.cuh"
#include "geom.cuh"
#include "sh_eval.cuh"

using namespace msb;

template <int DEG>
static void sh_all(int P, const float* dirs, float* B) {
    constexpr int D = (DEG + 1) * (DEG + 1);
    for (int i = 0; i < P; ++i) sh_basis<DEG>(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2], B + (size_t)i * D, 1);
}
template <int DEG>
static void sh_grad_all(int P, const float* dirs, const float* w, float* g) {
    constexpr int D = (DEG + 1) * (DEG + 1);
    for (int i = 0; i < P; ++i)
        sh_basis_grad<DEG>(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2], w + (size_t)i * D, 1, g[3 * i], g[3 * i + 1],
                           g[3 * i + 2]);
}


extern "C" {

void hc_project_fwd(int P, const float* xyz, const float* intr, const float* extr, int W, int H, float nearest,
                    float extent, float* uv, float* depth) {
    const Cam c = load_cam(intr, extr);
    for (int i = 0; i < P; ++i) {
        float u, v, d;
        if (!project_fwd(c, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], W, H, nearest, extent, u, v, d)) u = v = d = 0.f;
        uv[2 * i] = u;
        uv[2 * i + 1] = v;
        depth[i] = d;
    }
}

void hc_project_bwd(int P, const float* xyz, const float* intr, const float* extr, const float* depth,
                    const float* guv, const float* gd, float* dxyz, float* cam16) {
    const Cam c = load_cam(intr, extr);
    for (int k = 0; k < 16; ++k) cam16[k] = 0.f;
    for (int i = 0; i < P; ++i) {
        float dx = 0.f, dy = 0.f, dz = 0.f;
        if (depth[i] != 0.f)
            project_bwd<true>(c, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], guv[2 * i], guv[2 * i + 1], gd[i], dx, dy,
                              dz, cam16);
        dxyz[3 * i] = dx;
        dxyz[3 * i + 1] = dy;
        dxyz[3 * i + 2] = dz;
    }
}

void hc_cov3d_fwd(int P, const float* s, const float* q, float* cov) {
    for (int i = 0; i < P; ++i)
        cov3d_fwd(s[3 * i], s[3 * i + 1], s[3 * i + 2], q[4 * i], q[4 * i + 1], q[4 * i + 2], q[4 * i + 3], cov + 6 * i);
}

void hc_cov3d_bwd(int P, const float* s, const float* q, const float* g, float* ds, float* dq) {
    for (int i = 0; i < P; ++i)
        cov3d_bwd(s[3 * i], s[3 * i + 1], s[3 * i + 2], q[4 * i], q[4 * i + 1], q[4 * i + 2], q[4 * i + 3], g + 6 * i,
                  ds + 3 * i, dq + 4 * i);
}

void hc_ewa_fwd(int P, const float* xyz, const float* cov, const float* intr, const float* extr, const float* uv,
                const unsigned char* visible, int W, int H, float* conic, int* radius, int* tiles) {
    const Cam c = load_cam(intr, extr);
    const int gx = (W + 15) / 16, gy = (H + 15) / 16;
    for (int i = 0; i < P; ++i) {
        float cx = 0.f, cy = 0.f, cz = 0.f;
        int r = 0, t = 0;
        if (visible[i] && !ewa_fwd(c, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], cov + 6 * i, uv[2 * i], uv[2 * i + 1],
                                   gx, gy, cx, cy, cz, r, t)) {
            cx = cy = cz = 0.f;
            r = t = 0;
        }
        conic[3 * i] = cx;
        conic[3 * i + 1] = cy;
        conic[3 * i + 2] = cz;
        radius[i] = r;
        tiles[i] = t;
    }
}

void hc_ewa_bwd(int P, const float* xyz, const float* cov, const float* intr, const float* extr, const int* radius,
                const float* gconic, float* dxyz, float* dcov, float* cam16) {
    const Cam c = load_cam(intr, extr);
    for (int k = 0; k < 16; ++k) cam16[k] = 0.f;
    for (int i = 0; i < P; ++i) {
        float dx = 0.f, dy = 0.f, dz = 0.f, dcv[6] = {0, 0, 0, 0, 0, 0};
        if (radius[i] > 0) {
            if (!ewa_bwd<true>(c, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], cov + 6 * i, gconic[3 * i],
                               gconic[3 * i + 1], gconic[3 * i + 2], dx, dy, dz, dcv, cam16)) {
                dx = dy = dz = 0.f;
                for (int k = 0; k < 6; ++k) dcv[k] = 0.f;
            }
        }
        dxyz[3 * i] = dx;
        dxyz[3 * i + 1] = dy;
        dxyz[3 * i + 2] = dz;
        for (int k = 0; k < 6; ++k) dcov[6 * i + k] = dcv[k];
    }
}

int hc_sh_basis(int deg, int P, const float* dirs, float* B) {
    switch (deg) {
#define CASE(d) case d: sh_all<d>(P, dirs, B); return 0;
        CASE(0) CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10)
#undef CASE
    }
    return -1;
}

int hc_sh_grad(int deg, int P, const float* dirs, const float* w, float* g) {
    switch (deg) {
#define CASE(d) case d: sh_grad_all<d>(P, dirs, w, g); return 0;
        CASE(0) CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10)
#undef CASE
    }
    return -1;
}

// one pixel blended over a list with the product's pair math (forward only): returns T, writes F
float hc_blend_pixel(int n, const float* uv, const float* conic, const float* opacity, const float* feat, int C,
                     float px, float py, float* F, int* last_out) {
    float T = 1.0f;
    int last = 0;
    for (int k = 0; k < C; ++k) F[k] = 0.f;
    for (int j = 0; j < n; ++j) {
        const float dx = fadd(uv[2 * j], -px), dy = fadd(uv[2 * j + 1], -py);
        const float power = pair_power(dx, dy, conic[3 * j], conic[3 * j + 1], conic[3 * j + 2]);
        float G, alpha;
        if (!pair_alpha(power, opacity[j], G, alpha)) continue;
        const float nT = fmul(T, fadd(-alpha, 1.0f));
        if (nT < kTmin) break;
        for (int k = 0; k < C; ++k) F[k] = ffma(T, fmul(alpha, feat[j * C + k]), F[k]);
        T = nT;
        last = j + 1;
    }
    *last_out = last;
    return T;
}

// extents after the FP16 round trip of the blend record: out [P,4] = hx, hy, hs, ht
void hc_cull_extent(int P, const float* conic, const float* opacity, float* out) {
    for (int i = 0; i < P; ++i) {
        float hx, hy, hs, ht, p0, p1;
        cull_extent(conic[3 * i], conic[3 * i + 1], conic[3 * i + 2], opacity[i], hx, hy, hs, ht);
        cull_pack(hx, hy, hs, ht, p0, p1);
        cull_unpack(p0, p1, out[4 * i], out[4 * i + 1], out[4 * i + 2], out[4 * i + 3]);
    }
}
// FP16 round trip of four extents: out >= in must hold for every finite or infinite input
void hc_cull_pack_roundtrip(int n, const float* in, float* out) {
    for (int i = 0; i < n; ++i) {
        float p0, p1;
        cull_pack(in[4 * i], in[4 * i + 1], in[4 * i + 2], in[4 * i + 3], p0, p1);
        cull_unpack(p0, p1, out[4 * i], out[4 * i + 1], out[4 * i + 2], out[4 * i + 3]);
    }
}
// block test used by the blend kernels
int hc_cull_miss(float u, float v, const float* conic, float opacity, float x0, float y0, float w, float h) {
    float hx, hy, hs, ht, p0, p1;
    cull_extent(conic[0], conic[1], conic[2], opacity, hx, hy, hs, ht);
    cull_pack(hx, hy, hs, ht, p0, p1);
    return cull_miss(u, v, p0, p1, x0, y0, w, h) ? 1 : 0;
}

}  // extern "C"
