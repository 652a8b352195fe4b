/*
 * oracle/blend_ref.c -- CPU restatement of the reference's tile blend loops.
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): loaded by tests/, smoke() and the
 * cpu_baseline leg of bench.py via ctypes; never by msplat_b200.
 *
 * Follows, statement by statement, the per-pixel semantics of
 *   forward : /root/reference/msplat/src/alpha_blending.cu:16-110
 *   backward: /root/reference/msplat/src/alpha_blending.cu:112-246
 * restated as plain scalar loops over pixels (no tiles-of-threads, no atomics).
 * Per-pair arithmetic is float32; per-Gaussian gradient sums are accumulated in double
 * (the reference accumulates with float atomics in a non-deterministic order, so only
 * tolerance-level agreement is defined for gradients).
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC blend_ref.c -o libblend_ref.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TILE 16 /* BLOCK_X == BLOCK_Y == 16: msplat/include/config.h:7-8 */

/* alpha_blending.cu:76-80 -- power of the 2-D Gaussian at pixel offset (dx, dy) */
static inline float pair_power(float cx, float cy, float cz, float dx, float dy) {
    return -0.5f * (cx * dx * dx + cz * dy * dy) - cy * dx * dy;
}

/*
 * Forward.  feature is [P, C] row-major (the Python-level layout, alpha_blending.py:7-18);
 * image is [C, H, W]; final_T [H, W]; ncontrib [H, W].
 * Returns the number of (pixel, list-entry) pairs traversed = sum(ncontrib).
 */
int64_t blend_forward_ref(int P, int C, int W, int H, const float *uv, const float *conic,
                          const float *opacity, const float *feature, const int32_t *idx_sorted,
                          const int32_t *tile_range, float bg, float *image, float *final_T,
                          int32_t *ncontrib) {
    (void)P;
    const int gx = (W + TILE - 1) / TILE;
    int64_t pairs = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : pairs)
    for (int y = 0; y < H; ++y) {
        float *F = (float *)malloc(sizeof(float) * (size_t)(C > 0 ? C : 1));
        for (int x = 0; x < W; ++x) {
            const int tile = (y / TILE) * gx + (x / TILE);
            const int r0 = tile_range[2 * tile], r1 = tile_range[2 * tile + 1];
            float T = 1.0f; /* :54 */
            uint32_t contributor = 0, last = 0;
            for (int k = 0; k < C; ++k) F[k] = 0.0f;
            for (int e = r0; e < r1; ++e) {
                contributor++; /* :75 */
                const int g = idx_sorted[e];
                const float dx = uv[2 * g] - (float)x, dy = uv[2 * g + 1] - (float)y;
                const float power = pair_power(conic[3 * g], conic[3 * g + 1], conic[3 * g + 2], dx, dy);
                if (power > 0.0f) continue;                            /* :82 */
                const float alpha = fminf(0.99f, opacity[g] * expf(power)); /* :85 */
                if (alpha < 1.0f / 255.0f) continue;                   /* :87 */
                const float next_T = T * (1.0f - alpha);
                if (next_T < 0.0001f) break; /* :90-94: pixel done, entry not blended */
                for (int k = 0; k < C; ++k) F[k] += feature[(size_t)g * C + k] * alpha * T; /* :96-97 */
                T = next_T;
                last = contributor; /* :99-100 */
            }
            const size_t pix = (size_t)y * W + x;
            final_T[pix] = T;
            ncontrib[pix] = (int32_t)last;
            for (int k = 0; k < C; ++k) image[(size_t)k * H * W + pix] = F[k] + T * bg; /* :104-109 */
            pairs += last;
        }
        free(F);
    }
    return pairs;
}

/*
 * Backward.  dL_dimage is [C, H, W].  Outputs (zero-initialised here):
 * dL_duv [P,2], dL_dconic [P,3], dL_dopacity [P], dL_dfeature [P,C] (contiguous).
 */
void blend_backward_ref(int P, int C, int W, int H, const float *uv, const float *conic,
                        const float *opacity, const float *feature, const int32_t *idx_sorted,
                        const int32_t *tile_range, float bg, const float *final_T,
                        const int32_t *ncontrib, const float *dL_dimage, float *dL_duv,
                        float *dL_dconic, float *dL_dopacity, float *dL_dfeature) {
    const int gx = (W + TILE - 1) / TILE;
    const size_t nacc = (size_t)P * (size_t)(6 + C);
    double *acc = (double *)calloc(nacc > 0 ? nacc : 1, sizeof(double));
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; ++y) {
        float *accum_rec = (float *)malloc(sizeof(float) * (size_t)(3 * C + 1));
        float *last_feat = accum_rec + C;
        float *dpix = accum_rec + 2 * C;
        for (int x = 0; x < W; ++x) {
            const int tile = (y / TILE) * gx + (x / TILE);
            const int r0 = tile_range[2 * tile], r1 = tile_range[2 * tile + 1];
            const size_t pix = (size_t)y * W + x;
            const float T_final = final_T[pix]; /* :155 */
            float T = T_final;
            const int last_contributor = ncontrib[pix]; /* :159 */
            float bg_dot = 0.0f;
            for (int k = 0; k < C; ++k) {
                accum_rec[k] = 0.0f;
                last_feat[k] = 0.0f;
                dpix[k] = dL_dimage[(size_t)k * H * W + pix]; /* :163-165 */
                bg_dot += bg * dpix[k];                        /* :225-227 */
            }
            float last_alpha = 0.0f;
            /* :173 walks idx_sorted[range.y - progress - 1]; entries with list position
             * >= last_contributor are skipped (:185-187), so start at the last contributor. */
            int e_hi = r0 + last_contributor;
            if (e_hi > r1) e_hi = r1;
            for (int e = e_hi - 1; e >= r0; --e) {
                const int g = idx_sorted[e];
                const float dx = uv[2 * g] - (float)x, dy = uv[2 * g + 1] - (float)y;
                const float cx = conic[3 * g], cy = conic[3 * g + 1], cz = conic[3 * g + 2];
                const float power = pair_power(cx, cy, cz, dx, dy);
                if (power > 0.0f) continue; /* :196 */
                const float G = expf(power);
                const float opac = opacity[g];
                const float alpha = fminf(0.99f, opac * G); /* :201 */
                if (alpha < 1.0f / 255.0f) continue;
                T = T / (1.0f - alpha);          /* :205 */
                const float w = alpha * T;       /* dchannel_dcolor :206 */
                float dL_dalpha = 0.0f;
                double *a = acc + (size_t)g * (size_t)(6 + C);
                for (int k = 0; k < C; ++k) {
                    const float f = feature[(size_t)g * C + k];
                    accum_rec[k] = last_alpha * last_feat[k] + (1.0f - last_alpha) * accum_rec[k]; /* :213-214 */
                    last_feat[k] = f;
                    dL_dalpha += (f - accum_rec[k]) * dpix[k]; /* :217 */
                    const double v = (double)(w * dpix[k]);   /* :218-219 */
#pragma omp atomic
                    a[6 + k] += v;
                }
                dL_dalpha *= T; /* :222 */
                last_alpha = alpha;
                dL_dalpha += (-T_final / (1.0f - alpha)) * bg_dot; /* :229 */
                const float dL_dG = opac * dL_dalpha;              /* :231 */
                const float dGx = -G * dx * cx - G * dy * cy;      /* :232-234 */
                const float dGy = -G * dy * cz - G * dx * cy;
                const double g0 = (double)(dL_dG * dGx), g1 = (double)(dL_dG * dGy);
                const double g2 = (double)(-0.5f * G * dx * dx * dL_dG); /* :238-242 */
                const double g3 = (double)(-G * dx * dy * dL_dG);
                const double g4 = (double)(-0.5f * G * dy * dy * dL_dG);
                const double g5 = (double)(G * dL_dalpha); /* :243 */
#pragma omp atomic
                a[0] += g0;
#pragma omp atomic
                a[1] += g1;
#pragma omp atomic
                a[2] += g2;
#pragma omp atomic
                a[3] += g3;
#pragma omp atomic
                a[4] += g4;
#pragma omp atomic
                a[5] += g5;
            }
        }
        free(accum_rec);
    }
    for (int g = 0; g < P; ++g) {
        const double *a = acc + (size_t)g * (size_t)(6 + C);
        dL_duv[2 * g] = (float)a[0];
        dL_duv[2 * g + 1] = (float)a[1];
        dL_dconic[3 * g] = (float)a[2];
        dL_dconic[3 * g + 1] = (float)a[3];
        dL_dconic[3 * g + 2] = (float)a[4];
        dL_dopacity[g] = (float)a[5];
        for (int k = 0; k < C; ++k) dL_dfeature[(size_t)g * C + k] = (float)a[6 + k];
    }
    free(acc);
}
