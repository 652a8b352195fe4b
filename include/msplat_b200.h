/*
 * msplat_b200.h -- C ABI of libmsplat_b200.so: the B200-native (sm_100a) replacement for the
 * 12-function pybind11 module `msplat._C` of pointrix-project/msplat
 * (/root/reference/msplat/src/ext.cpp:13-26; prototypes /root/reference/msplat/include/*.h).
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless named *_host;
 *  - the caller owns all memory: outputs and workspaces are allocated by the caller (the Python
 *    side uses torch's stream-ordered caching allocator); the library never allocates, frees or
 *    synchronises, keeps no mutable global state and is re-entrant;
 *  - `stream` is a cudaStream_t; kernels run on the caller's current device;
 *  - return value: 0 = success, > 0 = cudaError_t of a failed launch, < 0 = argument error
 *    (-1 bad argument / misalignment, -2 workspace too small, -3 size out of range);
 *    msb_last_error() returns a thread-local description;
 *  - float tensors are float32, row-major contiguous and 16-byte aligned at their base;
 *    `visible` is one byte per Gaussian (torch.bool), NULL = all visible;
 *  - outputs for culled / invisible / degenerate Gaussians are written as zeros (the reference
 *    relies on torch::zeros pre-initialisation), so the caller may pass uninitialised buffers.
 */
#ifndef MSPLAT_B200_H
#define MSPLAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

int msb_version(void);
int msb_sm_count(void);
const char* msb_last_error(void);

/* ---- project_point ------------------------------------------------------------------------
 * replaces projectPointsForward  (include/project_point.h, src/project_point.cu:147-179)
 *          projectPointsBackward (src/project_point.cu:181-228)
 * xyz [P,3], intr [4] = fx fy cx cy, extr first 12 floats of [3,4]/[4,4]; uv [P,2], depth [P].
 * dL_dintr [4] / dL_dextr [12] may be NULL (gradient not wanted); otherwise they must be
 * zero-initialised by the caller and are accumulated into. */
int msb_project_point_fwd(const float* xyz, const float* intr, const float* extr, int P, int W, int H,
                          float nearest, float extent, float* uv, float* depth, void* stream);
int msb_project_point_bwd(const float* xyz, const float* intr, const float* extr, const float* depth,
                          const float* dL_duv, const float* dL_ddepth, int P, float* dL_dxyz, float* dL_dintr,
                          float* dL_dextr, void* stream);

/* ---- compute_cov3d ------------------------------------------------------------------------
 * replaces computeCov3DForward / computeCov3DBackward (src/compute_cov3d.cu:149-200)
 * scale [P,3], quat [P,4] (r,x,y,z, not normalised), cov3d [P,6] upper triangle. */
int msb_compute_cov3d_fwd(const float* scale, const float* quat, const uint8_t* visible, int P, float* cov3d,
                          void* stream);
int msb_compute_cov3d_bwd(const float* scale, const float* quat, const uint8_t* visible, const float* dL_dcov3d,
                          int P, float* dL_dscale, float* dL_dquat, void* stream);

/* ---- ewa_project --------------------------------------------------------------------------
 * replaces EWAProjectForward / EWAProjectBackward (src/ewa_project.cu:254-345)
 * conic [P,3], radius [P] int32, tiles [P] int32 -- radius and tiles are bit-exact with the
 * reference's sm_100 build. */
int msb_ewa_project_fwd(const float* xyz, const float* cov3d, const float* intr, const float* extr,
                        const float* uv, const uint8_t* visible, int P, int W, int H, float* conic,
                        int32_t* radius, int32_t* tiles, void* stream);
int msb_ewa_project_bwd(const float* xyz, const float* cov3d, const float* intr, const float* extr,
                        const int32_t* radius, const float* dL_dconic, int P, float* dL_dxyz, float* dL_dcov3d,
                        float* dL_dintr, float* dL_dextr, void* stream);

/* ---- fused per-Gaussian stage of rasterization() (msplat/__init__.py:70-81) ------------------
 * project + (visible = depth != 0) + cov3d + ewa in one pass; backward = ewa^T, cov3d^T, project^T.
 * dL_ddepth may be NULL. */
int msb_preprocess_fwd(const float* xyz, const float* scale, const float* quat, const float* intr,
                       const float* extr, int P, int W, int H, float nearest, float extent, float* uv,
                       float* depth, float* conic, int32_t* radius, int32_t* tiles, void* stream);
int msb_preprocess_bwd(const float* xyz, const float* scale, const float* quat, const float* intr,
                       const float* extr, const float* depth, const int32_t* radius, const float* dL_duv,
                       const float* dL_ddepth, const float* dL_dconic, int P, float* dL_dxyz, float* dL_dscale,
                       float* dL_dquat, float* dL_dintr, float* dL_dextr, void* stream);

/* ---- compute_sh ---------------------------------------------------------------------------
 * replaces computeSHForward / computeSHBackward (src/compute_sh.cu:1696-1754)
 * shs [P,Cs,D] (D = (deg+1)^2, deg 0..10, D innermost), dirs [P,3], value [P,Cs]. */
int msb_compute_sh_fwd(const float* shs, const float* dirs, const uint8_t* visible, int P, int Cs, int D,
                       float* value, void* stream);
int msb_compute_sh_bwd(const float* shs, const float* dirs, const uint8_t* visible, const float* dL_dvalue,
                       int P, int Cs, int D, float* dL_dshs, float* dL_ddirs, void* stream);

/* ---- sort_gaussian ------------------------------------------------------------------------
 * replaces torch.cumsum + computeGaussianKey + torch.sort + torch.gather +
 * computeTileGaussianRange (msplat/sort_gaussian.py:42-52, src/sort_gaussian.cu:74-142).
 * Two phases because M = sum(tiles) sizes the output:
 *   1. msb_sort_scan: the int64 total M is copied asynchronously to *total_host (PINNED host
 *      memory) -- synchronise `stream` before reading; offsets [P] (may be NULL) receives the
 *      inclusive int32 cumsum of tiles (torch.cumsum equivalent, not needed by phase 2);
 *   2. msb_sort_gaussian: stable LSD radix sort of the 64-bit key (tile << 32 | depth bits) over
 *      the duplication slots -- the 4 depth-digit onesweep passes run on the P Gaussians before
 *      duplication, the ceil(bits(T-1)/8) tile-digit passes on the M duplicates -- then tile
 *      ranges.  idx_sorted [M] int32, tile_range [T,2] int32 (empty tiles (0,0)); bit-exact with
 *      the reference.  msb_sort_num_passes = 4 + tile-digit passes. */
size_t msb_sort_scan_workspace_bytes(int P);
int msb_sort_scan(const int32_t* tiles, int P, int32_t* offsets, long long* total_host, void* ws,
                  size_t ws_bytes, void* stream);
int msb_sort_num_passes(int W, int H);
size_t msb_sort_workspace_bytes(int P, long long M, int W, int H);
int msb_sort_gaussian(const float* uv, const float* depth, const int32_t* radius, const int32_t* tiles, int P,
                      long long M, int W, int H, int32_t* idx_sorted, int32_t* tile_range, void* ws,
                      size_t ws_bytes, int sm_count, void* stream);
/* View batch: ONE sort for `views` views of the same P Gaussians (the batch is a virtual scene of
 * views * P Gaussians on views * T tiles).  uv/depth/radius/tiles are [views, P] view-major (P a multiple
 * of 4 when views > 1); idx_sorted [M] receives view * P + index, tile_range [views * T, 2] indexes
 * idx_sorted; M = sum of tiles over the batch.  Limits: M <= 2^31 - 1 (int32 positions, like the
 * reference's int32 cumsum, msplat/sort_gaussian.py:42), views * P and views * T < 2^31.  Within a view
 * the order is exactly that of msb_sort_gaussian. */
int msb_sort_num_passes_views(int W, int H, int views);
size_t msb_sort_workspace_bytes_views(int P, int views, long long M, int W, int H);
int msb_sort_gaussian_views(const float* uv, const float* depth, const int32_t* radius, const int32_t* tiles,
                            int P, int views, long long M, int W, int H, int32_t* idx_sorted,
                            int32_t* tile_range, void* ws, size_t ws_bytes, int sm_count, void* stream);

/* Level-1 replacements of the two sort-stage functions of msplat._C, for a maintainer who keeps the
 * reference's msplat/sort_gaussian.py unchanged (torch.cumsum / torch.sort / torch.gather stay in Python):
 * computeGaussianKey (src/sort_gaussian.cu:74-113): cumsum [P] = inclusive int32 cumsum of tiles, M =
 *   cumsum[P-1]; keys [M] int64 = (tile << 32) | depth bits and idx [M] int32, zeroed then filled;
 * computeTileGaussianRange (src/sort_gaussian.cu:115-142): tile_range [T,2] zeroed then filled. */
int msb_compute_gaussian_key(const float* uv, const float* depth, const int32_t* radius, const int32_t* cumsum,
                             int P, long long M, int W, int H, long long* keys, int32_t* idx, void* stream);
int msb_compute_tile_gaussian_range(const long long* keys_sorted, long long M, int W, int H, int32_t* tile_range,
                                    void* stream);

/* ---- alpha_blending -----------------------------------------------------------------------
 * replaces alphaBlendingForward / alphaBlendingBackward (src/alpha_blending.cu:248-573)
 * feature [P,C] row-major (the Python-level layout); image [C,H,W]; final_T [H,W];
 * ncontrib [H,W] int32.  `packed` (msb_blend_fwd_workspace_bytes) is written by the forward
 * call and must be handed unchanged to the backward call; `ws` (msb_blend_bwd_workspace_bytes)
 * is scratch.  dL_dfeature is contiguous [P,C]. */
int msb_blend_cpad(int C);
size_t msb_blend_fwd_workspace_bytes(int P, int C);
size_t msb_blend_bwd_workspace_bytes(int P, int C);
int msb_alpha_blending_fwd(const float* uv, const float* conic, const float* opacity, const float* feature,
                           const int32_t* idx_sorted, const int32_t* tile_range, float bg, int P, int C, int W,
                           int H, float* image, float* final_T, int32_t* ncontrib, void* packed,
                           size_t packed_bytes, void* stream);
int msb_alpha_blending_bwd(const float* feature, const int32_t* idx_sorted, const int32_t* tile_range, float bg,
                           int P, int C, int W, int H, const float* final_T, const int32_t* ncontrib,
                           const float* dL_dimage, const void* packed, float* dL_duv, float* dL_dconic,
                           float* dL_dopacity, float* dL_dfeature, void* ws, size_t ws_bytes, void* stream);

/* Rebuilds `packed` from the forward's inputs (for a backward entry that does not receive the forward's
 * workspace, like msplat._C.alpha_blending_backward). */
int msb_blend_pack(const float* uv, const float* conic, const float* opacity, const float* feature, int P, int C,
                   void* packed, size_t packed_bytes, void* stream);

/* ---- packed blend (inputs/gradients stay in the blend kernels' packed layout) --------------
 * Same kernels as msb_alpha_blending_fwd/bwd (src/alpha_blending.cu:248-573) without the
 * pack / unpack passes: rec [P,8] = {u, v, conic.x, conic.y | conic.z, opacity, c0, c1} where c0, c1
 * hold four FP16 culling extents (written by msb_render_preprocess_fwd; csrc/blend_math.cuh),
 * featp [P,Cpad], Cpad = msb_blend_cpad(C); grec [P,8] = {dL_duv.xy, dL_dconic.xyz,
 * dL_dopacity, 0, 0} and gfeat [P,Cpad] are zeroed and then accumulated by the backward call;
 * already_zero != 0 skips the zeroing (the caller cleared both buffers, e.g. on another stream). */
int msb_blend_packed_fwd(const float* rec, const float* featp, const int32_t* idx_sorted, const int32_t* tile_range,
                         float bg, int C, int W, int H, float* image, float* final_T, int32_t* ncontrib,
                         void* stream);
int msb_blend_packed_bwd(const float* rec, const float* featp, const int32_t* idx_sorted, const int32_t* tile_range,
                         float bg, int P, int C, int W, int H, const float* final_T, const int32_t* ncontrib,
                         const float* dL_dimage, float* grec, float* gfeat, int already_zero, void* stream);
/* View batch in one grid (blockIdx.z = view): rec [views*P,8], featp [views*P,Cpad], idx_sorted and
 * tile_range [views*T,2] from msb_sort_gaussian_views; image / dL_dimage [views,C,H,W], final_T /
 * ncontrib [views,H,W]; grec [views*P,8], gfeat [views*P,Cpad] (P = rows per view). */
int msb_blend_packed_fwd_views(const float* rec, const float* featp, const int32_t* idx_sorted,
                               const int32_t* tile_range, float bg, int C, int W, int H, int views, float* image,
                               float* final_T, int32_t* ncontrib, void* stream);
int msb_blend_packed_bwd_views(const float* rec, const float* featp, const int32_t* idx_sorted,
                               const int32_t* tile_range, float bg, int P, int C, int W, int H, int views,
                               const float* final_T, const int32_t* ncontrib, const float* dL_dimage, float* grec,
                               float* gfeat, int already_zero, void* stream);

/* ---- fused SH render preprocess ------------------------------------------------------------
 * One forward / one backward kernel for the whole per-Gaussian part of an SH-coloured render:
 * replaces the sequence projectPointsForward (src/project_point.cu:147-179), computeCov3DForward
 * (src/compute_cov3d.cu:149-170), EWAProjectForward (src/ewa_project.cu:254-297),
 * computeSHForward (src/compute_sh.cu:1696-1721) plus the torch glue of a 3DGS trainer
 * (view_dir = normalize(xyz - camera centre); colour = clamp_min(sh + sh_bias, 0);
 * feature = cat(colour, depth)) -- and the matching *Backward functions.  The camera centre is
 * -R^T t of extr.  shs [P,Cs,D], D = (deg+1)^2, deg <= 10.  C = Cs + (with_depth ? 1 : 0).
 * uv/depth/radius/tiles are bit-identical to msb_project_point_fwd / msb_ewa_project_fwd.
 * forward: total_dev (8 bytes of device memory) and total_host (PINNED host memory) are optional:
 * with total_dev, M = sum(tiles) is accumulated by the kernel into *total_dev (zeroed first); with
 * total_host it is also copied asynchronously to *total_host (synchronise the stream before reading
 * it).  Replaces msb_sort_scan on this path.
 * backward: accumulate != 0 adds into the outputs (view batches), else every element is written;
 * dL_dintr [4] / dL_dextr [12] may be NULL, otherwise they are accumulated into. */
int msb_render_preprocess_fwd(const float* xyz, const float* scale, const float* quat, const float* opacity,
                              const float* shs, const float* intr, const float* extr, int P, int Cs, int D,
                              int with_depth, int W, int H, float nearest, float extent, float sh_bias, int clamp,
                              float* rec, float* featp, float* uv, float* depth, int32_t* radius, int32_t* tiles,
                              long long* total_dev, long long* total_host, void* stream);
int msb_render_preprocess_bwd(const float* xyz, const float* scale, const float* quat, const float* shs,
                              const float* intr, const float* extr, const int32_t* tiles, const float* grec,
                              const float* gfeat, int P, int Cs, int D, int with_depth, float sh_bias, int clamp,
                              int accumulate, float* dL_dxyz, float* dL_dscale, float* dL_dquat, float* dL_dopacity,
                              float* dL_dshs, float* dL_dintr, float* dL_dextr, void* stream);

/* Diagnostic used by bench.py for the blend roofline: blended [views,H,W] int32 = number of list entries
 * that actually blend at each pixel (pass the power / alpha tests of src/alpha_blending.cu:85-94 before
 * the pixel terminates).  Same traversal as the forward kernel, no colours. */
int msb_blend_packed_count(const float* rec, const int32_t* idx_sorted, const int32_t* tile_range, int W, int H,
                           int views, int32_t* blended, void* stream);

/* View batch: ONE launch for `views` cameras over the same Gaussians (parameters and SH rows are read
 * once; the backward sums the gradients over the views in registers / at the L2).  intr [views,4];
 * extr [views,estride], estride = 12 ([3,4]) or 16 ([4,4]).  Per-view buffers are view-major with
 * `vstride` rows per view (vstride >= P, a multiple of 4 when views > 1; rows >= P are not touched):
 * rec [views,vstride,8], featp [views,vstride,Cpad], uv [views,vstride,2], depth / radius / tiles
 * [views,vstride], grec / gfeat like rec / featp; total_dev [views] int64 (optional) = M per view.
 * backward: per-Gaussian pointers may be offset to a slab of P Gaussians (tiles / grec / gfeat to the same
 * row of view 0).  dL_dintr [views,4] / dL_dextr [views,estride] optional, accumulated into; dL_dextr includes the view direction's dependence on the camera centre -R^T t. */
int msb_render_preprocess_fwd_views(const float* xyz, const float* scale, const float* quat, const float* opacity,
                                    const float* shs, const float* intr, const float* extr, int estride, int P,
                                    int views, long long vstride, int Cs, int D, int with_depth, int W, int H,
                                    float nearest, float extent, float sh_bias, int clamp, float* rec, float* featp,
                                    float* uv, float* depth, int32_t* radius, int32_t* tiles, long long* total_dev,
                                    void* stream);
int msb_render_preprocess_bwd_views(const float* xyz, const float* scale, const float* quat, const float* shs,
                                    const float* intr, const float* extr, int estride, const int32_t* tiles,
                                    const float* grec, const float* gfeat, int P, int views, long long vstride,
                                    int Cs, int D, int with_depth,
                                    float sh_bias, int clamp, int accumulate, float* dL_dxyz, float* dL_dscale,
                                    float* dL_dquat, float* dL_dopacity, float* dL_dshs, float* dL_dintr,
                                    float* dL_dextr, void* stream);

/* ---- fused multi-tensor Adam --------------------------------------------------------------------
 * replaces the per-tensor elementwise kernels of torch.optim.Adam in the reference's training loop
 * (/root/reference/tutorials/gs_2d.py:32-36 optimizer over xyz/scale/rotate/opacity/rgb, :66-87 loop): ONE launch
 * per 8 tensors.  torch.optim.Adam semantics (no weight decay, no amsgrad); params / grads / exp_avg /
 * exp_avg_sq are HOST arrays of `ntensors` device pointers, numel a host array of element counts, `step` the
 * 1-based step count of the bias corrections. */
int msb_adam_step(int ntensors, float* const* params, const float* const* grads, float* const* exp_avg,
                  float* const* exp_avg_sq, const long long* numel, float lr, float beta1, float beta2, float eps,
                  int step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MSPLAT_B200_H */
