// msplat_b200/csrc/sort.cu -- tile|depth key duplication + onesweep radix sort + tile ranges.
//
// Replaces the reference's sort stage:
//   torch.cumsum(int32)            /root/reference/msplat/sort_gaussian.py:42
//   computeGaussianKeyCUDAKernel   /root/reference/msplat/src/sort_gaussian.cu:17-43
//   torch.sort (64-bit) + gather   /root/reference/msplat/sort_gaussian.py:49-50
//   computeTileGaussianRange...    /root/reference/msplat/src/sort_gaussian.cu:45-71
// Contract (bit-exact): key = (tile_id << 32) | depth_bits, tile_id = y * ceil(W/16) + x,
// duplication order rows-then-columns starting at cumsum[idx-1]; slots that no Gaussian writes
// hold (key 0, idx 0); the sort is stable w.r.t. slot order; empty tiles have range (0, 0).
// Defined-behaviour choices where the reference is UB: depth bits are masked to 32 bits and a
// Gaussian never writes more than tiles[idx] entries.
//
// Pipeline (all integer work, HBM-bound):
//   phase 1  scan      : 3 small kernels -> inclusive offsets[P], M (copied to pinned host memory)
//   phase 2a duplicate : warp-cooperative key/value emission; the digit histograms of ALL radix
//                        passes are accumulated on the fly (depth digits once per Gaussian), so
//                        the keys are never re-read for a histogram pass            12 B/key
//   phase 2b onesweep  : p = ceil((32 + bits(T-1)) / 8) passes of 8 bits; each pass is ONE kernel:
//                        warp-ballot (match.any) ranking into shared-memory histograms, chained
//                        scan with decoupled look-back on a (value|flag) word per digit, keys
//                        exchanged through shared memory so global writes are coalesced  24 B/key/pass
//   phase 2c ranges    : boundary detection on the sorted keys                          8 B/key
#include "geom.cuh"

namespace msb {

// ------------------------------------------------------------------------------------------------
// phase 1: inclusive scan of tiles[P]
// ------------------------------------------------------------------------------------------------
constexpr int SC_NT = 256;
constexpr int SC_IPT = 8;
constexpr int SC_TILE = SC_NT * SC_IPT;  // 2048

__device__ __forceinline__ int warp_incl_scan(int v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}

// block-wide exclusive scan of one int per thread (NTH threads); returns exclusive prefix and total
template <int NTH>
__device__ __forceinline__ int block_excl_scan(int v, int* s_warp /*[NTH/32 + 1]*/, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int inc = warp_incl_scan(v);
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < NTH / 32 ? s_warp[lane] : 0;
        const int winc = warp_incl_scan(w);
        if (lane < NTH / 32) s_warp[lane] = winc - w;
        if (lane == 31) s_warp[NTH / 32] = winc;
    }
    __syncthreads();
    total = s_warp[NTH / 32];
    const int res = s_warp[warp] + inc - v;
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(SC_NT) scan_block_sums_kernel(int P, const int* __restrict__ tiles,
                                                                long long* __restrict__ bsum) {
    __shared__ long long s_part[SC_NT / 32];
    const long long base = (long long)blockIdx.x * SC_TILE;
    long long acc = 0;
#pragma unroll
    for (int k = 0; k < SC_IPT; ++k) {
        const long long i = base + k * SC_NT + threadIdx.x;
        if (i < P) acc += max(tiles[i], 0);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int w = 0; w < SC_NT / 32; ++w) t += s_part[w];
        bsum[blockIdx.x] = t;
    }
}

// single block: exclusive scan of bsum[nb] in place; total -> *total_dev (saturated to INT_MAX+)
__global__ void __launch_bounds__(1024) scan_spine_kernel(int nb, long long* __restrict__ bsum,
                                                          long long* __restrict__ total_dev) {
    __shared__ long long s_w[33];
    long long carry = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const long long v = i < nb ? bsum[i] : 0;
        long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long n = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += n;
        }
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            long long w = s_w[lane];
            long long winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long n = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += n;
            }
            s_w[lane] = winc - w;
            if (lane == 31) s_w[32] = winc;
        }
        __syncthreads();
        if (i < nb) bsum[i] = carry + s_w[warp] + inc - v;
        carry += s_w[32];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_dev = carry;
}

__global__ void __launch_bounds__(SC_NT) scan_apply_kernel(int P, const int* __restrict__ tiles,
                                                           const long long* __restrict__ bsum,
                                                           int* __restrict__ offsets) {
    __shared__ int s_warp[SC_NT / 32 + 1];
    const long long base = (long long)blockIdx.x * SC_TILE + (long long)threadIdx.x * SC_IPT;
    int v[SC_IPT];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < SC_IPT; ++k) {
        const long long i = base + k;
        v[k] = i < P ? max(tiles[i], 0) : 0;
        sum += v[k];
    }
    int total;
    int run = block_excl_scan<SC_NT>(sum, s_warp, total) + (int)bsum[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SC_IPT; ++k) {
        run += v[k];
        const long long i = base + k;
        if (i < P) offsets[i] = run;  // inclusive, like torch.cumsum
    }
}

// ------------------------------------------------------------------------------------------------
// phase 2a: key/value duplication + histograms of every radix pass
// ------------------------------------------------------------------------------------------------
constexpr int MAX_PASS = 8;
constexpr int DUP_NT = 256;
constexpr int DUP_SMALL = 8;

__global__ void __launch_bounds__(DUP_NT) duplicate_kernel(int P, const float2* __restrict__ uv,
                                                           const float* __restrict__ depth,
                                                           const int* __restrict__ radius,
                                                           const int* __restrict__ tiles,
                                                           const int* __restrict__ offsets, int gx, int gy,
                                                           int npass, unsigned long long* __restrict__ keys,
                                                           int* __restrict__ vals,
                                                           unsigned int* __restrict__ hist /*[npass][256]*/) {
    __shared__ unsigned int s_hist[MAX_PASS * 256];
    for (int i = threadIdx.x; i < npass * 256; i += DUP_NT) s_hist[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long nchunks = ((long long)P + DUP_NT - 1) / DUP_NT;
    for (long long chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const long long i = chunk * DUP_NT + threadIdx.x;
        int n = 0, z = 0, start = 0, x0 = 0, y0 = 0, w = 1;
        unsigned int dbits = 0;
        if (i < P) {
            const int slots = max(tiles[i], 0);
            start = offsets[i] - slots;  // == cumsum[i-1] (sort_gaussian.cu:32)
            const int rad = radius[i];
            if (rad > 0) {  // sort_gaussian.cu:26
                const float2 c = uv[i];
                const Rect q = get_rect(c.x, c.y, rad, gx, gy);
                x0 = q.x0;
                y0 = q.y0;
                w = max(q.x1 - q.x0, 1);
                n = min(max((q.x1 - q.x0) * (q.y1 - q.y0), 0), slots);
                dbits = __float_as_uint(depth[i]);
            }
            z = slots - n;
            if (n > 0) {
                for (int p = 0; p < 4 && p < npass; ++p)
                    atomicAdd(&s_hist[p * 256 + ((dbits >> (8 * p)) & 255u)], (unsigned)n);
            }
            if (z > 0) {
                for (int p = 0; p < npass; ++p) atomicAdd(&s_hist[p * 256], (unsigned)z);
                for (int e = 0; e < z; ++e) {  // never-written slots stay (0, 0): sort_gaussian.cu:98-99
                    keys[start + n + e] = 0ull;
                    vals[start + n + e] = 0;
                }
            }
        }
        // small footprints: each lane emits its own run
        if (n > 0 && n <= DUP_SMALL) {
            int tx = x0, ty = y0;
            for (int e = 0; e < n; ++e) {
                const unsigned int tile = (unsigned)(ty * gx + tx);
                keys[start + e] = ((unsigned long long)tile << 32) | dbits;
                vals[start + e] = (int)i;
                for (int p = 4; p < npass; ++p) atomicAdd(&s_hist[p * 256 + ((tile >> (8 * (p - 4))) & 255u)], 1u);
                if (++tx == x0 + w) {
                    tx = x0;
                    ++ty;
                }
            }
        }
        // large footprints: the whole warp emits one Gaussian at a time (reference: one thread
        // serially writes up to T entries, sort_gaussian.cu:35-42)
        unsigned big = __ballot_sync(0xffffffffu, n > DUP_SMALL);
        while (big) {
            const int src = __ffs(big) - 1;
            big &= big - 1;
            const int bn = __shfl_sync(0xffffffffu, n, src);
            const int bstart = __shfl_sync(0xffffffffu, start, src);
            const int bx0 = __shfl_sync(0xffffffffu, x0, src);
            const int by0 = __shfl_sync(0xffffffffu, y0, src);
            const int bw = __shfl_sync(0xffffffffu, w, src);
            const unsigned bd = __shfl_sync(0xffffffffu, dbits, src);
            const int bi = (int)(chunk * DUP_NT + (threadIdx.x & ~31) + src);
            for (int e = lane; e < bn; e += 32) {
                const int ry = e / bw, rx = e - ry * bw;
                const unsigned int tile = (unsigned)((by0 + ry) * gx + bx0 + rx);
                keys[bstart + e] = ((unsigned long long)tile << 32) | bd;
                vals[bstart + e] = bi;
                for (int p = 4; p < npass; ++p) atomicAdd(&s_hist[p * 256 + ((tile >> (8 * (p - 4))) & 255u)], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < npass * 256; i += DUP_NT) {
        const unsigned int c = s_hist[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

// ------------------------------------------------------------------------------------------------
// phase 2b: one onesweep pass (8-bit digit) over (key64, val32) pairs
// ------------------------------------------------------------------------------------------------
constexpr int RS_NT = 256;
constexpr int RS_WARPS = RS_NT / 32;
constexpr int RS_IPT = 16;
constexpr int RS_TILE = RS_NT * RS_IPT;  // 4096 keys per CTA
constexpr unsigned int RS_FLAG_AGG = 1u << 30;
constexpr unsigned int RS_FLAG_PRE = 2u << 30;
constexpr unsigned int RS_VALUE_MASK = (1u << 30) - 1u;

struct RsSmem {
    unsigned long long keys[RS_TILE];    // 32 KB exchange buffer
    int vals[RS_TILE];                   // 16 KB
    unsigned int whist[RS_WARPS][256];   //  8 KB per-warp digit counts -> per-warp offsets
    unsigned int bin_start[256];         // tile-local exclusive digit offsets
    long long gadj[256];                 // global base of the digit minus bin_start
    int scan_tmp[RS_NT / 32 + 1];
    int tile_id;
};

__device__ __forceinline__ unsigned int ld_relaxed(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(unsigned int* p, unsigned int v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(RS_NT) onesweep_kernel(int M, int pass, const unsigned long long* __restrict__ kin,
                                                         const int* __restrict__ vin,
                                                         unsigned long long* __restrict__ kout,
                                                         int* __restrict__ vout,
                                                         const unsigned int* __restrict__ hist /*[256] this pass*/,
                                                         unsigned int* __restrict__ status /*[ntiles][256]*/,
                                                         unsigned int* __restrict__ ticket) {
    extern __shared__ __align__(16) unsigned char rs_raw[];
    RsSmem& sm = *reinterpret_cast<RsSmem*>(rs_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int shift = 8 * pass;

    if (tid == 0) sm.tile_id = (int)atomicAdd(ticket, 1u);  // tiles are ordered by start time
#pragma unroll
    for (int k = 0; k < 256 / 32; ++k) sm.whist[warp][lane + 32 * k] = 0;
    __syncthreads();
    const int tile = sm.tile_id;
    const long long tile_base = (long long)tile * RS_TILE;
    const int valid = (int)min((long long)RS_TILE, (long long)M - tile_base);

    // ---- load (warp-striped: element order = warp, item, lane) --------------------------------
    unsigned long long key[RS_IPT];
    int val[RS_IPT];
    unsigned short rank[RS_IPT];
    const int wbase = warp * (32 * RS_IPT);
#pragma unroll
    for (int i = 0; i < RS_IPT; ++i) {
        const int e = wbase + i * 32 + lane;
        if (e < valid) {
            key[i] = kin[tile_base + e];
            val[i] = vin[tile_base + e];
        } else {
            key[i] = ~0ull;  // sorts after everything; digit 255 in every pass
            val[i] = 0;
        }
    }

    // ---- rank inside the warp with match.any (warp-ballot ranking) ------------------------------
    const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < RS_IPT; ++i) {
        const unsigned d = (unsigned)(key[i] >> shift) & 255u;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        unsigned base = 0;
        if (lane == leader) {
            base = sm.whist[warp][d];
            sm.whist[warp][d] = base + __popc(peers);
        }
        base = __shfl_sync(0xffffffffu, base, leader);
        rank[i] = (unsigned short)(base + __popc(peers & lt_mask));
        __syncwarp();
    }
    __syncthreads();

    // ---- per digit (thread tid == digit): warp prefix, tile count, look-back ---------------------
    unsigned int count = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
        const unsigned int c = sm.whist[w][tid];
        sm.whist[w][tid] = count;
        count += c;
    }
    if (tid == 255) count -= (unsigned)(RS_TILE - valid);  // padding keys are not real
    unsigned int* my_status = status + (size_t)tile * 256 + tid;
    st_relaxed(my_status, (tile == 0 ? RS_FLAG_PRE : RS_FLAG_AGG) | count);

    // exclusive scan over digits of the global histogram and of the tile histogram
    int tot;
    const unsigned int gbase = (unsigned)block_excl_scan<RS_NT>((int)hist[tid], sm.scan_tmp, tot);
    const unsigned int lbase = (unsigned)block_excl_scan<RS_NT>((int)count, sm.scan_tmp, tot);

    unsigned int excl = 0;
    if (tile > 0) {
        int j = tile - 1;
        while (true) {
            const unsigned int* p = status + (size_t)j * 256 + tid;
            unsigned int v = ld_relaxed(p);
            while ((v & ~RS_VALUE_MASK) == 0u) {
                __nanosleep(20);
                v = ld_relaxed(p);
            }
            excl += v & RS_VALUE_MASK;
            if (v & RS_FLAG_PRE) break;
            --j;
        }
        st_relaxed(my_status, RS_FLAG_PRE | (excl + count));
    }
    sm.bin_start[tid] = lbase;
    sm.gadj[tid] = (long long)gbase + (long long)excl - (long long)lbase;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) sm.whist[w][tid] += lbase;  // warp offset inside the tile
    __syncthreads();

    // ---- exchange through shared memory (tile-local sorted order) ------------------------------
#pragma unroll
    for (int i = 0; i < RS_IPT; ++i) {
        const unsigned d = (unsigned)(key[i] >> shift) & 255u;
        const unsigned pos = sm.whist[warp][d] + rank[i];
        sm.keys[pos] = key[i];
        sm.vals[pos] = val[i];
    }
    __syncthreads();

    // ---- coalesced scatter ----------------------------------------------------------------------
#pragma unroll 4
    for (int j = tid; j < valid; j += RS_NT) {
        const unsigned long long k = sm.keys[j];
        const unsigned d = (unsigned)(k >> shift) & 255u;
        const long long g = sm.gadj[d] + j;
        kout[g] = k;
        vout[g] = sm.vals[j];
    }
}

// ------------------------------------------------------------------------------------------------
// phase 2c: tile ranges from the sorted keys (sort_gaussian.cu:45-71); tile_range pre-zeroed
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tile_range_kernel(int M, const unsigned long long* __restrict__ keys,
                                                         int2* __restrict__ tile_range, int T) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const unsigned int cur = (unsigned int)(keys[i] >> 32);
    if (cur >= (unsigned)T) return;  // cannot happen for keys produced by duplicate_kernel
    if (i == 0) tile_range[cur].x = 0;
    if (i == M - 1) tile_range[cur].y = M;
    if (i == 0) return;
    const unsigned int prev = (unsigned int)(keys[i - 1] >> 32);
    if (prev != cur) {
        if (prev < (unsigned)T) tile_range[prev].y = (int)i;
        tile_range[cur].x = (int)i;
    }
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static int num_passes(int T) {
    int bits = 0;
    while (bits < 31 && (1ll << bits) < (long long)T) ++bits;  // bits needed for tile ids 0..T-1
    return (32 + bits + 7) / 8;
}

struct SortLayout {
    size_t keys_a, keys_b, vals_tmp, hist, ticket, status, total;
    int npass, ntiles_rs;
};

static SortLayout sort_layout(long long M, int T) {
    SortLayout L;
    L.npass = num_passes(T);
    L.ntiles_rs = (int)((M + RS_TILE - 1) / RS_TILE);
    size_t off = 0;
    L.keys_a = off; off = align_up(off + (size_t)M * 8, 256);
    L.keys_b = off; off = align_up(off + (size_t)M * 8, 256);
    L.vals_tmp = off; off = align_up(off + (size_t)M * 4, 256);
    // zero-initialised region: hist | ticket | status
    L.hist = off; off = align_up(off + (size_t)MAX_PASS * 256 * 4, 256);
    L.ticket = off; off = align_up(off + (size_t)MAX_PASS * 4, 256);
    L.status = off; off = align_up(off + (size_t)L.npass * (size_t)L.ntiles_rs * 256 * 4, 256);
    L.total = off;
    return L;
}

}  // namespace msb

using namespace msb;

extern "C" {

// Workspace (bytes) for msb_sort_scan: block sums + the device-side total.
size_t msb_sort_scan_workspace_bytes(int P) {
    const size_t nb = ((size_t)(P > 0 ? P : 0) + SC_TILE - 1) / SC_TILE;
    return (nb + 2) * sizeof(long long);
}

// Phase 1.  offsets[P] = inclusive int32 cumsum of max(tiles, 0); the 64-bit total is copied
// asynchronously into *total_host (pinned host memory): the caller synchronises the stream
// before reading it.
int msb_sort_scan(const int32_t* tiles, int P, int32_t* offsets, long long* total_host, void* ws, size_t ws_bytes,
                  void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (P < 0 || !total_host) return set_error(MSB_ERR_ARG, "sort_scan: bad argument");
    if (P == 0) {
        *total_host = 0;
        return MSB_OK;
    }
    if (!tiles || !offsets || !ws) return set_error(MSB_ERR_ARG, "sort_scan: null pointer");
    if (ws_bytes < msb_sort_scan_workspace_bytes(P)) return set_error(MSB_ERR_WORKSPACE, "sort_scan: workspace too small");
    const int nb = (P + SC_TILE - 1) / SC_TILE;
    long long* bsum = reinterpret_cast<long long*>(ws);
    long long* total_dev = bsum + nb;
    scan_block_sums_kernel<<<nb, SC_NT, 0, st>>>(P, tiles, bsum);
    scan_spine_kernel<<<1, 1024, 0, st>>>(nb, bsum, total_dev);
    scan_apply_kernel<<<nb, SC_NT, 0, st>>>(P, tiles, bsum, offsets);
    int rc = check_launch("sort_scan");
    if (rc) return rc;
    cudaError_t e = cudaMemcpyAsync(total_host, total_dev, sizeof(long long), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return set_error((int)e, "sort_scan: cudaMemcpyAsync failed");
    return MSB_OK;
}

int msb_sort_num_passes(int W, int H) {
    const int gx = (W + MSB_TILE - 1) / MSB_TILE, gy = (H + MSB_TILE - 1) / MSB_TILE;
    return num_passes(gx * gy);
}

size_t msb_sort_workspace_bytes(long long M, int W, int H) {
    const int gx = (W + MSB_TILE - 1) / MSB_TILE, gy = (H + MSB_TILE - 1) / MSB_TILE;
    if (M <= 0) return 256;
    return sort_layout(M, gx * gy).total;
}

// Phase 2.  idx_sorted[M] (int32) and tile_range[T, 2] (int32) are outputs owned by the caller.
int msb_sort_gaussian(const float* uv, const float* depth, const int32_t* radius, const int32_t* tiles,
                      const int32_t* offsets, int P, long long M, int W, int H, int32_t* idx_sorted,
                      int32_t* tile_range, void* ws, size_t ws_bytes, int sm_count, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int gx = (W + MSB_TILE - 1) / MSB_TILE, gy = (H + MSB_TILE - 1) / MSB_TILE;
    const int T = gx * gy;
    if (P < 0 || M < 0 || W <= 0 || H <= 0 || !tile_range) return set_error(MSB_ERR_ARG, "sort_gaussian: bad argument");
    if (M >= (1ll << 30)) return set_error(MSB_ERR_RANGE, "sort_gaussian: more than 2^30 tile intersections");
    cudaError_t e = cudaMemsetAsync(tile_range, 0, (size_t)T * 2 * sizeof(int32_t), st);
    if (e != cudaSuccess) return set_error((int)e, "sort_gaussian: memset tile_range failed");
    if (M == 0 || P == 0) return MSB_OK;
    if (!uv || !depth || !radius || !tiles || !offsets || !idx_sorted || !ws)
        return set_error(MSB_ERR_ARG, "sort_gaussian: null pointer");
    const SortLayout L = sort_layout(M, T);
    if (ws_bytes < L.total) return set_error(MSB_ERR_WORKSPACE, "sort_gaussian: workspace too small");
    unsigned char* base = reinterpret_cast<unsigned char*>(ws);
    unsigned long long* kbuf[2] = {reinterpret_cast<unsigned long long*>(base + L.keys_a),
                                   reinterpret_cast<unsigned long long*>(base + L.keys_b)};
    int* vtmp = reinterpret_cast<int*>(base + L.vals_tmp);
    // values ping-pong so that the last pass lands in idx_sorted
    int* vbuf[2];
    vbuf[L.npass % 2] = idx_sorted;
    vbuf[(L.npass + 1) % 2] = vtmp;
    unsigned int* hist = reinterpret_cast<unsigned int*>(base + L.hist);
    unsigned int* ticket = reinterpret_cast<unsigned int*>(base + L.ticket);
    unsigned int* status = reinterpret_cast<unsigned int*>(base + L.status);
    e = cudaMemsetAsync(base + L.hist, 0, L.total - L.hist, st);
    if (e != cudaSuccess) return set_error((int)e, "sort_gaussian: memset workspace failed");

    const long long nchunks = ((long long)P + DUP_NT - 1) / DUP_NT;
    const int sms = sm_count > 0 ? sm_count : 148;
    const unsigned dup_grid = (unsigned)min(nchunks, (long long)sms * 8);
    duplicate_kernel<<<dup_grid, DUP_NT, 0, st>>>(P, reinterpret_cast<const float2*>(uv), depth, radius, tiles, offsets,
                                                  gx, gy, L.npass, kbuf[0], vbuf[0], hist);
    int rc = check_launch("sort_gaussian/duplicate");
    if (rc) return rc;

    static_assert(sizeof(RsSmem) <= 100 * 1024, "onesweep shared memory");
    e = cudaFuncSetAttribute(onesweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmem));
    if (e != cudaSuccess) return set_error((int)e, "sort_gaussian: cudaFuncSetAttribute failed");
    for (int p = 0; p < L.npass; ++p) {
        onesweep_kernel<<<L.ntiles_rs, RS_NT, sizeof(RsSmem), st>>>(
            (int)M, p, kbuf[p % 2], vbuf[p % 2], kbuf[(p + 1) % 2], vbuf[(p + 1) % 2], hist + p * 256,
            status + (size_t)p * L.ntiles_rs * 256, ticket + p);
        rc = check_launch("sort_gaussian/onesweep");
        if (rc) return rc;
    }
    tile_range_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>((int)M, kbuf[L.npass % 2],
                                                                  reinterpret_cast<int2*>(tile_range), T);
    return check_launch("sort_gaussian/tile_range");
}

}  // extern "C"
