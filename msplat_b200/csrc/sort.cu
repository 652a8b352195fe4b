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
// The order the reference produces is the stable LSD radix order of the 64-bit key over the
// duplication slots.  All duplicates of a Gaussian share the low key word (its depth bits), so
// the four depth-digit passes commute with the duplication: they are run on the P Gaussians
// (8 B per element) instead of on the M = sum(tiles) duplicates (12 B per element), and only
// the tile-digit passes (ceil(bits(T-1)/8): 2 at 1080p and 4K) touch the duplicates, with a
// 32-bit key (the tile id).  Stability of every pass keeps ties (same tile, same depth bits) in
// Gaussian-index order, exactly like the stable sort of the slots.  Never-written slots are
// all (key 0, idx 0) and therefore form a prefix of tile 0: they are emitted first.
//
// Pipeline (all integer work):
//   phase 1  count     : M = sum(max(tiles,0)) -> pinned host memory (sizes the output)
//   phase 2a keygen    : the Pc Gaussians that emit at least one entry are compacted (index order
//                        kept, no chained scan) into (depth key, id) pairs; per Gaussian a 16-byte
//                        record {x0, y0, w, n} of its tile rectangle; digit histograms of the 4 depth
//                        passes; phantom-slot count Z.  Pc stays on the device.
//   phase 2b onesweep  : 4 passes of 8 bits over (depth key, Gaussian id), Pc elements;
//                        each pass is ONE kernel: ballot-built peer masks -> ranks in shared-memory
//                        histograms, chained scan with decoupled look-back on a (value|flag) word per
//                        digit, keys exchanged through shared memory so global writes are coalesced;
//                        the LAST pass writes, instead of (key, id), the depth-ordered records
//                        {x0 | w << 16, y0, n, id}: the pipeline's one random gather (rect[id])
//   phase 2c offsets   : single-pass chained exclusive scan of n in depth order (+Z), sequential reads
//   phase 2d duplicate : warp-cooperative emission of (tile id, Gaussian id) in depth order from the
//                        records (sequential reads); the histograms of the tile passes are accumulated on the fly
//   phase 2e onesweep  : ceil(bits(T-1)/8) passes over (tile id, Gaussian id), M elements; the top
//                        digit compares only the bits the tile ids use
//   phase 2f ranges    : boundary detection on the sorted tile ids
// Algorithmic bytes: 20 B/Gaussian keygen reads (+20 again from L2) + per emitter 8 B pair + 16 B
//                    record, 4 * 16 B/emitter depth passes (the last: + 16 B gather, 16 B record instead of
//                    the 8 B pair), 16 + 4 B/emitter offsets, 20 B/emitter duplicate reads,
//                    8 B/key duplicate write + pt * 16 B/key + 4 B/key ranges.
// View batches: the arrays hold views * P entries; see keygen_kernel.
#include <stdlib.h>

#include "geom.cuh"

namespace msb {

// ------------------------------------------------------------------------------------------------
// scans
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

// value i of the scanned sequence
__device__ __forceinline__ int scan_value(const int* __restrict__ vals, long long i) { return max(vals[i], 0); }

__global__ void __launch_bounds__(SC_NT) scan_block_sums_kernel(int P, const int* __restrict__ vals,
                                                                long long* __restrict__ bsum) {
    __shared__ long long s_part[SC_NT / 32];
    const long long base = (long long)blockIdx.x * SC_TILE;
    long long acc = 0;
#pragma unroll
    for (int k = 0; k < SC_IPT; ++k) {
        const long long i = base + k * SC_NT + threadIdx.x;
        if (i < P) acc += scan_value(vals, i);
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

// single block: exclusive scan of bsum[nb] in place; total -> *total_dev
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

// out[i] = (EXCL ? exclusive : inclusive) prefix of the sequence
template <bool EXCL>
__global__ void __launch_bounds__(SC_NT) scan_apply_kernel(int P, const int* __restrict__ vals,
                                                           const long long* __restrict__ bsum,
                                                           int* __restrict__ out) {
    __shared__ int s_warp[SC_NT / 32 + 1];
    const long long base = (long long)blockIdx.x * SC_TILE + (long long)threadIdx.x * SC_IPT;
    int v[SC_IPT];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < SC_IPT; ++k) {
        const long long i = base + k;
        v[k] = i < P ? scan_value(vals, i) : 0;
        sum += v[k];
    }
    int total;
    int run = block_excl_scan<SC_NT>(sum, s_warp, total) + (int)bsum[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SC_IPT; ++k) {
        const long long i = base + k;
        if (EXCL) {
            if (i < P) out[i] = run;
            run += v[k];
        } else {
            run += v[k];
            if (i < P) out[i] = run;  // inclusive, like torch.cumsum
        }
    }
}

// ------------------------------------------------------------------------------------------------
// phase 2a: depth key + emitted-entry count of every Gaussian that emits, stably COMPACTED
// ------------------------------------------------------------------------------------------------
// Gaussians that emit no entry (culled, zero footprint; ~40 % of BASELINE config #3) are dropped
// here instead of being carried through the four depth passes as 0xFFFFFFFF keys: the passes,
// the offset scan and the duplication then run over Pc = #emitting Gaussians, a count that only
// exists on the device (*pcount).  The output keeps Gaussian-index order (the stability of the
// sort relies on it) without a chained scan: the grid is at most one wave of CTAs; virtual CTA v
// (ticket order) owns the contiguous index range [v R, (v+1) R), each of its warps a contiguous
// sub-range.  Pass A counts the emitters per warp; the CTA publishes (flag | count) and sums the
// counts of the virtual CTAs before it -- all of which have started, so the wait is one round
// trip, not a chain; pass B re-reads the range (L2 hits) and writes each warp's emitters at its
// prefix with ballot ranks.
constexpr int KG_NT = 256;
constexpr int KG_WARPS = KG_NT / 32;
constexpr int MAX_TPASS = 4;  // tile id < 2^31
// Status word of the chained scans / decoupled look-backs: 0 = nothing published yet; bit 31 set =
// inclusive PREFIX (31-bit value); otherwise an AGGREGATE stored as value + 1.  Every value is
// < 2^31, the bound the reference's int32 cumsum puts on M (msplat/sort_gaussian.py:42).
constexpr unsigned int RS_PRE_BIT = 0x80000000u;
__device__ __forceinline__ unsigned int rs_agg(unsigned int v) { return v + 1u; }
__device__ __forceinline__ unsigned int rs_pre(unsigned int v) { return v | RS_PRE_BIT; }
__device__ __forceinline__ bool rs_is_pre(unsigned int w) { return (w & RS_PRE_BIT) != 0u; }
__device__ __forceinline__ unsigned int rs_value(unsigned int w) { return rs_is_pre(w) ? (w & ~RS_PRE_BIT) : (w - 1u); }

__device__ __forceinline__ unsigned int ld_relaxed(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(unsigned int* p, unsigned int v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// emitted entries of a Gaussian (0 if it takes no part) and its tile rectangle
__device__ __forceinline__ int keygen_count(int slots, int rad, float2 p, int gx, int gy, int& x0, int& y0, int& w) {
    int n = 0;
    x0 = y0 = 0;
    w = 1;
    if (rad > 0 && slots > 0) {  // sort_gaussian.cu:26
        const Rect q = get_rect(p.x, p.y, rad, gx, gy);
        n = min(max((q.x1 - q.x0) * (q.y1 - q.y0), 0), slots);
        x0 = q.x0;
        y0 = q.y0;
        w = max(q.x1 - q.x0, 1);
    }
    return n;
}

constexpr int KG_ROWS = 4;  // rows of 32 Gaussians whose loads are issued together (memory-level parallelism)

// View batches (views > 1): the arrays hold views * Pv entries, view-major; entry i belongs to view
// i / Pv and its tile rectangle is shifted down by view * gy rows of a virtual gx x (views gy) grid, so
// that the whole batch is ONE sort whose tile id is view * T + tile and whose Gaussian id is
// view * Pv + index (what the batched blend kernels index their packed records with).
__global__ void __launch_bounds__(KG_NT) keygen_kernel(int P, int Pv /*Gaussians per view*/,
                                                       int seg /*Gaussians per warp, multiple of 32*/,
                                                       const float2* __restrict__ uv,
                                                       const float* __restrict__ depth,
                                                       const int* __restrict__ radius,
                                                       const int* __restrict__ tiles, int gx, int gy,
                                                       unsigned int* __restrict__ dkeys, int* __restrict__ dvals,
                                                       int4* __restrict__ rect /*[Pc] {x0 | w << 16, y0, n, id} of the emitters, compact*/,
                                                       unsigned int* __restrict__ hist /*[4][256]*/,
                                                       unsigned int* __restrict__ zcount,
                                                       unsigned int* __restrict__ status /*[gridDim.x]*/,
                                                       unsigned int* __restrict__ ticket,
                                                       unsigned int* __restrict__ pcount) {
    __shared__ unsigned int s_hist[4 * 256];
    __shared__ unsigned int s_wcnt[KG_WARPS];
    __shared__ unsigned int s_part[KG_WARPS];
    __shared__ int s_vcta;
    for (int i = threadIdx.x; i < 4 * 256; i += KG_NT) s_hist[i] = 0;
    if (threadIdx.x == 0) s_vcta = (int)atomicAdd(ticket, 1u);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int vcta = s_vcta;
    const long long w0 = ((long long)vcta * KG_WARPS + warp) * seg;
    const long long w1 = min(w0 + seg, (long long)P);

    // ---- pass A: emitters of this warp's range ---------------------------------------------------
    unsigned int wcount = 0, zsum = 0;
    for (long long i0 = w0; i0 < w1; i0 += 32 * KG_ROWS) {
        int sl[KG_ROWS], rd[KG_ROWS];
        float2 pp[KG_ROWS];
#pragma unroll
        for (int r = 0; r < KG_ROWS; ++r) {
            const long long i = i0 + 32 * r + lane;
            const bool in = i < w1;
            sl[r] = in ? max(tiles[i], 0) : 0;
            rd[r] = in ? radius[i] : 0;
            pp[r] = in ? uv[i] : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int r = 0; r < KG_ROWS; ++r) {
            int x0, y0, w;
            const int n = keygen_count(sl[r], rd[r], pp[r], gx, gy, x0, y0, w);
            zsum += (unsigned)(sl[r] - n);  // never-written slots stay (0, 0): sort_gaussian.cu:98-99
            wcount += __popc(__ballot_sync(0xffffffffu, n > 0));
        }
    }
    if (lane == 0) s_wcnt[warp] = wcount;
    __syncthreads();
    unsigned int wpre = 0, total = 0;
#pragma unroll
    for (int w = 0; w < KG_WARPS; ++w) {
        const unsigned int c = s_wcnt[w];
        wpre += w < warp ? c : 0u;
        total += c;
    }
    if (threadIdx.x == 0) st_relaxed(status + vcta, rs_pre(total));

    // ---- prefix over the virtual CTAs before this one (all of them have started) -----------------
    unsigned int part = 0;
    for (int j = threadIdx.x; j < vcta; j += KG_NT) {
        unsigned int v = ld_relaxed(status + j);
        while (v == 0u) {
            __nanosleep(20);
            v = ld_relaxed(status + j);
        }
        part += rs_value(v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) s_part[warp] = part;
    __syncthreads();
    unsigned int pos = wpre;
#pragma unroll
    for (int w = 0; w < KG_WARPS; ++w) pos += s_part[w];
    if (threadIdx.x == 0 && vcta == (int)gridDim.x - 1) *pcount = pos + total;  // wpre == 0 for thread 0

    // ---- pass B: compacted (depth key, id) pairs, rectangles, depth-digit histograms (L2 re-read) -
    const unsigned lt_mask = (1u << lane) - 1u;
    for (long long i0 = w0; i0 < w1; i0 += 32 * KG_ROWS) {
        int sl[KG_ROWS], rd[KG_ROWS];
        float2 pp[KG_ROWS];
        unsigned int dk[KG_ROWS];
#pragma unroll
        for (int r = 0; r < KG_ROWS; ++r) {
            const long long i = i0 + 32 * r + lane;
            const bool in = i < w1;
            sl[r] = in ? max(tiles[i], 0) : 0;
            rd[r] = in ? radius[i] : 0;
            pp[r] = in ? uv[i] : make_float2(0.f, 0.f);
            dk[r] = in ? __float_as_uint(depth[i]) : 0u;
        }
#pragma unroll
        for (int r = 0; r < KG_ROWS; ++r) {
            const long long i = i0 + 32 * r + lane;
            int x0, y0, w;
            const int n = keygen_count(sl[r], rd[r], pp[r], gx, gy, x0, y0, w);
            const unsigned em = __ballot_sync(0xffffffffu, n > 0);
            if (n > 0) {
                const unsigned int o = pos + __popc(em & lt_mask);
                dkeys[o] = dk[r];
                dvals[o] = (int)o;  // the depth passes carry the emitter's compact slot; its id sits in the record
                rect[o] = make_int4(x0 | (w << 16), y0 + (int)(i / Pv) * gy, n, (int)i);
#pragma unroll
                for (int p = 0; p < 4; ++p) atomicAdd(&s_hist[p * 256 + ((dk[r] >> (8 * p)) & 255u)], 1u);
            }
            pos += __popc(em);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) zsum += __shfl_xor_sync(0xffffffffu, zsum, o);
    if (lane == 0 && zsum) atomicAdd(zcount, zsum);
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * 256; i += KG_NT) {
        const unsigned int c = s_hist[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

// ------------------------------------------------------------------------------------------------
// phase 2c: start offsets in depth order -- single-pass chained scan of n[order[i]] (+ Z)
// ------------------------------------------------------------------------------------------------
// One kernel instead of block sums + spine + apply.  The emitted counts come from the depth-ordered records the
// last depth pass wrote (no gather here).  Tiles of 2048 are taken by ticket; warp 0 resolves the tile's
// exclusive prefix by decoupled look-back over one (flag | inclusive prefix) word per tile, 32
// predecessors per round trip.
__global__ void __launch_bounds__(SC_NT) scan_offsets_kernel(const unsigned int* __restrict__ n_dev,
                                                             const int4* __restrict__ rsorted /*depth order {x0|w<<16, y0, n, id}*/,
                                                             const unsigned int* __restrict__ zcount,
                                                             int* __restrict__ start,
                                                             unsigned int* __restrict__ status,
                                                             unsigned int* __restrict__ ticket) {
    __shared__ int s_warp[SC_NT / 32 + 1];
    __shared__ int s_tile;
    __shared__ unsigned int s_prefix;
    if (threadIdx.x == 0) s_tile = (int)atomicAdd(ticket, 1u);
    __syncthreads();
    const int tile = s_tile;
    const int N = (int)*n_dev;
    const long long base = (long long)tile * SC_TILE + (long long)threadIdx.x * SC_IPT;
    if ((long long)tile * SC_TILE >= N) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int v[SC_IPT];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < SC_IPT; ++k) {
        v[k] = base + k < N ? rsorted[base + k].z : 0;
        sum += v[k];
    }
    int total;
    const int excl = block_excl_scan<SC_NT>(sum, s_warp, total);
    if (gridDim.x <= 4096) {
        // few tiles (all resident at once on this GPU up to ~1.8 M emitters): every tile publishes its total
        // and sums the totals of ALL tiles before it -- one round trip instead of a look-back chain
        __shared__ unsigned int s_part[SC_NT / 32];
        if (threadIdx.x == 0) st_relaxed(status + tile, rs_pre((unsigned)total));
        unsigned int part = 0;
        for (int j = threadIdx.x; j < tile; j += SC_NT) {
            unsigned int w = ld_relaxed(status + j);
            while (w == 0u) {
                __nanosleep(20);
                w = ld_relaxed(status + j);
            }
            part += rs_value(w);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) s_part[warp] = part;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned int pre = 0;
#pragma unroll
            for (int w = 0; w < SC_NT / 32; ++w) pre += s_part[w];
            s_prefix = pre;
        }
    } else if (warp == 0) {
        unsigned int pre = 0;
        if (tile > 0) {
            if (lane == 0) st_relaxed(status + tile, rs_agg((unsigned)total));
            int j = tile - 1;
            while (true) {
                const unsigned int w = (j - lane >= 0) ? ld_relaxed(status + (j - lane)) : RS_PRE_BIT;
                const unsigned ready = __ballot_sync(0xffffffffu, w != 0u);
                const unsigned isp = __ballot_sync(0xffffffffu, rs_is_pre(w));
                const int run = (~ready) ? __ffs(~ready) - 1 : 32;  // predecessors ready without a gap
                const int fp = isp ? __ffs(isp) - 1 : 32;           // nearest one holding a full prefix
                const int take = min(run, fp + 1);
                unsigned int add = lane < take ? rs_value(w) : 0u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) add += __shfl_xor_sync(0xffffffffu, add, o);
                pre += add;
                if (fp < run) break;
                j -= take;
                if (take == 0) __nanosleep(20);
            }
        }
        if (lane == 0) {
            st_relaxed(status + tile, rs_pre(pre + (unsigned)total));
            s_prefix = pre;
        }
    }
    __syncthreads();
    int run = (int)(s_prefix + *zcount) + excl;
#pragma unroll
    for (int k = 0; k < SC_IPT; ++k) {
        if (base + k < N) start[base + k] = run;
        run += v[k];
    }
}

// ------------------------------------------------------------------------------------------------
// phase 2d: duplication in depth order + histograms of the tile-digit passes
// ------------------------------------------------------------------------------------------------
constexpr int DUP_NT = 256;
constexpr int DUP_COOP = 64;  // footprints up to this many tiles are expanded cooperatively

// tile-digit histograms of one emitted entry per lane (valid lanes only)
__device__ __forceinline__ void dup_hist(unsigned int* s_hist, unsigned int tile, bool valid, int npass) {
    if (valid) atomicAdd(&s_hist[tile & 255u], 1u);  // low digit: neighbouring entries are neighbouring tiles
    for (int p = 1; p < npass; ++p) {
        // higher digits are (almost) warp-uniform: one aggregated add per distinct value
        const unsigned d = valid ? ((tile >> (8 * p)) & 255u) : 0xffffu;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        if (valid && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&s_hist[p * 256 + d], (unsigned)__popc(peers));
    }
}

__global__ void __launch_bounds__(DUP_NT) duplicate_kernel(const unsigned int* __restrict__ pcount,
                                                           const int4* __restrict__ rsorted /*[Pc] depth order {x0|w<<16, y0, n, id}*/,
                                                           const int* __restrict__ start_sorted, int gx, int npass,
                                                           const unsigned int* __restrict__ zcount,
                                                           unsigned int* __restrict__ keys, int* __restrict__ vals,
                                                           unsigned int* __restrict__ hist /*[npass][256]*/) {
    __shared__ unsigned int s_hist[MAX_TPASS * 256];
    for (int i = threadIdx.x; i < npass * 256; i += DUP_NT) s_hist[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int P = (int)*pcount;  // emitting Gaussians (keygen_kernel)
    // phantom prefix: Z entries (tile 0, idx 0)
    const unsigned int Z = *zcount;
    for (long long e = (long long)blockIdx.x * DUP_NT + threadIdx.x; e < Z; e += (long long)gridDim.x * DUP_NT) {
        keys[e] = 0u;
        vals[e] = 0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && Z)
        for (int p = 0; p < npass; ++p) atomicAdd(&s_hist[p * 256], Z);
    const bool packable = gx < 4096;  // x0, y0 < 4096 fit the packed shuffle word (y0 < gy is not bounded by it)
    const long long nchunks = ((long long)P + DUP_NT - 1) / DUP_NT;
    for (long long chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const long long i = chunk * DUP_NT + threadIdx.x;
        int n = 0, start = 0, x0 = 0, y0 = 0, w = 1, id = 0;
        if (i < P) {
            start = start_sorted[i];
            const int4 r = rsorted[i];  // sequential: the last depth pass gathered the record
            x0 = r.x & 0xffff;
            w = (int)((unsigned)r.x >> 16);
            y0 = r.y;
            n = r.z;
            id = r.w;
        }
        // ---- footprints of up to DUP_COOP tiles: load-balanced expansion ----------------------------------
        // The warp's Gaussians own consecutive output runs; lane L emits entry e = base + L of the
        // concatenation, finds its source lane by binary search over the exclusive prefix of n (5 shuffles)
        // and fetches {output offset, x0 | y0, w | magic, id} with 4 more: stores are coalesced and no lane
        // idles behind a neighbour's larger footprint.  r / w for r w < 4096 is (r m) >> 12, m = 4096 / w + 1.
        const bool coop = packable && n > 0 && n <= DUP_COOP && w <= DUP_COOP && y0 < (1 << 19);
        const int nc = coop ? n : 0;
        int rel = nc;  // exclusive prefix of nc over the lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, rel, o);
            if (lane >= o) rel += v;
        }
        const int total = __shfl_sync(0xffffffffu, rel, 31);
        rel -= nc;
        const int delta = start - rel;
        const int pxy = x0 | (y0 << 12);
        const int wm = w | ((4096 / max(w, 1) + 1) << 8);
        for (int base = 0; base < total; base += 32) {
            const int e = base + lane;
            const bool valid = e < total;
            int src = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const int v = __shfl_sync(0xffffffffu, rel, src + step);
                if (v <= e) src += step;
            }
            const int srel = __shfl_sync(0xffffffffu, rel, src);
            const int sdelta = __shfl_sync(0xffffffffu, delta, src);
            const int sxy = __shfl_sync(0xffffffffu, pxy, src);
            const int swm = __shfl_sync(0xffffffffu, wm, src);
            const int sid = __shfl_sync(0xffffffffu, id, src);
            const int r = e - srel;
            const int sw = swm & 255, ry = (r * (swm >> 8)) >> 12, rx = r - ry * sw;
            const unsigned int tile = (unsigned)(((sxy >> 12) + ry) * gx + (sxy & 4095) + rx);
            if (valid) {
                keys[sdelta + e] = tile;
                vals[sdelta + e] = sid;
            }
            dup_hist(s_hist, tile, valid, npass);
        }
        // ---- larger footprints: the whole warp emits one Gaussian at a time (reference: one thread
        // serially writes up to T entries, sort_gaussian.cu:35-42)
        unsigned big = __ballot_sync(0xffffffffu, n > 0 && !coop);
        while (big) {
            const int src = __ffs(big) - 1;
            big &= big - 1;
            const int bn = __shfl_sync(0xffffffffu, n, src);
            const int bstart = __shfl_sync(0xffffffffu, start, src);
            const int bx0 = __shfl_sync(0xffffffffu, x0, src);
            const int by0 = __shfl_sync(0xffffffffu, y0, src);
            const int bw = __shfl_sync(0xffffffffu, w, src);
            const int bid = __shfl_sync(0xffffffffu, id, src);
            for (int e0 = 0; e0 < bn; e0 += 32) {
                const int e = e0 + lane;
                const bool valid = e < bn;
                const int ry = e / bw, rx = e - ry * bw;
                const unsigned int tile = (unsigned)((by0 + ry) * gx + bx0 + rx);
                if (valid) {
                    keys[bstart + e] = tile;
                    vals[bstart + e] = bid;
                }
                dup_hist(s_hist, tile, valid, npass);
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
// one onesweep pass (8-bit digit) over (key32, val32) pairs
// ------------------------------------------------------------------------------------------------
constexpr int RS_NT = 256;
constexpr int RS_WARPS = RS_NT / 32;
constexpr int RS_IPT = 16;               // keys per thread of the tile passes (4096 keys per CTA)
constexpr int RS_TILE = RS_NT * RS_IPT;

template <int IPT>
struct RsSmem {
    unsigned int keys[RS_NT * IPT];      // 16 KB exchange buffer (IPT = 16)
    int vals[RS_NT * IPT];               // 16 KB
    unsigned int whist[RS_WARPS][256];   //  8 KB per-warp digit counts -> per-warp offsets
    int gadj[256];                       // global base of the digit minus its tile-local start
    int scan_tmp[RS_NT / 32 + 1];
    int tile_id;
};

// m &= (lanes whose digit agrees with mine in bit B): LOP3 -> predicate, VOTE, SEL, LOP3 m & (bal ^ y)
template <int B>
__device__ __forceinline__ void peer_bit(unsigned d, unsigned& m) {
    unsigned bal, y;
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
        "and.b32 t, %2, %3;\n\t"
        "setp.ne.u32 p, t, 0;\n\t"
        "vote.sync.ballot.b32 %0, p, 0xffffffff;\n\t"
        "selp.b32 %1, 0, -1, p;\n\t}"
        : "=r"(bal), "=r"(y)
        : "r"(d), "n"(1u << B));
    asm("lop3.b32 %0, %0, %1, %2, 0x60;" : "+r"(m) : "r"(bal), "r"(y));
}

// DEVN: the element count is read from *n_dev (compacted Gaussians; the grid is sized for an upper
// bound).  LB: predecessors inspected per look-back round trip.  When all tiles of a pass are
// resident at once (the L2-resident depth passes: <= one wave) every tile publishes its aggregate
// at about the same time and the full prefixes then advance LB tiles per round trip while each
// waiting tile walks back LB per round trip: tile k resolves after ~k / (2 LB) round trips, so
// those passes want a wide window; the multi-wave tile passes resolve within the first window.
// IPT: keys per thread (16).  A/B on BASELINE config #3 (profiles/r1_ab_experiments.md): 8 keys per thread
// at 6 CTAs/SM and look-back windows of 16 / 32 were not faster for the L2-resident depth passes;
// what helped them was the single-lane wait below (gate_ns).
// GATHER (last depth pass): while scattering, the kernel fetches the 16-byte record {x0 | w << 16, y0, n, id} that
// keygen wrote at the emitter's compact slot (the value the depth passes carry; the pipeline's one random gather, over
// a dense array) and writes it at the Gaussian's depth-order position -- the offset scan and the duplication then
// stream sequentially.
template <bool DEVN, int LB, int NB, int IPT, bool GATHER = false>
__global__ void __launch_bounds__(RS_NT, IPT >= 16 ? 4 : 6) onesweep_kernel(int N, const unsigned int* __restrict__ n_dev, int shift, const unsigned int* __restrict__ kin,
                                                            const int* __restrict__ vin,
                                                            unsigned int* __restrict__ kout, int* __restrict__ vout,
                                                            const unsigned int* __restrict__ hist /*[256] this pass*/,
                                                            unsigned int* __restrict__ status /*[ntiles][256]*/,
                                                            unsigned int* __restrict__ ticket, int gate_ns,
                                                            const int4* __restrict__ rect = nullptr,
                                                            int4* __restrict__ rsorted = nullptr) {
    extern __shared__ __align__(16) unsigned char rs_raw[];
    RsSmem<IPT>& sm = *reinterpret_cast<RsSmem<IPT>*>(rs_raw);
    constexpr int TILE = RS_NT * IPT;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) sm.tile_id = (int)atomicAdd(ticket, 1u);  // tiles are ordered by start time
#pragma unroll
    for (int k = 0; k < 256 / 32; ++k) sm.whist[warp][lane + 32 * k] = 0;
    if (DEVN) N = (int)*n_dev;
    __syncthreads();
    const int tile = sm.tile_id;
    const long long tile_base = (long long)tile * TILE;
    const int valid = (int)min((long long)TILE, (long long)N - tile_base);
    if (DEVN && valid <= 0) return;  // the grid is sized for the host-side upper bound of N

    // ---- load (warp-striped: element order = warp, item, lane) --------------------------------
    unsigned int key[IPT];
    unsigned short rank[IPT];
    const int wbase = warp * (32 * IPT);
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const int e = wbase + i * 32 + lane;
        // kept ROTATED so that this pass's digit sits in the low byte (bit tests and digit extraction
        // need no shift); rotated back on the way out.  Padding sorts last: digit 255 in every pass.
        const unsigned int k = e < valid ? kin[tile_base + e] : 0xffffffffu;
        key[i] = __funnelshift_r(k, k, shift);
    }

    // ---- rank inside the warp with match.any (warp-ballot ranking) ------------------------------
    // All 16 match.any are issued first (independent, pipelined); the only serial chain left is one
    // shared-memory atomic per round on the warp's private digit counters.
    const unsigned lt_mask = (1u << lane) - 1u;
    // The peer mask (lanes holding the same digit) is built from 8 ballots, one per digit bit:
    // match.any.sync costs time proportional to the number of DISTINCT values in the warp (measured:
    // a pass over uniformly distributed digits ran 1.7x slower than one over 4 distinct digits).
    unsigned peers[IPT];
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const unsigned d = key[i];
        unsigned m = 0xffffffffu;
        // digit bits >= NB are zero in every real key of this pass (top tile-id digit; the host passes
        // one bit more than the keys need, so that the 0xFFFFFFFF padding of the last tile never
        // matches a real key)
        peer_bit<0>(d, m);
        if constexpr (NB > 1) peer_bit<1>(d, m);
        if constexpr (NB > 2) peer_bit<2>(d, m);
        if constexpr (NB > 3) peer_bit<3>(d, m);
        if constexpr (NB > 4) peer_bit<4>(d, m);
        if constexpr (NB > 5) peer_bit<5>(d, m);
        if constexpr (NB > 6) peer_bit<6>(d, m);
        if constexpr (NB > 7) peer_bit<7>(d, m);
        peers[i] = m;
    }
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const unsigned d = key[i] & 255u;
        const int leader = __ffs(peers[i]) - 1;
        unsigned base = 0;
        if (lane == leader) base = atomicAdd(&sm.whist[warp][d], (unsigned)__popc(peers[i]));
        __syncwarp();  // orders the counters between rounds (stability of the rank)
        base = __shfl_sync(0xffffffffu, base, leader);
        rank[i] = (unsigned short)(base + __popc(peers[i] & lt_mask));
    }
    // payload: fetched now (the match masks are dead), lands while the digit offsets are resolved
    int val[IPT];
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const int e = wbase + i * 32 + lane;
        val[i] = e < valid ? vin[tile_base + e] : 0;
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
    if (tid == 255) count -= (unsigned)(TILE - valid);  // padding keys are not real
    unsigned int* my_status = status + (size_t)tile * 256 + tid;
    if (tile != 0) st_relaxed(my_status, rs_agg(count));  // as early as possible: successors sum it

    // tile-local exclusive digit offsets; tile 0 also scans the global histogram and folds the
    // global digit base into the prefix it publishes, so every other tile receives it through
    // the look-back and needs no scan of its own
    int tot;
    const unsigned int lbase = (unsigned)block_excl_scan<RS_NT>((int)count, sm.scan_tmp, tot);
    unsigned int excl = 0;
    if (tile == 0) {
        excl = (unsigned)block_excl_scan<RS_NT>((int)hist[tid], sm.scan_tmp, tot);
        st_relaxed(my_status, rs_pre(excl + count));
    } else {
        // decoupled look-back, LB predecessors per round trip (independent loads in flight);
        // a serial walk costs one L2 latency per predecessor and dominated the pass
        int j = tile - 1;
        while (true) {
            unsigned int v[LB];
#pragma unroll
            for (int k = 0; k < LB; ++k)
                v[k] = (j - k >= 0) ? ld_relaxed(status + (size_t)(j - k) * 256 + tid) : RS_PRE_BIT;
            int used = 0;
            bool done = false;
#pragma unroll
            for (int k = 0; k < LB; ++k) {
                if (!done && used == k && v[k] != 0u) {
                    excl += rs_value(v[k]);
                    used = k + 1;
                    done = rs_is_pre(v[k]);
                }
            }
            if (done) break;
            j -= used;
            if (gate_ns > 0) {
                // nothing published by tile j yet: ONE lane per warp waits for it (a tile's 256 words are
                // written together), so waiting warps do not take issue slots from the tiles still ranking
                if (__all_sync(0xffffffffu, used == 0)) {
                    if (lane == 0)
                        while (ld_relaxed(status + (size_t)j * 256 + tid) == 0u) __nanosleep(gate_ns);
                    __syncwarp();
                }
            } else if (used == 0) {
                __nanosleep(20);
            }
        }
        st_relaxed(my_status, rs_pre(excl + count));
    }
    sm.gadj[tid] = (int)excl - (int)lbase;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) sm.whist[w][tid] += lbase;  // warp offset inside the tile
    __syncthreads();

    // ---- exchange through shared memory (tile-local sorted order) ------------------------------
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const unsigned d = key[i] & 255u;
        const unsigned pos = sm.whist[warp][d] + rank[i];
        sm.keys[pos] = key[i];
        sm.vals[pos] = val[i];
    }
    __syncthreads();

    // ---- coalesced scatter ----------------------------------------------------------------------
#pragma unroll 4
    for (int j = tid; j < valid; j += RS_NT) {
        const unsigned int k = sm.keys[j];  // rotated key
        const unsigned d = k & 255u;
        const long long g = (long long)sm.gadj[d] + j;
        if (GATHER) {
            rsorted[g] = rect[sm.vals[j]];  // 16-byte record of the emitter whose compact slot the passes carried
        } else {
            kout[g] = __funnelshift_l(k, k, shift);
            vout[g] = sm.vals[j];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// phase 2f: tile ranges from the sorted tile ids (sort_gaussian.cu:45-71); tile_range pre-zeroed
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tile_range_kernel(int M, const unsigned int* __restrict__ keys,
                                                         int2* __restrict__ tile_range, int T) {
    // four keys per thread: one 16-byte load + the left neighbour
    const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i0 >= M) return;
    unsigned int k[4];
    if (i0 + 3 < M) {
        const uint4 v = *reinterpret_cast<const uint4*>(keys + i0);
        k[0] = v.x; k[1] = v.y; k[2] = v.z; k[3] = v.w;
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) k[e] = i0 + e < M ? keys[i0 + e] : 0u;
    }
    unsigned int prev = i0 > 0 ? keys[i0 - 1] : 0u;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const long long i = i0 + e;
        if (i >= M) break;
        const unsigned int cur = k[e];
        if (cur < (unsigned)T) {  // always true for keys produced by duplicate_kernel
            if (i == 0) tile_range[cur].x = 0;
            if (i == M - 1) tile_range[cur].y = M;
            if (i > 0 && prev != cur) {
                if (prev < (unsigned)T) tile_range[prev].y = (int)i;
                tile_range[cur].x = (int)i;
            }
        }
        prev = cur;
    }
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// significant bits of the tile ids 0..T-1, and the digit passes over them
static int tile_bits(int T) {
    int bits = 0;
    while (bits < 31 && (1ll << bits) < (long long)T) ++bits;
    return bits;
}
static int tile_passes(int T) { return (tile_bits(T) + 7) / 8; }

struct SortLayout {
    size_t dkeys[2], dvals[2], rect, rsorted, start, tkeys[2], tvals, zero0, hist, ticket, zcount, pcount, status_k, status_s, status_d, status_t,
        total;
    int tpass, ntiles_d, ntiles_t, nb_scan, nchunks_k;
};

static SortLayout sort_layout(int P, long long M, int T) {
    SortLayout L;
    L.tpass = tile_passes(T);
    L.ntiles_d = (int)(((long long)P + RS_TILE - 1) / RS_TILE);
    L.ntiles_t = (int)((M + RS_TILE - 1) / RS_TILE);
    L.nb_scan = (int)(((long long)P + SC_TILE - 1) / SC_TILE);
    L.nchunks_k = (int)(((long long)P + 255) / 256) + 1;  // >= keygen grid (<= ceil(P / 256) CTAs)
    size_t off = 0;
    auto take = [&](size_t bytes) {
        const size_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    };
    for (int k = 0; k < 2; ++k) L.dkeys[k] = take((size_t)P * 4);
    for (int k = 0; k < 2; ++k) L.dvals[k] = take((size_t)P * 4);
    L.rect = take((size_t)P * 16);
    L.rsorted = take((size_t)P * 16);
    L.start = take((size_t)P * 4);
    for (int k = 0; k < 2; ++k) L.tkeys[k] = take((size_t)M * 4);
    L.tvals = take((size_t)M * 4);
    // zero-initialised region: hist | ticket | zcount | status
    L.zero0 = off;
    L.hist = take((size_t)(4 + MAX_TPASS) * 256 * 4);
    L.ticket = take((size_t)(4 + MAX_TPASS + 2) * 4);  // one per onesweep pass + keygen + offset scan
    L.zcount = take(4);
    L.pcount = take(4);
    L.status_k = take((size_t)L.nchunks_k * 4);
    L.status_s = take((size_t)(L.nb_scan + 1) * 4);
    L.status_d = take((size_t)4 * L.ntiles_d * 256 * 4);
    L.status_t = take((size_t)L.tpass * L.ntiles_t * 256 * 4);
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

// Phase 1.  The 64-bit total M = sum(max(tiles, 0)) is copied asynchronously into *total_host
// (pinned host memory): the caller synchronises the stream before reading it.  If offsets is not
// NULL it also receives the inclusive int32 cumsum (torch.cumsum equivalent; the sort itself no
// longer needs it).
int msb_sort_scan(const int32_t* tiles, int P, int32_t* offsets, long long* total_host, void* ws, size_t ws_bytes,
                  void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (P < 0 || !total_host) return set_error(MSB_ERR_ARG, "sort_scan: bad argument");
    if (P == 0) {
        *total_host = 0;
        return MSB_OK;
    }
    if (!tiles || !ws) return set_error(MSB_ERR_ARG, "sort_scan: null pointer");
    if (ws_bytes < msb_sort_scan_workspace_bytes(P)) return set_error(MSB_ERR_WORKSPACE, "sort_scan: workspace too small");
    const int nb = (P + SC_TILE - 1) / SC_TILE;
    long long* bsum = reinterpret_cast<long long*>(ws);
    long long* total_dev = bsum + nb;
    scan_block_sums_kernel<<<nb, SC_NT, 0, st>>>(P, tiles, bsum);
    scan_spine_kernel<<<1, 1024, 0, st>>>(nb, bsum, total_dev);
    if (offsets) scan_apply_kernel<false><<<nb, SC_NT, 0, st>>>(P, tiles, bsum, offsets);
    int rc = check_launch("sort_scan");
    if (rc) return rc;
    cudaError_t e = cudaMemcpyAsync(total_host, total_dev, sizeof(long long), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return set_error((int)e, "sort_scan: cudaMemcpyAsync failed");
    return MSB_OK;
}

// ------------------------------------------------------------------------------------------------
// Level-1 drop-in of the two sort-stage functions of msplat._C (integration/_C.py): the reference's
// Python keeps torch.cumsum / torch.sort / torch.gather between them (msplat/sort_gaussian.py:42-52)
// ------------------------------------------------------------------------------------------------
// computeGaussianKeyCUDAKernel (src/sort_gaussian.cu:17-43): key = (tile << 32) | depth bits at slots
// [cumsum[i-1], ...), rows then columns.  A warp takes 32 Gaussians and writes each footprint cooperatively
// (the reference: one thread, serially).  Divergences as in the fused sort: depth bits masked to 32 bits,
// at most cumsum[i] - cumsum[i-1] entries per Gaussian; unwritten slots stay (0, 0) (zeroed by the caller).
__global__ void __launch_bounds__(256) gaussian_key_kernel(int P, const float2* __restrict__ uv,
                                                           const float* __restrict__ depth,
                                                           const int* __restrict__ radius,
                                                           const int* __restrict__ cumsum, int gx, int gy,
                                                           long long* __restrict__ keys, int* __restrict__ idx_out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int n = 0, start = 0, x0 = 0, y0 = 0, w = 1;
    unsigned int dbits = 0;
    if (i < P) {
        start = i == 0 ? 0 : cumsum[i - 1];
        const int slots = cumsum[i] - start;
        n = keygen_count(slots, radius[i], uv[i], gx, gy, x0, y0, w);
        dbits = __float_as_uint(depth[i]);
    }
    unsigned todo = __ballot_sync(0xffffffffu, n > 0);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int bn = __shfl_sync(0xffffffffu, n, src), bstart = __shfl_sync(0xffffffffu, start, src);
        const int bx0 = __shfl_sync(0xffffffffu, x0, src), by0 = __shfl_sync(0xffffffffu, y0, src);
        const int bw = __shfl_sync(0xffffffffu, w, src);
        const unsigned int bd = __shfl_sync(0xffffffffu, dbits, src);
        const int bi = (int)(i - lane + src);
        for (int e = lane; e < bn; e += 32) {
            const int ry = e / bw, rx = e - ry * bw;
            const long long tile = (long long)(by0 + ry) * gx + bx0 + rx;
            keys[bstart + e] = (tile << 32) | (long long)bd;
            idx_out[bstart + e] = bi;
        }
    }
}

// computeTileGaussianRangeCUDAKernel (src/sort_gaussian.cu:45-71) on sorted 64-bit keys; tile_range pre-zeroed
__global__ void __launch_bounds__(256) tile_range64_kernel(long long M, const long long* __restrict__ keys,
                                                           int2* __restrict__ tile_range, int T) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const int cur = (int)(keys[i] >> 32);
    if (cur < 0 || cur >= T) return;
    if (i == 0) tile_range[cur].x = 0;
    if (i == M - 1) tile_range[cur].y = (int)M;
    if (i == 0) return;
    const int prev = (int)(keys[i - 1] >> 32);
    if (prev != cur) {
        if (prev >= 0 && prev < T) tile_range[prev].y = (int)i;
        tile_range[cur].x = (int)i;
    }
}

// Number of 8-bit digit passes over the 64-bit key: 4 depth passes (on the Gaussians) plus the
// tile-id passes (on the duplicates).  A batch of `views` views sorts (view * T + tile) ids.
int msb_sort_num_passes_views(int W, int H, int views) {
    const long long gx = (W + MSB_TILE - 1) / MSB_TILE, gy = (H + MSB_TILE - 1) / MSB_TILE;
    const long long T = gx * gy * (views > 0 ? views : 1);
    return 4 + tile_passes((int)min(T, (long long)INT32_MAX));
}
int msb_sort_num_passes(int W, int H) { return msb_sort_num_passes_views(W, H, 1); }

size_t msb_sort_workspace_bytes_views(int P, int views, long long M, int W, int H) {
    const long long gx = (W + MSB_TILE - 1) / MSB_TILE, gy = (H + MSB_TILE - 1) / MSB_TILE;
    const long long Pt = (long long)P * views, T = gx * gy * views;
    if (M <= 0 || P <= 0 || views <= 0 || Pt > INT32_MAX || T > INT32_MAX) return 256;
    return sort_layout((int)Pt, M, (int)T).total;
}
size_t msb_sort_workspace_bytes(int P, long long M, int W, int H) { return msb_sort_workspace_bytes_views(P, 1, M, W, H); }

// Phase 2 for a batch of `views` views over the same P Gaussians (views = 1: the reference's call).
// uv/depth/radius/tiles hold [views, P] entries, view-major.  idx_sorted [M] receives view * P + index,
// tile_range [views * T, 2] indexes idx_sorted; M = sum of tiles over the whole batch, < 2^31.
int msb_sort_gaussian_views(const float* uv, const float* depth, const int32_t* radius, const int32_t* tiles, int Pv,
                            int views, long long M, int W, int H, int32_t* idx_sorted, int32_t* tile_range, void* ws,
                            size_t ws_bytes, int sm_count, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int gx = (W + MSB_TILE - 1) / MSB_TILE, gy = (H + MSB_TILE - 1) / MSB_TILE;
    if (Pv < 0 || views <= 0 || M < 0 || W <= 0 || H <= 0 || !tile_range)
        return set_error(MSB_ERR_ARG, "sort_gaussian: bad argument");
    if ((long long)Pv * views > INT32_MAX || (long long)gx * gy * views > INT32_MAX)
        return set_error(MSB_ERR_RANGE, "sort_gaussian: views * P and views * tiles must stay below 2^31");
    if (gx >= 65536) return set_error(MSB_ERR_RANGE, "sort_gaussian: more than 65535 tile columns (W >= 2^20 pixels)");
    // int32 positions, like the reference's int32 cumsum (msplat/sort_gaussian.py:42)
    if (M > (long long)INT32_MAX) return set_error(MSB_ERR_RANGE, "sort_gaussian: more than 2^31 - 1 tile intersections");
    const int P = Pv * views;
    const int T = gx * gy * views;
    cudaError_t e = cudaMemsetAsync(tile_range, 0, (size_t)T * 2 * sizeof(int32_t), st);
    if (e != cudaSuccess) return set_error((int)e, "sort_gaussian: memset tile_range failed");
    if (M == 0 || P == 0) return MSB_OK;
    if (!uv || !depth || !radius || !tiles || !idx_sorted || !ws)
        return set_error(MSB_ERR_ARG, "sort_gaussian: null pointer");
    const SortLayout L = sort_layout(P, M, T);
    if (ws_bytes < L.total) return set_error(MSB_ERR_WORKSPACE, "sort_gaussian: workspace too small");
    unsigned char* base = reinterpret_cast<unsigned char*>(ws);
    auto U32 = [&](size_t o) { return reinterpret_cast<unsigned int*>(base + o); };
    auto I32 = [&](size_t o) { return reinterpret_cast<int*>(base + o); };
    unsigned int* hist = U32(L.hist);
    unsigned int* ticket = U32(L.ticket);
    unsigned int* zcount = U32(L.zcount);
    e = cudaMemsetAsync(base + L.zero0, 0, L.total - L.zero0, st);
    if (e != cudaSuccess) return set_error((int)e, "sort_gaussian: memset workspace failed");
    const int sms = sm_count > 0 ? sm_count : 148;
    unsigned int* pcount = U32(L.pcount);

    // 2a: depth keys of the emitting Gaussians, compacted in index order; *pcount = Pc.  At most one
    // wave of CTAs (4 per SM); each warp owns `seg` consecutive Gaussians.
    const long long warps_needed = ((long long)P + 31) / 32;
    const unsigned kgrid = (unsigned)max(1ll, min((warps_needed + KG_WARPS - 1) / KG_WARPS, (long long)sms * 4));
    const long long per_warp = ((long long)P + (long long)kgrid * KG_WARPS - 1) / ((long long)kgrid * KG_WARPS);
    const int seg = (int)((per_warp + 31) / 32 * 32);
    if ((size_t)kgrid > (size_t)L.nchunks_k) return set_error(MSB_ERR_WORKSPACE, "sort_gaussian: keygen status");
    keygen_kernel<<<kgrid, KG_NT, 0, st>>>(P, max(Pv, 1), seg, reinterpret_cast<const float2*>(uv), depth, radius, tiles, gx, gy,
                                           U32(L.dkeys[0]), I32(L.dvals[0]), reinterpret_cast<int4*>(base + L.rect), hist,
                                           zcount,
                                           U32(L.status_k), ticket + 4 + MAX_TPASS, pcount);
    int rc = check_launch("sort_gaussian/keygen");
    if (rc) return rc;

    static_assert(sizeof(RsSmem<RS_IPT>) <= 48 * 1024, "onesweep shared memory");
    // 2b: four depth-digit passes over the Pc emitting Gaussians (ends in buffer 0).  Grids are sized
    // for the host-side bound min(P, M) >= Pc; surplus CTAs exit after taking their ticket.
    const long long pc_max = min((long long)P, M);
    const int nb_scan = (int)((pc_max + SC_TILE - 1) / SC_TILE);
    const int ntiles_d = (int)((pc_max + RS_TILE - 1) / RS_TILE);
    static const int gate_ns = [] { const char* e = getenv("MSB_SORT_GATE"); return e ? atoi(e) : 100; }();
    for (int p = 0; p < 4; ++p) {
        unsigned int* kin = U32(L.dkeys[p % 2]);
        int* vin = I32(L.dvals[p % 2]);
        unsigned int* kout = U32(L.dkeys[(p + 1) % 2]);
        int* vout = I32(L.dvals[(p + 1) % 2]);
        unsigned int* stat = U32(L.status_d) + (size_t)p * L.ntiles_d * 256;
        if (p < 3)
            onesweep_kernel<true, 8, 8, RS_IPT><<<ntiles_d, RS_NT, sizeof(RsSmem<RS_IPT>), st>>>(
                P, pcount, 8 * p, kin, vin, kout, vout, hist + p * 256, stat, ticket + p, gate_ns);
        else  // the last depth pass writes the depth-ordered {rectangle, id} records instead of (key, id)
            onesweep_kernel<true, 8, 8, RS_IPT, true><<<ntiles_d, RS_NT, sizeof(RsSmem<RS_IPT>), st>>>(
                P, pcount, 8 * p, kin, vin, kout, vout, hist + p * 256, stat, ticket + p, gate_ns,
                reinterpret_cast<const int4*>(base + L.rect), reinterpret_cast<int4*>(base + L.rsorted));
        rc = check_launch("sort_gaussian/onesweep(depth)");
        if (rc) return rc;
    }
    const int4* rsorted = reinterpret_cast<const int4*>(base + L.rsorted);

    // 2c: start offsets in depth order (exclusive scan of the emitted counts, + Z)
    scan_offsets_kernel<<<nb_scan, SC_NT, 0, st>>>(pcount, rsorted, zcount, I32(L.start), U32(L.status_s),
                                                   ticket + 4 + MAX_TPASS + 1);
    rc = check_launch("sort_gaussian/offsets");
    if (rc) return rc;

    // 2d: duplication; values ping-pong so that the last tile pass lands in idx_sorted
    unsigned int* tk[2] = {U32(L.tkeys[0]), U32(L.tkeys[1])};
    int* tv[2];
    tv[L.tpass % 2] = idx_sorted;
    tv[(L.tpass + 1) % 2] = I32(L.tvals);
    unsigned int* thist = hist + 4 * 256;
    const unsigned dgrid = (unsigned)max(1ll, min((pc_max + DUP_NT - 1) / DUP_NT, (long long)sms * 8));
    duplicate_kernel<<<dgrid, DUP_NT, 0, st>>>(pcount, rsorted, I32(L.start), gx, L.tpass, zcount, tk[0], tv[0], thist);
    rc = check_launch("sort_gaussian/duplicate");
    if (rc) return rc;

    // 2e: tile-digit passes over the duplicates
    const int tbits = tile_bits(T);
    for (int p = 0; p < L.tpass; ++p) {
        // digit bits the ranking compares: those the tile ids use in this digit, plus one (see the kernel)
        const int nb = min(8, max(1, tbits - 8 * p) + 1);
        const unsigned int* kin = tk[p % 2];
        const int* vin = tv[p % 2];
        unsigned int* kout = tk[(p + 1) % 2];
        int* vout = tv[(p + 1) % 2];
        const unsigned int* h = thist + p * 256;
        unsigned int* stat = U32(L.status_t) + (size_t)p * L.ntiles_t * 256;
        unsigned int* tick = ticket + 4 + p;
#define MSB_TILE_PASS(NB_) \
    onesweep_kernel<false, 8, NB_, RS_IPT><<<L.ntiles_t, RS_NT, sizeof(RsSmem<RS_IPT>), st>>>((int)M, nullptr, 8 * p, kin, vin, kout, \
                                                                             vout, h, stat, tick, gate_ns)
        switch (nb) {
            case 2: MSB_TILE_PASS(2); break;
            case 3: MSB_TILE_PASS(3); break;
            case 4: MSB_TILE_PASS(4); break;
            case 5: MSB_TILE_PASS(5); break;
            case 6: MSB_TILE_PASS(6); break;
            case 7: MSB_TILE_PASS(7); break;
            default: MSB_TILE_PASS(8); break;
        }
#undef MSB_TILE_PASS
        rc = check_launch("sort_gaussian/onesweep(tile)");
        if (rc) return rc;
    }
    tile_range_kernel<<<(unsigned)((M + 1023) / 1024), 256, 0, st>>>((int)M, tk[L.tpass % 2],
                                                                  reinterpret_cast<int2*>(tile_range), T);
    return check_launch("sort_gaussian/tile_range");
}

// msplat._C.compute_gaussian_key: cumsum [P] = inclusive int32 cumsum of tiles_touched, M = cumsum[P-1];
// keys [M] int64 and idx [M] int32 are zeroed and then filled.
int msb_compute_gaussian_key(const float* uv, const float* depth, const int32_t* radius, const int32_t* cumsum, int P,
                             long long M, int W, int H, long long* keys, int32_t* idx, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (P < 0 || M < 0 || W <= 0 || H <= 0) return set_error(MSB_ERR_ARG, "compute_gaussian_key: bad argument");
    if (M == 0 || P == 0) return MSB_OK;
    if (!uv || !depth || !radius || !cumsum || !keys || !idx) return set_error(MSB_ERR_ARG, "compute_gaussian_key: null pointer");
    if (M > (long long)INT32_MAX) return set_error(MSB_ERR_RANGE, "compute_gaussian_key: more than 2^31 - 1 tile intersections");
    cudaError_t e = cudaMemsetAsync(keys, 0, (size_t)M * sizeof(long long), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(idx, 0, (size_t)M * sizeof(int32_t), st);
    if (e != cudaSuccess) return set_error((int)e, "compute_gaussian_key: memset failed");
    const int gx = (W + MSB_TILE - 1) / MSB_TILE, gy = (H + MSB_TILE - 1) / MSB_TILE;
    gaussian_key_kernel<<<(unsigned)(((long long)P + 255) / 256), 256, 0, st>>>(
        P, reinterpret_cast<const float2*>(uv), depth, radius, cumsum, gx, gy, keys, idx);
    return check_launch("compute_gaussian_key");
}

// msplat._C.compute_tile_gaussian_range: tile_range [T,2] is zeroed and then filled from the sorted keys.
int msb_compute_tile_gaussian_range(const long long* keys_sorted, long long M, int W, int H, int32_t* tile_range,
                                    void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (M < 0 || W <= 0 || H <= 0 || !tile_range) return set_error(MSB_ERR_ARG, "compute_tile_gaussian_range: bad argument");
    const int T = ((W + MSB_TILE - 1) / MSB_TILE) * ((H + MSB_TILE - 1) / MSB_TILE);
    cudaError_t e = cudaMemsetAsync(tile_range, 0, (size_t)T * 2 * sizeof(int32_t), st);
    if (e != cudaSuccess) return set_error((int)e, "compute_tile_gaussian_range: memset failed");
    if (M == 0) return MSB_OK;
    if (!keys_sorted) return set_error(MSB_ERR_ARG, "compute_tile_gaussian_range: null pointer");
    tile_range64_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(M, keys_sorted, reinterpret_cast<int2*>(tile_range), T);
    return check_launch("compute_tile_gaussian_range");
}

// Phase 2.  idx_sorted[M] (int32) and tile_range[T, 2] (int32) are outputs owned by the caller.
int msb_sort_gaussian(const float* uv, const float* depth, const int32_t* radius, const int32_t* tiles, int P,
                      long long M, int W, int H, int32_t* idx_sorted, int32_t* tile_range, void* ws, size_t ws_bytes,
                      int sm_count, void* stream) {
    return msb_sort_gaussian_views(uv, depth, radius, tiles, P, 1, M, W, H, idx_sorted, tile_range, ws, ws_bytes,
                                   sm_count, stream);
}

}  // extern "C"
