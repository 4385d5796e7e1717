// binning.cu -- tile-bucketed binning: replaces the reference's prefix sum (K2),
// duplicateWithKeys (K3), 64-bit global radix sort (K4) and identifyTileRanges (K5)
// (rasterizer_impl.cu:70-138, 278-315).
//
// The reference sorts R (tile | depth) 64-bit keys with a 6-pass device-wide onesweep
// radix sort (~152 B of HBM traffic per instance).  Here the tile id is never sorted:
//   1. preprocess counts instances per tile (tile_count),
//   2. tile_scan: exclusive scan over the tiles -> ranges[tile] = [first,last), cursors,
//      num_rendered (no per-Gaussian scan, no host round trip for launch sizes),
//   3. scatter: every (Gaussian, tile) instance claims a slot in its tile's bucket and
//      stores depth_bits<<32 | gaussian_idx            (8 B written per instance),
//   4. per-tile sort in shared memory by that 64-bit value (monotone depth-bucket sort),
//      writing the gaussian index list                  (8 B read + 4 B written).
// Sorting by (depth bits, gaussian idx) reproduces the order of the reference's stable
// sort exactly: ties in (tile, depth) keep emission order, which is ascending Gaussian
// index (rasterizer_impl.cu:88-108).  point_list and ranges are therefore bit-identical
// to the reference at ~20 B per instance.
#include "surfel_common.cuh"
#include "surfel_kernels.h"

namespace srf {

constexpr int kSmallCap = 2048;   // entries sorted by the 256-thread per-tile kernel (two 16 KB key arrays)

// Ascending bitonic network over data[0..n) (n arbitrary; indices >= n act as +inf and
// are never touched).  All compare-exchanges are ascending ("normalised" network), so
// padding needs no storage.  Works on shared or global memory; the CTA must call it
// convergently.
template <typename Ptr>
__device__ __forceinline__ void bitonic_sort_cta(Ptr data, int n, int tid, int nthreads) {
    if (n < 2) return;
    int m = 1;
    while (m < n) m <<= 1;
    const int npairs = m >> 1;
    for (int k = 2, lk = 1; k <= m; k <<= 1, ++lk) {
        const int hk = k >> 1;
        for (int i = tid; i < npairs; i += nthreads) {
            const int blk = i >> (lk - 1), off = i & (hk - 1);
            const int lo = (blk << lk) + off, hi = (blk << lk) + k - 1 - off;
            if (hi < n) {
                const uint64_t x = data[lo], y = data[hi];
                if (x > y) { data[lo] = y; data[hi] = x; }
            }
        }
        __syncthreads();
        for (int j = k >> 2; j > 0; j >>= 1) {
            for (int i = tid; i < npairs; i += nthreads) {
                const int lo = 2 * i - (i & (j - 1)), hi = lo + j;
                if (hi < n) {
                    const uint64_t x = data[lo], y = data[hi];
                    if (x > y) { data[lo] = y; data[hi] = x; }
                }
            }
            __syncthreads();
        }
    }
}

constexpr int kSortSmallThreads = 256;

constexpr int kScanCache = 8192;   // tiles whose counts the scan keeps in shared memory (2048 x 1024 px images)

__global__ void __launch_bounds__(1024) tile_scan_kernel(BinArgs a) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_cnt[kScanCache];
    {   // view of this CTA: own geom / tile workspace and binning buffers
        const int view = blockIdx.y;
        a.tile_count = view_ptr(a.tile_count, view, a.tile_stride);
        a.ranges = view_ptr(a.ranges, view, a.tile_stride);
        a.counters = view_ptr(a.counters, view, a.tile_stride);
        a.big_list = view_ptr(a.big_list, view, a.tile_stride);
        a.tile_order = view_ptr(a.tile_order, view, a.tile_stride);
    }
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // The counters sit in one 256-byte block per tile (so that K1's atomics spread over the L2 slices): every read
    // is its own sector and a full L2 round trip.  The kernel needs each count four times (scan, ranges, LPT maximum,
    // LPT buckets); they are fetched ONCE, all loads of a thread in flight together, and kept in shared memory.
    const bool cached = a.ntiles <= kScanCache;
    if (cached) {
        for (int t = tid; t < a.ntiles; t += 1024) s_cnt[t] = a.tile_count[(size_t)t * SRF_TILE_CTR_STRIDE];
        __syncthreads();
    }
    auto count_of = [&](int t) -> uint32_t { return cached ? s_cnt[t] : a.tile_count[(size_t)t * SRF_TILE_CTR_STRIDE]; };
    const int chunk = (a.ntiles + 1023) / 1024;
    const int begin = min(tid * chunk, a.ntiles), end = min(begin + chunk, a.ntiles);
    uint32_t local = 0;
    for (int t = begin; t < end; ++t) local += count_of(t);
    uint32_t incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = s_warp[lane];
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += v;
        }
        s_warp[lane] = wi - w;  // exclusive prefix of warp totals
        if (lane == 31) {
            a.counters[0] = wi;  // num_rendered
            if (a.count_host != nullptr) {          // zero-copy read-back: no copy-engine operation in the stream
                a.count_host[blockIdx.y] = wi;
                __threadfence_system();
            }
        }
    }
    __syncthreads();
    uint32_t running = s_warp[wid] + incl - local;
    for (int t = begin; t < end; ++t) {
        const uint32_t c = count_of(t);
        a.ranges[t] = c ? make_uint2(running, running + c) : make_uint2(0u, 0u);
        a.tile_count[(size_t)t * SRF_TILE_CTR_STRIDE + 1] = running;   // bucket cursor
        if (c > (uint32_t)kSmallCap) {
            const uint32_t slot = atomicAdd(&a.counters[1], 1u);
            a.big_list[slot] = (uint32_t)t;
        }
        running += c;
    }
    // Longest-processing-time-first launch order for the per-tile blend CTAs: the grid is only
    // a few waves deep and per-tile work varies by >10x, so scheduling the heavy tiles first
    // removes most of the tail.  A 64-bucket counting sort on the instance count is enough.
    {
        __shared__ uint32_t s_bucket[64];
        __shared__ uint32_t s_max;
        if (tid < 64) s_bucket[tid] = 0;
        if (tid == 0) s_max = 0;
        __syncthreads();
        uint32_t lmax = 0;
        for (int t = tid; t < a.ntiles; t += 1024) lmax = max(lmax, count_of(t));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
        if (lane == 0) atomicMax(&s_max, lmax);
        __syncthreads();
        const uint32_t cmax = s_max + 1;
        for (int t = tid; t < a.ntiles; t += 1024) {
            const uint32_t c = count_of(t);
            const uint32_t bkt = 63u - (uint32_t)(((uint64_t)c * 64u) / cmax);   // heavy tiles -> low buckets
            atomicAdd(&s_bucket[bkt], 1u);
        }
        __syncthreads();
        if (wid == 0) {
            // exclusive scan of the 64 bucket sizes (two values per lane)
            const uint32_t v0 = s_bucket[2 * lane], v1 = s_bucket[2 * lane + 1];
            uint32_t sum = v0 + v1, inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            s_bucket[2 * lane] = inc - sum;
            s_bucket[2 * lane + 1] = inc - sum + v0;
        }
        __syncthreads();
        for (int t = tid; t < a.ntiles; t += 1024) {
            const uint32_t c = count_of(t);
            const uint32_t bkt = 63u - (uint32_t)(((uint64_t)c * 64u) / cmax);
            a.tile_order[atomicAdd(&s_bucket[bkt], 1u)] = (uint32_t)t;
        }
    }
}

__global__ void __launch_bounds__(256) scatter_kernel(BinArgs a) {
    {   // view of this CTA: own geom / tile workspace and binning buffers
        const int view = blockIdx.y;
        a.depths = view_ptr(a.depths, view, a.geom_stride);
        a.rects = view_ptr(a.rects, view, a.geom_stride);
        a.tile_count = view_ptr(a.tile_count, view, a.tile_stride);
        a.ranges = view_ptr(a.ranges, view, a.tile_stride);
        a.counters = view_ptr(a.counters, view, a.tile_stride);
        a.big_list = view_ptr(a.big_list, view, a.tile_stride);
        a.tile_order = view_ptr(a.tile_order, view, a.tile_stride);
        a.entries = view_ptr(a.entries, view, a.entries_stride);
        a.point_list = view_ptr(a.point_list, view, a.plist_stride);
    }
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int x0 = 0, y0 = 0, x1 = 0, y1 = 0, ntiles = 0;
    uint64_t key = 0;

    if (idx < a.P) {
        const uint2 r = a.rects[idx];   // (0,0,0,0) for culled Gaussians
        x0 = r.x & 0xffff; y0 = r.x >> 16; x1 = r.y & 0xffff; y1 = r.y >> 16;
        ntiles = (x1 - x0) * (y1 - y0);
        if (ntiles > 0) key = ((uint64_t)__float_as_uint(a.depths[idx]) << 32) | (uint32_t)idx;
    }
    const int kSerialMax = 8;
    if (ntiles > 0 && ntiles <= kSerialMax) {
        // claim all slots first (independent returning atomics stay in flight together), then store
        const int w = x1 - x0;
        uint32_t slot[kSerialMax];
#pragma unroll
        for (int k = 0; k < kSerialMax; ++k) {
            slot[k] = 0xffffffffu;
            if (k < ntiles) {
                const int ty = k / w, tx = k - ty * w;
                slot[k] = atomicAdd(&a.tile_count[(size_t)((y0 + ty) * a.gx + x0 + tx) * SRF_TILE_CTR_STRIDE + 1], 1u);
            }
        }
#pragma unroll
        for (int k = 0; k < kSerialMax; ++k)
            if (k < ntiles && slot[k] < a.capacity) a.entries[slot[k]] = key;
    }
    // rectangles of more than kSerialMax tiles: the warp shares the work, 32 tiles per step.  Every step is
    // "claim (returning atomic) -> store"; the stores of one step are issued only after the claims of the NEXT step
    // are in flight, so a warp with several large splats pays one atomic round trip, not one per splat.
    unsigned big = __ballot_sync(0xffffffffu, ntiles > kSerialMax);
    uint32_t pend_slot = 0xffffffffu;
    uint64_t pend_key = 0;
    while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const int bx0 = __shfl_sync(0xffffffffu, x0, src), by0 = __shfl_sync(0xffffffffu, y0, src);
        const int bx1 = __shfl_sync(0xffffffffu, x1, src), by1 = __shfl_sync(0xffffffffu, y1, src);
        const uint64_t bkey = __shfl_sync(0xffffffffu, key, src);
        const int w = bx1 - bx0, n = w * (by1 - by0);
        for (int t = lane; t < n + lane; t += 32) {          // same trip count for all lanes of the warp
            uint32_t slot = 0xffffffffu;
            if (t < n) {
                const int ty = t / w, tx = t - ty * w;
                slot = atomicAdd(&a.tile_count[(size_t)((by0 + ty) * a.gx + bx0 + tx) * SRF_TILE_CTR_STRIDE + 1], 1u);
            }
            if (pend_slot < a.capacity) a.entries[pend_slot] = pend_key;      // previous step's store
            pend_slot = slot; pend_key = bkey;
        }
    }
    if (pend_slot < a.capacity) a.entries[pend_slot] = pend_key;
}

// Per-tile sort: a monotone depth-bucket sort.
//   1. min / max depth of the tile,
//   2. every instance goes to one of NB depth buckets, b = trunc((d - dmin) * scale): a
//      monotone non-decreasing function of d, so ordering by (b, key) equals ordering by key,
//   3. instances are grouped by bucket in shared memory (histogram, scan, scatter),
//   4. each instance ranks itself inside its bucket by counting the smaller 64-bit keys (a
//      bucket holds a handful of instances unless depths collide), and writes its Gaussian index
//      to point_list[first + bucket_start + rank].
// About 50 instructions per instance and 6 barriers per tile, against ~550 instructions and
// 55 barriers for a full bitonic network at n ~ 600; the result is the same total order.
// NB must be a multiple of NT (each thread scans NB/NT consecutive buckets).
template <int NT, int NB>
__device__ __forceinline__ void depth_bucket_sort(const uint64_t* __restrict__ src, uint32_t* __restrict__ dst, int n,
                                                  uint64_t* s_keys, uint64_t* s_grouped, uint32_t* s_hist,
                                                  uint32_t* s_start, uint32_t* s_misc, int tid) {
    constexpr int PER = NB / NT;
    const int lane = tid & 31, wid = tid >> 5;
    for (int i = tid; i < NB; i += NT) s_hist[i] = 0;
    if (tid == 0) { s_misc[0] = 0xffffffffu; s_misc[1] = 0u; }
    __syncthreads();
    // view-space depths are positive, so their bit patterns order like the floats
    uint32_t dmin = 0xffffffffu, dmax = 0u;
    for (int i = tid; i < n; i += NT) {
        const uint64_t k = src[i];
        s_keys[i] = k;
        const uint32_t d = (uint32_t)(k >> 32);
        dmin = min(dmin, d); dmax = max(dmax, d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dmin = min(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
        dmax = max(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    }
    if (lane == 0) { atomicMin(&s_misc[0], dmin); atomicMax(&s_misc[1], dmax); }
    __syncthreads();
    const float fmin_ = __uint_as_float(s_misc[0]), fmax_ = __uint_as_float(s_misc[1]);
    const float range = fmax_ - fmin_;
    const float scale = range > 0.0f ? (float)NB / range : 0.0f;
    for (int i = tid; i < n; i += NT) {
        const float d = __uint_as_float((uint32_t)(s_keys[i] >> 32));
        const int b = min(NB - 1, (int)((d - fmin_) * scale));
        atomicAdd(&s_hist[b], 1u);
    }
    __syncthreads();
    {   // block-wide exclusive scan of the NB bucket sizes (PER consecutive buckets per thread)
        uint32_t v[PER], sum = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k) { v[k] = s_hist[tid * PER + k]; sum += v[k]; }
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_misc[2 + wid] = inc;
        __syncthreads();
        if (wid == 0) {
            const uint32_t w = (lane < NT / 32) ? s_misc[2 + lane] : 0u;
            uint32_t wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            if (lane < NT / 32) s_misc[2 + lane] = wi - w;
        }
        __syncthreads();
        uint32_t run = s_misc[2 + wid] + inc - sum;
#pragma unroll
        for (int k = 0; k < PER; ++k) { s_start[tid * PER + k] = run; s_hist[tid * PER + k] = run; run += v[k]; }
        if (tid == NT - 1) s_start[NB] = run;
    }
    __syncthreads();
    // Degenerate depth distribution (many instances share a bucket, in the limit all depths are equal): ranking inside
    // a bucket is quadratic in its size.  Past kMaxBucket the tile falls back to the comparison network
    // (n log^2 n whatever the keys are); the result is the same total order.
    {
        constexpr uint32_t kMaxBucket = 96;
        uint32_t big = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k) big = max(big, s_start[tid * PER + k + 1 > NB ? NB : tid * PER + k + 1] - s_start[tid * PER + k]);
        if (__syncthreads_or(big > kMaxBucket)) {
            bitonic_sort_cta(s_keys, n, tid, NT);
            for (int i = tid; i < n; i += NT) dst[i] = (uint32_t)s_keys[i];
            return;
        }
    }
    for (int i = tid; i < n; i += NT) {
        const uint64_t k = s_keys[i];
        const float d = __uint_as_float((uint32_t)(k >> 32));
        const int b = min(NB - 1, (int)((d - fmin_) * scale));
        s_grouped[atomicAdd(&s_hist[b], 1u)] = k;     // s_hist doubles as the bucket cursor
    }
    __syncthreads();
    for (int i = tid; i < n; i += NT) {
        const uint64_t k = s_grouped[i];
        const float d = __uint_as_float((uint32_t)(k >> 32));
        const int b = min(NB - 1, (int)((d - fmin_) * scale));
        const uint32_t st = s_start[b], en = s_start[b + 1];
        uint32_t rank = 0;
        for (uint32_t j = st; j < en; ++j) rank += (s_grouped[j] < k) ? 1u : 0u;
        dst[st + rank] = (uint32_t)k;
    }
}

constexpr int kSmallBuckets = 256;
constexpr int kBigBuckets = 2048;
constexpr int kBigSmemCap = 8192;   // instances the persistent kernel bucket-sorts in shared memory

__global__ void __launch_bounds__(kSortSmallThreads) sort_small_kernel(BinArgs a) {
    __shared__ uint64_t s_keys[kSmallCap];
    __shared__ uint64_t s_grouped[kSmallCap];
    __shared__ uint32_t s_hist[kSmallBuckets];
    __shared__ uint32_t s_start[kSmallBuckets + 1];
    __shared__ uint32_t s_misc[2 + 32];
    {   // view of this CTA: own geom / tile workspace and binning buffers
        const int view = blockIdx.y;
        a.depths = view_ptr(a.depths, view, a.geom_stride);
        a.rects = view_ptr(a.rects, view, a.geom_stride);
        a.tile_count = view_ptr(a.tile_count, view, a.tile_stride);
        a.ranges = view_ptr(a.ranges, view, a.tile_stride);
        a.counters = view_ptr(a.counters, view, a.tile_stride);
        a.big_list = view_ptr(a.big_list, view, a.tile_stride);
        a.tile_order = view_ptr(a.tile_order, view, a.tile_stride);
        a.entries = view_ptr(a.entries, view, a.entries_stride);
        a.point_list = view_ptr(a.point_list, view, a.plist_stride);
    }
    const uint2 r = a.ranges[blockIdx.x];
    const int n = (int)(r.y - r.x);
    const int tid = threadIdx.x;
    // rewind this tile's bucket cursor so that a re-run with a larger capacity (after an
    // optimistic-capacity overflow) can scatter again without repeating the scan
    if (tid == 0) a.tile_count[(size_t)blockIdx.x * SRF_TILE_CTR_STRIDE + 1] = r.x;
    if (n <= 0 || n > kSmallCap || r.y > a.capacity) return;
    const uint64_t* src = a.entries + r.x;
    uint32_t* dst = a.point_list + r.x;
    if (n == 1) {
        if (tid == 0) dst[0] = (uint32_t)src[0];
        return;
    }
    depth_bucket_sort<kSortSmallThreads, kSmallBuckets>(src, dst, n, s_keys, s_grouped, s_hist, s_start, s_misc, tid);
}

// Tiles with more than kSmallCap instances: a persistent 1024-thread kernel walks the list the
// scan produced.  Up to kBigSmemCap instances use the same depth-bucket sort (2048 buckets);
// beyond that (one tile holding > 8192 splats) an in-place bitonic network in L2/HBM.
__global__ void __launch_bounds__(1024) sort_big_kernel(BinArgs a) {
    extern __shared__ __align__(16) uint64_t s_dyn[];
    uint64_t* s_keys = s_dyn;
    uint64_t* s_grouped = s_dyn + kBigSmemCap;
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(s_dyn + 2 * kBigSmemCap);
    uint32_t* s_start = s_hist + kBigBuckets;
    uint32_t* s_misc = s_start + kBigBuckets + 1;
    {   // view of this CTA: own geom / tile workspace and binning buffers
        const int view = blockIdx.y;
        a.depths = view_ptr(a.depths, view, a.geom_stride);
        a.rects = view_ptr(a.rects, view, a.geom_stride);
        a.tile_count = view_ptr(a.tile_count, view, a.tile_stride);
        a.ranges = view_ptr(a.ranges, view, a.tile_stride);
        a.counters = view_ptr(a.counters, view, a.tile_stride);
        a.big_list = view_ptr(a.big_list, view, a.tile_stride);
        a.tile_order = view_ptr(a.tile_order, view, a.tile_stride);
        a.entries = view_ptr(a.entries, view, a.entries_stride);
        a.point_list = view_ptr(a.point_list, view, a.plist_stride);
    }
    const int tid = threadIdx.x;
    const uint32_t nbig = a.counters[1];
    for (uint32_t b = blockIdx.x; b < nbig; b += gridDim.x) {
        const uint2 r = a.ranges[a.big_list[b]];
        const int n = (int)(r.y - r.x);
        if (r.y > a.capacity) continue;
        uint64_t* src = a.entries + r.x;
        uint32_t* dst = a.point_list + r.x;
        __syncthreads();
        if (n <= kBigSmemCap) {
            depth_bucket_sort<1024, kBigBuckets>(src, dst, n, s_keys, s_grouped, s_hist, s_start, s_misc, tid);
        } else {
            bitonic_sort_cta(src, n, tid, 1024);
            for (int i = tid; i < n; i += 1024) dst[i] = (uint32_t)src[i];
        }
    }
}

cudaError_t launch_tile_scan(const BinArgs& a, cudaStream_t stream) {
    if (a.nviews <= 0) return cudaSuccess;
    prof_start(K_TILE_SCAN, stream);
    tile_scan_kernel<<<dim3(1, a.nviews), 1024, 0, stream>>>(a);
    prof_stop(K_TILE_SCAN, stream);
    return cudaGetLastError();
}

cudaError_t launch_bin_and_sort(const BinArgs& a, cudaStream_t stream) {
    const size_t big_smem = (size_t)2 * kBigSmemCap * sizeof(uint64_t) + (2 * kBigBuckets + 1 + 2 + 32 + 4) * sizeof(uint32_t);
    if (a.P <= 0 || a.nviews <= 0) return cudaSuccess;
    // the opt-in is per device (and cheap), so it is made on every call for the current device
    cudaError_t e = cudaFuncSetAttribute(sort_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big_smem);
    if (e != cudaSuccess) return e;
    prof_start(K_SCATTER, stream);
    scatter_kernel<<<dim3((a.P + 255) / 256, a.nviews), 256, 0, stream>>>(a);
    prof_stop(K_SCATTER, stream);
    prof_start(K_SORT_SMALL, stream);
    sort_small_kernel<<<dim3(a.ntiles, a.nviews), kSortSmallThreads, 0, stream>>>(a);
    prof_stop(K_SORT_SMALL, stream);
    // big tiles (> kSmallCap instances) are rare: a few persistent CTAs per view walk the list the scan produced
    const int per_view = max(8, sm_count() / a.nviews);
    prof_start(K_SORT_BIG, stream);
    sort_big_kernel<<<dim3(per_view, a.nviews), 1024, big_smem, stream>>>(a);
    prof_stop(K_SORT_BIG, stream);
    return cudaGetLastError();
}

}  // namespace srf
