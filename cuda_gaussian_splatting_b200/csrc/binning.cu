// binning.cu — device-wide exclusive scan, duplicateWithKeys and tile-range identification.
//
// Replaces, in the reference (rasterizer/sorting.cu): the libtorch cumsum + .item() + zeros +
// slice copy (:145-152), k_fill_sort_pairs (:30-72) and k_compute_tile_ranges (:82-109).
//
//  * scan: single-pass chained scan with decoupled look-back (one read + one write of N int32,
//    8 B/Gaussian), dynamic block ticket so look-back can never wait on an unscheduled block.
//  * duplicateWithKeys: warp-level load-balanced expansion. A warp owns 32 consecutive
//    Gaussians whose output slots are contiguous (offsets are an exclusive scan), so output slot
//    k is resolved to its owner lane with a 5-step shuffle binary search and every 8-byte key /
//    4-byte value store of the warp is fully coalesced, regardless of how many tiles a single
//    Gaussian covers (the reference serialises all tiles of a Gaussian in one thread).
#include "common.cuh"

namespace cugs {

// ------------------------------------------------------------------------------------------------
// scan
// ------------------------------------------------------------------------------------------------
constexpr int kScanBlock = 256;
constexpr int kScanItems = 8;                            // per thread, two int4
constexpr int kScanTile = kScanBlock * kScanItems;       // 2048 items per block
constexpr uint64_t kFlagAggregate = 1ull << 62;
constexpr uint64_t kFlagPrefix = 2ull << 62;
constexpr uint64_t kValueMask = (1ull << 62) - 1;

__global__ void __launch_bounds__(kScanBlock)
k_scan_exclusive(int64_t n, const int* __restrict__ in, int* __restrict__ out,
                 unsigned* __restrict__ ticket, volatile uint64_t* __restrict__ status,
                 int64_t* __restrict__ total_dev, int64_t* __restrict__ total_pinned,
                 const unsigned* __restrict__ aux_pair /* e.g. depth min/max, copied to pinned[1] */,
                 const uint64_t* __restrict__ gather /* optional: item i is in[(u32)gather[i]] */) {
    __shared__ unsigned s_block;
    __shared__ int64_t s_warp[kScanBlock / 32];
    __shared__ int64_t s_prefix;
    if (threadIdx.x == 0) s_block = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned bid = s_block;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t base = (int64_t)bid * kScanTile + (int64_t)threadIdx.x * kScanItems;

    int v[kScanItems];
    if (gather != nullptr) {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) v[k] = (base + k < n) ? in[(unsigned)gather[base + k]] : 0;
    } else if (base + kScanItems <= n) {
        const int4 a = *reinterpret_cast<const int4*>(in + base);
        const int4 b = *reinterpret_cast<const int4*>(in + base + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) v[k] = (base + k < n) ? in[base + k] : 0;
    }
    int64_t tsum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) tsum += v[k];

    // block-wide exclusive scan of the thread sums
    int64_t incl = tsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int64_t o = __shfl_up_sync(kFull, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int64_t warp_off = 0, block_sum = 0;
#pragma unroll
    for (int w = 0; w < kScanBlock / 32; ++w) {
        const int64_t s = s_warp[w];
        if (w < warp) warp_off += s;
        block_sum += s;
    }

    // decoupled look-back by warp 0: 32 predecessors per round trip (one per lane), consumed up to the
    // nearest inclusive prefix; a window with an unpublished status in front of that prefix is re-read
    if (warp == 0) {
        int64_t excl = 0;
        if (bid == 0) {
            if (lane == 0) status[0] = kFlagPrefix | (uint64_t)block_sum;
        } else {
            if (lane == 0) {
                status[bid] = kFlagAggregate | (uint64_t)block_sum;
                __threadfence();
            }
            int64_t j = (int64_t)bid - 1;
            while (true) {
                const int64_t idx = j - lane;
                uint64_t st = 2ull << 62;  // in front of block 0: an inclusive prefix of zero
                if (idx >= 0) st = status[idx];
                const unsigned pref = __ballot_sync(kFull, (st & kFlagPrefix) != 0);
                const unsigned none = __ballot_sync(kFull, (st & (kFlagPrefix | kFlagAggregate)) == 0);
                const int first = pref ? __ffs(pref) - 1 : 31;  // last lane that is consumed this round
                const unsigned need = (first == 31) ? kFull : ((2u << first) - 1);
                if (none & need) continue;
                int64_t val = (lane <= first) ? (int64_t)(st & kValueMask) : 0;
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) val += __shfl_xor_sync(kFull, val, d);
                excl += val;
                if (pref) break;
                j -= 32;
            }
            if (lane == 0) status[bid] = kFlagPrefix | (uint64_t)(excl + block_sum);
        }
        if (lane == 0) {
            s_prefix = excl;
            if ((int64_t)(bid + 1) * kScanTile >= n) {  // last logical block
                const int64_t total = excl + block_sum;
                if (total_dev) *total_dev = total;
                if (total_pinned) {
                    total_pinned[0] = total;
                    if (aux_pair) total_pinned[1] = (int64_t)((uint64_t)aux_pair[0] | ((uint64_t)aux_pair[1] << 32));
                }
            }
        }
    }
    __syncthreads();
    int64_t run = s_prefix + warp_off + (incl - tsum);
    int o[kScanItems];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) { o[k] = (int)run; run += v[k]; }
    if (base + kScanItems <= n) {
        *reinterpret_cast<int4*>(out + base) = make_int4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<int4*>(out + base + 4) = make_int4(o[4], o[5], o[6], o[7]);
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k)
            if (base + k < n) out[base + k] = o[k];
    }
}

// ------------------------------------------------------------------------------------------------
// duplicateWithKeys
// ------------------------------------------------------------------------------------------------
constexpr int kDupBlock = 256;

__global__ void __launch_bounds__(kDupBlock)
k_duplicate_with_keys(int64_t n, int width, int height, int ntx, int nty,
                      const float* __restrict__ means_2d, const float* __restrict__ depths,
                      const int* __restrict__ radii, const int* __restrict__ tiles_touched,
                      const int* __restrict__ offsets, int64_t p, uint64_t* __restrict__ keys,
                      int* __restrict__ values) {
    const int lane = threadIdx.x & 31;
    const int64_t g0 = ((int64_t)blockIdx.x * (kDupBlock / 32) + (threadIdx.x >> 5)) * 32;
    if (g0 >= n) return;
    const int64_t i = g0 + lane;

    int reserved = 0, emit = 0, tx0 = 0, ty0 = 0, w = 1;
    unsigned dbits = 0;
    int64_t off = 0;
    if (i < n) {
        reserved = tiles_touched[i];
        off = offsets[i];
        const int radius = radii[i];
        if (radius > 0) {  // sorting.cu:44-45
            const float2 m = reinterpret_cast<const float2*>(means_2d)[i];
            const TileRect r = tile_rect(m.x, m.y, radius, width, height, ntx, nty);
            const int ww = r.tx1 - r.tx0, hh = r.ty1 - r.ty0;
            if (ww > 0 && hh > 0) { emit = ww * hh; w = ww; }
            tx0 = r.tx0; ty0 = r.ty0;
            dbits = __float_as_uint(depths[i]);
        }
    }
    // exclusive prefix of reserved slots within the warp
    int incl = reserved;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(kFull, incl, d);
        if (lane >= d) incl += o;
    }
    const int wpre = incl - reserved;
    const int total = __shfl_sync(kFull, incl, 31);
    const int64_t base = __shfl_sync(kFull, off, 0);

    for (int k0 = 0; k0 < total; k0 += 32) {
        const int k = k0 + lane;
        int owner = 0;  // largest lane with wpre <= k
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1) {
            const int cand = owner + step;
            const int vpre = __shfl_sync(kFull, wpre, cand & 31);
            if (cand < 32 && vpre <= k) owner = cand;
        }
        const int j = k - __shfl_sync(kFull, wpre, owner);
        const int o_emit = __shfl_sync(kFull, emit, owner);
        const int o_w = __shfl_sync(kFull, w, owner);
        const int o_tx0 = __shfl_sync(kFull, tx0, owner);
        const int o_ty0 = __shfl_sync(kFull, ty0, owner);
        const unsigned o_db = __shfl_sync(kFull, dbits, owner);
        if (k < total && base + k < p) {
            uint64_t key = 0;  // A.2: reserved-but-not-emitted slots hold key 0 / value 0
            int val = 0;
            if (j < o_emit) {
                const int ty = o_ty0 + j / o_w, tx = o_tx0 + j % o_w;  // ty outer, tx inner (:63-64)
                key = ((uint64_t)(unsigned)(ty * ntx + tx) << 32) | (uint64_t)o_db;
                val = (int)(g0 + owner);
            }
            keys[base + k] = key;
            values[base + k] = val;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// tile ranges (sorting.cu:82-109)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_tile_ranges(int64_t p, const uint64_t* __restrict__ keys, int num_tiles, int* __restrict__ ranges) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p) return;
    const unsigned cur = (unsigned)(keys[i] >> 32);
    if (cur >= (unsigned)num_tiles) return;  // cannot happen for keys produced by this library
    if (i == 0) {
        ranges[cur * 2 + 0] = 0;
    } else {
        const unsigned prev = (unsigned)(keys[i - 1] >> 32);
        if (cur != prev) {
            if (prev < (unsigned)num_tiles) ranges[prev * 2 + 1] = (int)i;
            ranges[cur * 2 + 0] = (int)i;
        }
    }
    if (i == p - 1) ranges[cur * 2 + 1] = (int)p;
}

}  // namespace cugs

using namespace cugs;

extern "C" size_t cugs_b200_scan_temp_bytes(int64_t n) {
    const int64_t blocks = (n + kScanTile - 1) / kScanTile;
    return 16 + (size_t)(blocks > 0 ? blocks : 1) * sizeof(uint64_t);
}

int cugs_scan_launch(cugs_handle_t* h, cudaStream_t s, int64_t n, const int32_t* tiles_touched,
                     int32_t* offsets, int64_t* total_dev, int64_t* total_pinned, void* scan_temp,
                     const unsigned* aux_pair, const uint64_t* gather) {
    const int64_t blocks = (n + kScanTile - 1) / kScanTile;
    CUGS_CUDA_TRY(h, cudaMemsetAsync(scan_temp, 0, cugs_b200_scan_temp_bytes(n), s));
    unsigned* ticket = reinterpret_cast<unsigned*>(scan_temp);
    uint64_t* status = reinterpret_cast<uint64_t*>(reinterpret_cast<char*>(scan_temp) + 16);
    k_scan_exclusive<<<(unsigned)blocks, kScanBlock, 0, s>>>(n, tiles_touched, offsets, ticket, status,
                                                             total_dev, total_pinned, aux_pair, gather);
    CUGS_LAUNCH_CHECK(h, "k_scan_exclusive");
    return CUGS_OK;
}

extern "C" int cugs_b200_scan(cugs_handle_t* h, void* stream, int64_t n, const int32_t* tiles_touched,
                              int32_t* offsets, int64_t* total_dev, int64_t* total_host,
                              void* scan_temp, size_t scan_temp_bytes) {
    CUGS_REQUIRE(h, h != nullptr, "handle is null");
    CUGS_REQUIRE(h, n >= 0, "n must be >= 0");
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) {
        if (total_dev) CUGS_CUDA_TRY(h, cudaMemsetAsync(total_dev, 0, sizeof(int64_t), s));
        if (total_host) { CUGS_CUDA_TRY(h, cudaStreamSynchronize(s)); *total_host = 0; }
        return CUGS_OK;
    }
    CUGS_REQUIRE(h, tiles_touched && offsets && scan_temp, "null pointer");
    if (scan_temp_bytes < cugs_b200_scan_temp_bytes(n))
        return set_error(h, CUGS_ERR_WORKSPACE, "scan_temp too small: %zu < %zu", scan_temp_bytes,
                         cugs_b200_scan_temp_bytes(n));
    // every blocking call gets its own pinned word (a ring), so that a scan issued between the plan and
    // the finish of a frame, or by another stream, cannot overwrite a count somebody still has to read
    int64_t* slot = total_host ? cugs_pinned_slot(h) : nullptr;
    if (int e = cugs_scan_launch(h, s, n, tiles_touched, offsets, total_dev, slot, scan_temp, nullptr, nullptr))
        return e;
    if (total_host) {
        CUGS_CUDA_TRY(h, cudaStreamSynchronize(s));  // the one blocking read (sorting.cu:146)
        *total_host = *slot;
    }
    return CUGS_OK;
}

extern "C" int cugs_b200_duplicate_with_keys(cugs_handle_t* h, void* stream, int64_t n, int width,
                                             int height, const float* means_2d, const float* depths,
                                             const int32_t* radii, const int32_t* tiles_touched,
                                             const int32_t* offsets, int64_t p, uint64_t* keys,
                                             int32_t* values) {
    CUGS_REQUIRE(h, h != nullptr, "handle is null");
    CUGS_REQUIRE(h, n >= 0 && p >= 0, "n and p must be >= 0");
    CUGS_REQUIRE(h, width > 0 && height > 0, "image size must be positive");
    if (n == 0 || p == 0) return CUGS_OK;
    CUGS_REQUIRE(h, means_2d && depths && radii && tiles_touched && offsets && keys && values,
                 "null pointer");
    const int ntx = (width + kTile - 1) / kTile, nty = (height + kTile - 1) / kTile;
    const unsigned grid = (unsigned)((n + kDupBlock - 1) / kDupBlock);
    k_duplicate_with_keys<<<grid, kDupBlock, 0, (cudaStream_t)stream>>>(
        n, width, height, ntx, nty, means_2d, depths, radii, tiles_touched, offsets, p, keys, values);
    CUGS_LAUNCH_CHECK(h, "k_duplicate_with_keys");
    return CUGS_OK;
}

extern "C" int cugs_b200_tile_ranges(cugs_handle_t* h, void* stream, int64_t p,
                                     const uint64_t* keys_sorted, int num_tiles, int32_t* tile_ranges) {
    CUGS_REQUIRE(h, h != nullptr, "handle is null");
    CUGS_REQUIRE(h, p >= 0 && num_tiles >= 0, "p and num_tiles must be >= 0");
    if (num_tiles == 0) return CUGS_OK;
    CUGS_REQUIRE(h, tile_ranges != nullptr, "null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    CUGS_CUDA_TRY(h, cudaMemsetAsync(tile_ranges, 0, (size_t)num_tiles * 2 * sizeof(int), s));  // :216
    if (p == 0) return CUGS_OK;
    CUGS_REQUIRE(h, keys_sorted != nullptr, "null pointer");
    k_tile_ranges<<<(unsigned)((p + 255) / 256), 256, 0, s>>>(p, keys_sorted, num_tiles, tile_ranges);
    CUGS_LAUNCH_CHECK(h, "k_tile_ranges");
    return CUGS_OK;
}
