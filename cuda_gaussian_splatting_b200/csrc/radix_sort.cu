// radix_sort.cu — hand-written onesweep LSD radix sort of (u64 tile|depth key, u32 value) pairs.
//
// Replaces cub::DeviceRadixSort::SortPairs over all 64 key bits (reference
// rasterizer/sorting.cu:190-211; CUB Policy1000 = histogram + onesweep, 8 passes of 8 bits).
// Only the key bits that can differ are sorted: depth bits [0, depth_bits) and tile bits
// [32, 32 + tile_bits), viewed as one compact key k' = (tile << depth_bits) | depth. At 1080p
// (8160 tiles -> 13 bits) that is <= 45 bits = 6 passes instead of 8; with the depth range trim
// (highest differing depth bit from preprocess) typically 5.
//
// Structure per sort:
//   k_sort_histogram : one read of the keys builds the digit histograms of ALL passes
//   k_sort_scan_bins : exclusive scan of each pass's 256 bins -> global bin bases
//   k_onesweep (x passes): each block takes one tile of keys (ticketed), ranks them stably with
//       warp match + per-warp counters, resolves its global bin offsets with a decoupled
//       look-back over the preceding tiles' per-bin counts (single pass over the data: 12 B read
//       + 12 B written per pair and pass), and scatters through shared memory so that runs of
//       equal digits leave the SM as contiguous segments.
// A stable LSD sort over the differing bits gives exactly the permutation of the full 64-bit
// stable sort, so keys/values are bit-identical to the reference's CUB call.
#include "common.cuh"

namespace cugs {

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kMaxPasses = 8;

constexpr int kSortThreads = 384;
constexpr int kSortItems = 12;
constexpr int kSortTile = kSortThreads * kSortItems;  // 4608 pairs per block
constexpr int kSortWarps = kSortThreads / 32;

constexpr size_t kSortSmemBytes = (size_t)kSortTile * 8 + (size_t)kRadix * 8 +
                                  (size_t)kSortWarps * kRadix * 4 + (size_t)kRadix * 4 + 64;

constexpr unsigned kLbAggregate = 1u << 30;
constexpr unsigned kLbPrefix = 2u << 30;
constexpr unsigned kLbValue = (1u << 30) - 1;

struct SortPlan {
    int passes;
    int depth_bits;          // db: low key bits kept
    int shift[kMaxPasses];   // bit offset in the compact key
    int bits[kMaxPasses];    // digit width of the pass (<= 8)
};

__host__ __device__ __forceinline__ uint64_t compact_key(uint64_t key, int db) {
    const uint64_t lo = (db >= 32) ? (key & 0xffffffffull) : (key & ((1ull << db) - 1));
    return ((key >> 32) << db) | lo;
}
__device__ __forceinline__ unsigned digit_of(uint64_t key, int db, int shift, unsigned mask) {
    return (unsigned)(compact_key(key, db) >> shift) & mask;
}

// ------------------------------------------------------------------------------------------------
// histogram of every pass in one read of the keys (8 B/pair)
// ------------------------------------------------------------------------------------------------
constexpr int kHistThreads = 512;
constexpr int kHistItems = 8;

__global__ void __launch_bounds__(kHistThreads)
k_sort_histogram(int64_t p, const uint64_t* __restrict__ keys, SortPlan plan, unsigned* __restrict__ hist) {
    __shared__ unsigned sh[kMaxPasses * kRadix];
    for (int b = threadIdx.x; b < plan.passes * kRadix; b += kHistThreads) sh[b] = 0;
    __syncthreads();
    const int64_t tile = (int64_t)kHistThreads * kHistItems;
    for (int64_t base = (int64_t)blockIdx.x * tile; base < p; base += (int64_t)gridDim.x * tile) {
        uint64_t k[kHistItems];
#pragma unroll
        for (int i = 0; i < kHistItems; ++i) {
            const int64_t idx = base + (int64_t)i * kHistThreads + threadIdx.x;
            k[i] = (idx < p) ? __ldcs(keys + idx) : 0ull;
        }
#pragma unroll
        for (int i = 0; i < kHistItems; ++i) {
            const int64_t idx = base + (int64_t)i * kHistThreads + threadIdx.x;
            if (idx < p) {
                const uint64_t ck = compact_key(k[i], plan.depth_bits);
                for (int ps = 0; ps < plan.passes; ++ps) {
                    const unsigned d = (unsigned)(ck >> plan.shift[ps]) & ((1u << plan.bits[ps]) - 1);
                    atomicAdd(&sh[ps * kRadix + d], 1u);
                }
            }
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < plan.passes * kRadix; b += kHistThreads) {
        const unsigned c = sh[b];
        if (c) atomicAdd(&hist[b], c);
    }
}

// exclusive scan of each pass's bins, in place: hist[pass][bin] -> first global index of the bin
__global__ void __launch_bounds__(kRadix) k_sort_scan_bins(unsigned* __restrict__ hist) {
    __shared__ unsigned swarp[kRadix / 32];
    unsigned* hp = hist + blockIdx.x * kRadix;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned c = hp[threadIdx.x];
    unsigned incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned o = __shfl_up_sync(kFull, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) swarp[warp] = incl;
    __syncthreads();
    unsigned off = 0;
    for (int w = 0; w < warp; ++w) off += swarp[w];
    hp[threadIdx.x] = off + incl - c;
}

// ------------------------------------------------------------------------------------------------
// one onesweep pass
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSortThreads)
k_onesweep(int64_t p, const uint64_t* __restrict__ keys_in, const int* __restrict__ vals_in,
           uint64_t* __restrict__ keys_out, int* __restrict__ vals_out,
           const unsigned* __restrict__ bin_base, volatile unsigned* __restrict__ lookback,
           unsigned* __restrict__ ticket, int db, int shift, int bits) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    uint64_t* s_keys = reinterpret_cast<uint64_t*>(s_raw);                   // [kSortTile], reused for values
    int64_t* s_bin_global = reinterpret_cast<int64_t*>(s_keys + kSortTile);   // global index = [d] + slot
    unsigned(*s_warp_hist)[kRadix] =                                          // counts, then per-warp offsets
        reinterpret_cast<unsigned(*)[kRadix]>(s_bin_global + kRadix);
    unsigned* s_bin_start = &s_warp_hist[kSortWarps][0];                      // local exclusive scan of bin counts
    unsigned* s_scan = s_bin_start + kRadix;
    __shared__ unsigned s_tile;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    for (int b = lane; b < kRadix; b += 32) s_warp_hist[warp][b] = 0;
    __syncthreads();
    const unsigned tile = s_tile;
    const int64_t tile_base = (int64_t)tile * kSortTile;
    const int valid = (int)min((int64_t)kSortTile, p - tile_base);
    const int64_t seg = tile_base + (int64_t)warp * (32 * kSortItems);
    const unsigned mask = (1u << bits) - 1;
    const unsigned lt_mask = (1u << lane) - 1;

    // ---- load keys (warp-striped: consecutive lanes, consecutive keys) ----
    uint64_t key[kSortItems];
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const int64_t idx = seg + i * 32 + lane;
        key[i] = (idx < p) ? __ldcs(keys_in + idx) : ~0ull;
    }

    // ---- stable ranking inside the warp: match peers with the same digit, per-warp counters ----
    unsigned short rank[kSortItems];
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const unsigned d = digit_of(key[i], db, shift, mask);
        const unsigned peers = match_digit(d, bits);
        const int leader = __ffs(peers) - 1;
        unsigned old = 0;
        if (lane == leader) {
            old = s_warp_hist[warp][d];
            s_warp_hist[warp][d] = old + __popc(peers);
        }
        old = __shfl_sync(kFull, old, leader);
        rank[i] = (unsigned short)(old + __popc(peers & lt_mask));
        __syncwarp();
    }
    __syncthreads();

    // ---- per-bin: scan over warps, block count, local start, global base via look-back ----
    unsigned bin_count = 0;
    if (tid < kRadix) {
        unsigned run = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            const unsigned c = s_warp_hist[w][tid];
            s_warp_hist[w][tid] = run;
            run += c;
        }
        bin_count = run;
        unsigned incl = run;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned o = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) s_scan[warp] = incl;
        // keep incl for after the barrier
        bin_count = run;
        s_bin_start[tid] = incl - run;  // provisional: exclusive within the warp
    }
    __syncthreads();
    if (tid < kRadix) {
        unsigned off = 0;
        for (int w = 0; w < warp; ++w) off += s_scan[w];
        const unsigned local_start = s_bin_start[tid] + off;

        volatile unsigned* lb = lookback + (size_t)tile * kRadix;
        unsigned excl = 0;
        if (tile == 0) {
            lb[tid] = kLbPrefix | bin_count;
        } else {
            lb[tid] = kLbAggregate | bin_count;
            int64_t j = (int64_t)tile - 1;
            while (true) {
                const unsigned s = lookback[(size_t)j * kRadix + tid];
                if (s & kLbPrefix) { excl += s & kLbValue; break; }
                if (s & kLbAggregate) { excl += s & kLbValue; --j; }
            }
            lb[tid] = kLbPrefix | ((excl + bin_count) & kLbValue);
        }
        s_bin_global[tid] = (int64_t)bin_base[tid] + (int64_t)excl - (int64_t)local_start;
        __syncwarp();
        s_bin_start[tid] = local_start;
    }
    __syncthreads();

    // ---- scatter keys into their local sorted slot ----
    unsigned short pos[kSortItems];
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const unsigned d = digit_of(key[i], db, shift, mask);
        pos[i] = (unsigned short)(s_bin_start[d] + s_warp_hist[warp][d] + rank[i]);
        s_keys[pos[i]] = key[i];
    }
    __syncthreads();

    // ---- write keys: slot -> global index (runs of equal digits are contiguous) ----
    int64_t gpos[kSortItems];
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const int slot = tid + i * kSortThreads;
        const uint64_t k = s_keys[slot];
        const unsigned d = digit_of(k, db, shift, mask);
        gpos[i] = s_bin_global[d] + slot;
        if (slot < valid) keys_out[gpos[i]] = k;
    }

    // ---- values follow the same permutation ----
    int val[kSortItems];
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const int64_t idx = seg + i * 32 + lane;
        val[i] = (idx < p) ? __ldcs(vals_in + idx) : 0;
    }
    __syncthreads();
    int* s_vals = reinterpret_cast<int*>(s_keys);
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) s_vals[pos[i]] = val[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const int slot = tid + i * kSortThreads;
        if (slot < valid) vals_out[gpos[i]] = s_vals[slot];
    }
}

inline SortPlan make_sort_plan(int depth_bits, int tile_bits) {
    SortPlan pl{};
    pl.depth_bits = depth_bits;
    const int total = depth_bits + tile_bits;
    pl.passes = (total + kRadixBits - 1) / kRadixBits;
    if (pl.passes < 1) pl.passes = 1;
    int shift = 0;
    for (int i = 0; i < pl.passes; ++i) {  // spread the bits evenly: e.g. 45 -> 8,8,8,7,7,7
        const int left = total - shift, passes_left = pl.passes - i;
        int b = (left + passes_left - 1) / passes_left;
        if (b < 1) b = 1;
        pl.shift[i] = shift;
        pl.bits[i] = b;
        shift += b;
    }
    return pl;
}

}  // namespace cugs

using namespace cugs;

extern "C" int cugs_b200_sort_num_passes(int depth_bits, int tile_bits) {
    if (depth_bits < 0 || depth_bits > 32 || tile_bits < 0 || tile_bits > 32) return -1;
    return make_sort_plan(depth_bits, tile_bits).passes;
}

extern "C" size_t cugs_b200_sort_temp_bytes(int64_t p) {
    const int64_t tiles = (p + kSortTile - 1) / kSortTile;
    return (size_t)kMaxPasses * kRadix * 4 + 64 +
           (size_t)kMaxPasses * (size_t)(tiles > 0 ? tiles : 1) * kRadix * 4;
}

// Runs the passes src -> dst -> src ...; the result is in (keys_b, vals_b) if the number of passes
// is odd and in (keys_a, vals_a) if it is even; *result_in_b tells which.
extern "C" int cugs_b200_sort_pairs_pingpong(cugs_handle_t* h, void* stream, int64_t p, int depth_bits,
                                             int tile_bits, uint64_t* keys_a, int32_t* vals_a,
                                             uint64_t* keys_b, int32_t* vals_b, void* temp,
                                             size_t temp_bytes, int* result_in_b) {
    CUGS_REQUIRE(h, h != nullptr, "handle is null");
    CUGS_REQUIRE(h, p >= 0, "p must be >= 0");
    CUGS_REQUIRE(h, depth_bits >= 0 && depth_bits <= 32 && tile_bits >= 0 && tile_bits <= 32,
                 "depth_bits / tile_bits must be in 0..32");
    const SortPlan plan = make_sort_plan(depth_bits, tile_bits);
    if (result_in_b) *result_in_b = plan.passes & 1;
    if (p == 0) return CUGS_OK;
    if (p >= (1ll << 30))
        return set_error(h, CUGS_ERR_UNSUPPORTED, "P = %lld >= 2^30 pairs is not supported", (long long)p);
    CUGS_REQUIRE(h, keys_a && vals_a && keys_b && vals_b && temp, "null pointer");
    if (temp_bytes < cugs_b200_sort_temp_bytes(p))
        return set_error(h, CUGS_ERR_WORKSPACE, "sort temp too small: %zu < %zu", temp_bytes,
                         cugs_b200_sort_temp_bytes(p));
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t tiles = (p + kSortTile - 1) / kSortTile;
    unsigned* hist = reinterpret_cast<unsigned*>(temp);
    unsigned* tickets = hist + kMaxPasses * kRadix;
    unsigned* lookback = tickets + 16;
    const size_t used = (size_t)kMaxPasses * kRadix * 4 + 64 + (size_t)plan.passes * tiles * kRadix * 4;
    CUGS_CUDA_TRY(h, cudaMemsetAsync(temp, 0, used, s));

    int hist_blocks = h->sm_count * 2;
    const int64_t hist_tile = (int64_t)kHistThreads * kHistItems;
    if ((int64_t)hist_blocks * hist_tile > p) hist_blocks = (int)((p + hist_tile - 1) / hist_tile);
    k_sort_histogram<<<hist_blocks, kHistThreads, 0, s>>>(p, keys_a, plan, hist);
    CUGS_LAUNCH_CHECK(h, "k_sort_histogram");
    k_sort_scan_bins<<<plan.passes, kRadix, 0, s>>>(hist);
    CUGS_LAUNCH_CHECK(h, "k_sort_scan_bins");

    CUGS_CUDA_TRY(h, cudaFuncSetAttribute(k_onesweep, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)kSortSmemBytes));
    uint64_t* ksrc = keys_a; int32_t* vsrc = vals_a;
    uint64_t* kdst = keys_b; int32_t* vdst = vals_b;
    for (int ps = 0; ps < plan.passes; ++ps) {
        k_onesweep<<<(unsigned)tiles, kSortThreads, kSortSmemBytes, s>>>(
            p, ksrc, vsrc, kdst, vdst, hist + ps * kRadix, lookback + (size_t)ps * tiles * kRadix,
            tickets + ps, plan.depth_bits, plan.shift[ps], plan.bits[ps]);
        CUGS_LAUNCH_CHECK(h, "k_onesweep");
        uint64_t* tk = ksrc; ksrc = kdst; kdst = tk;
        int32_t* tv = vsrc; vsrc = vdst; vdst = tv;
    }
    return CUGS_OK;
}

extern "C" int cugs_b200_sort_pairs(cugs_handle_t* h, void* stream, int64_t p, int depth_bits,
                                    int tile_bits, uint64_t* keys_in, int32_t* values_in,
                                    uint64_t* keys_out, int32_t* values_out, void* temp,
                                    size_t temp_bytes) {
    CUGS_REQUIRE(h, h != nullptr, "handle is null");
    CUGS_REQUIRE(h, depth_bits >= 0 && depth_bits <= 32 && tile_bits >= 0 && tile_bits <= 32,
                 "depth_bits / tile_bits must be in 0..32");
    if (p <= 0) return p == 0 ? CUGS_OK : set_error(h, CUGS_ERR_INVALID_ARG, "p must be >= 0");
    const int passes = make_sort_plan(depth_bits, tile_bits).passes;
    cudaStream_t s = (cudaStream_t)stream;
    int in_b = 0;
    if (passes & 1)
        return cugs_b200_sort_pairs_pingpong(h, stream, p, depth_bits, tile_bits, keys_in, values_in,
                                             keys_out, values_out, temp, temp_bytes, &in_b);
    // even number of passes: start from the output buffers so that the result lands there
    CUGS_REQUIRE(h, keys_in && values_in && keys_out && values_out, "null pointer");
    CUGS_CUDA_TRY(h, cudaMemcpyAsync(keys_out, keys_in, (size_t)p * 8, cudaMemcpyDeviceToDevice, s));
    CUGS_CUDA_TRY(h, cudaMemcpyAsync(values_out, values_in, (size_t)p * 4, cudaMemcpyDeviceToDevice, s));
    return cugs_b200_sort_pairs_pingpong(h, stream, p, depth_bits, tile_bits, keys_out, values_out,
                                         keys_in, values_in, temp, temp_bytes, &in_b);
}
