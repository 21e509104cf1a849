// tile_binning.cu — the binning pipeline of the fused render path (render_plan / render_finish).
//
// Produces exactly what the reference's sort_gaussians produces (rasterizer/sorting.cu:115-227:
// stable ascending sort of the 64-bit keys tile<<32 | depth_bits with CUB over all 64 bits, then
// k_compute_tile_ranges): the sorted Gaussian indices and the tile ranges, bit for bit — but as an
// LSD radix sort whose passes over the DEPTH bits are hoisted in front of duplicateWithKeys:
// every (tile, Gaussian) pair of one Gaussian carries the same depth digits, so the low 32 key
// bits are sorted once per Gaussian (N elements) instead of once per pair (P ~ 6 N elements), and
// only the tile bits are sorted on the pairs:
//
//   1. preprocess writes one packed element per Gaussian: depth_key << 32 | index, depth_key =
//      float bits of the depth (0 for the filler entries of quirk A.2, 0xffffffff if the Gaussian
//      emits no pair)                                                       [preprocess.cu]
//   2. stable LSD sort of the N packed elements by depth_key              (4 onesweep passes x 8 bits)
//   3. exclusive scan of tiles_touched in depth order -> pair offsets, P  [binning.cu, gather-scan]
//   4. duplicateWithKeys in depth order: packed pair = tile_id << 32 | gaussian index
//   5. tile histogram (-> tile ranges by an exclusive scan, no pass over sorted keys) and stable
//      LSD sort of the P pairs by tile id                                  (ceil(tile_bits/8) passes)
//
// Order of the result: tile ascending; inside a tile the emission order of step 4 = (depth bits
// ascending, Gaussian index ascending) — the order of the reference's stable 64-bit sort.
// HBM traffic: 8 B/pair per pass instead of 12, 2 passes over the pairs instead of 6 (1080p).
#include "common.cuh"

namespace cugs {

constexpr int kPkRadixBits = 8;
constexpr int kPkRadix = 256;
constexpr int kPkThreads = 512;
constexpr int kPkItems = 8;
// (kPkThreads * kPkItems = 4096 elements per block)
// Experiment (kept as a build option, off by default): short inputs -- the depth sort of the N Gaussians, 3 M
// elements = 733 tiles of 4096 on 444 block slots -- are latency bound by a block's load -> rank -> look-back ->
// scatter chain, so a 4-items-per-thread variant (2048-element tiles, twice as many, shorter blocks) was tried.
// Measured on B200 (profiles/r02/NOTES.md): SLOWER, 35.1 us instead of 29 us per depth pass at 3 M (the look-back
// chain doubles in length and the per-block fixed costs are paid twice); no difference at 100 k.
constexpr int kPkItemsSmall = 4;
#ifndef CUGS_PK_SMALL_LIMIT
#define CUGS_PK_SMALL_LIMIT 0  // elements up to which the small-tile variant is used; 0 = never
#endif
// Long inputs (the P pairs): 16 items per thread (8192-element tiles, two blocks per SM) halve the number of tiles
// and with it the look-back polling per element, which is a third of the instructions of a pass at 8 items.
// Measured on B200 at P = 18.6 M (profiles/r02/bench_variant_large_tiles.json): sort stage 0.381 -> 0.358 ms.
constexpr int kPkItemsLarge = 16;
#ifndef CUGS_PK_LARGE_LIMIT
#define CUGS_PK_LARGE_LIMIT (8ll << 20)  // elements (launch capacity) from which the large-tile variant is used
#endif
inline int pk_items_for(int64_t n) {
    if (n > 0 && n <= (int64_t)CUGS_PK_SMALL_LIMIT) return kPkItemsSmall;
    if (n >= (int64_t)CUGS_PK_LARGE_LIMIT) return kPkItemsLarge;
    return kPkItems;
}
constexpr int kPkWarps = kPkThreads / 32;
constexpr int kPkMaxPasses = 4;

constexpr unsigned kPkAggregate = 1u << 30;
constexpr unsigned kPkPrefix = 2u << 30;
constexpr unsigned kPkValue = (1u << 30) - 1;

struct PackedPlan {
    int passes;
    int shift[kPkMaxPasses];  // bit offset inside the HIGH 32 bits of the element
    int bits[kPkMaxPasses];
};

inline PackedPlan make_packed_plan(int key_bits) {
    PackedPlan pl{};
    pl.passes = (key_bits + kPkRadixBits - 1) / kPkRadixBits;
    if (pl.passes < 1) pl.passes = 1;
    if (pl.passes > kPkMaxPasses) pl.passes = kPkMaxPasses;
    int shift = 0;
    for (int i = 0; i < pl.passes; ++i) {  // spread the bits evenly: 13 -> 7,6
        const int left = key_bits - shift, passes_left = pl.passes - i;
        int b = (left + passes_left - 1) / passes_left;
        if (b < 1) b = 1;
        if (b > kPkRadixBits) b = kPkRadixBits;
        pl.shift[i] = shift;
        pl.bits[i] = b;
        shift += b;
    }
    return pl;
}

__device__ __forceinline__ unsigned pk_digit(uint64_t e, int shift, unsigned mask) {
    return ((unsigned)(e >> 32) >> shift) & mask;
}

// ------------------------------------------------------------------------------------------------
// histograms: digit histograms of every pass (+ optionally the full per-tile histogram) in ONE read
// ------------------------------------------------------------------------------------------------
constexpr int kPkHistThreads = 1024;  // one fat block per SM: the flush of the tile histogram (one global atomic per
                                      // block and non-empty tile) is what bounds the pair histogram, not its body
constexpr int kPkHistItems = 8;

template <bool kTileHist>
__global__ void __launch_bounds__(kPkHistThreads)
k_packed_histogram(int64_t n_cap, const int64_t* __restrict__ n_dev, const uint64_t* __restrict__ elts, PackedPlan plan,
                   unsigned* __restrict__ digit_hist /* [passes][256] */, int num_tiles,
                   unsigned* __restrict__ tile_hist /* [num_tiles] */) {
    extern __shared__ unsigned s_hist[];  // [passes*256] (+ [num_tiles])
    // element count: host value, or (capacity-sized launch, no host round trip) the device word
    // written by the scan, clamped to the capacity of the buffers
    const int64_t n = n_dev ? min(*n_dev, n_cap) : n_cap;
    unsigned* s_tile = s_hist + plan.passes * kPkRadix;
    const int total_bins = plan.passes * kPkRadix + (kTileHist ? num_tiles : 0);
    for (int b = threadIdx.x; b < total_bins; b += kPkHistThreads) s_hist[b] = 0;
    __syncthreads();
    const int64_t tile = (int64_t)kPkHistThreads * kPkHistItems;
    for (int64_t base = (int64_t)blockIdx.x * tile; base < n; base += (int64_t)gridDim.x * tile) {
        uint64_t e[kPkHistItems];
#pragma unroll
        for (int i = 0; i < kPkHistItems; ++i) {
            const int64_t idx = base + (int64_t)i * kPkHistThreads + threadIdx.x;
            e[i] = (idx < n) ? __ldcs(elts + idx) : 0ull;
        }
#pragma unroll
        for (int i = 0; i < kPkHistItems; ++i) {
            const int64_t idx = base + (int64_t)i * kPkHistThreads + threadIdx.x;
            if (idx < n) {
                const unsigned key = (unsigned)(e[i] >> 32);
                if (kTileHist && key < (unsigned)num_tiles) {
                    // one atomic per pair: the digit histograms are folded out of the tile histogram below
                    atomicAdd(&s_tile[key], 1u);
                } else {
                    for (int ps = 0; ps < plan.passes; ++ps)
                        atomicAdd(&s_hist[ps * kPkRadix + ((key >> plan.shift[ps]) & ((1u << plan.bits[ps]) - 1))], 1u);
                }
            }
        }
    }
    __syncthreads();
    if (kTileHist) {  // every digit of a key is a function of its tile id
        for (int b = threadIdx.x; b < num_tiles; b += kPkHistThreads) {
            const unsigned c = s_tile[b];
            if (c)
                for (int ps = 0; ps < plan.passes; ++ps)
                    atomicAdd(&s_hist[ps * kPkRadix + (((unsigned)b >> plan.shift[ps]) & ((1u << plan.bits[ps]) - 1))], c);
        }
        __syncthreads();
    }
    for (int b = threadIdx.x; b < plan.passes * kPkRadix; b += kPkHistThreads) {
        const unsigned c = s_hist[b];
        if (c) atomicAdd(&digit_hist[b], c);
    }
    if (kTileHist)
        for (int b = threadIdx.x; b < num_tiles; b += kPkHistThreads) {
            const unsigned c = s_tile[b];
            if (c) atomicAdd(&tile_hist[b], c);
        }
}

// exclusive scan of each pass's 256 bins, in place
__global__ void __launch_bounds__(kPkRadix) k_packed_scan_bins(unsigned* __restrict__ hist) {
    __shared__ unsigned swarp[kPkRadix / 32];
    unsigned* hp = hist + blockIdx.x * kPkRadix;
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

// tile histogram -> tile ranges [start, end) (rasterizer/sorting.cu:82-109 semantics: tiles
// without pairs stay {0, 0}). One block; num_tiles is at most a few ten thousand.
__global__ void __launch_bounds__(1024)
k_tile_hist_to_ranges(int num_tiles, const unsigned* __restrict__ tile_hist, int* __restrict__ ranges) {
    // one pass: every thread owns `per` consecutive tiles (8 at 1080p), so the block scans once
    __shared__ unsigned s_warp[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per = (num_tiles + 1023) / 1024;
    const int t0 = threadIdx.x * per, t1 = min(num_tiles, t0 + per);
    unsigned sum = 0;
    for (int t = t0; t < t1; ++t) sum += tile_hist[t];
    unsigned incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned o = __shfl_up_sync(kFull, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned start = incl - sum;
    for (int w = 0; w < warp; ++w) start += s_warp[w];
    for (int t = t0; t < t1; ++t) {
        const unsigned c = tile_hist[t];
        reinterpret_cast<int2*>(ranges)[t] = c ? make_int2((int)start, (int)(start + c)) : make_int2(0, 0);
        start += c;
    }
}

// ------------------------------------------------------------------------------------------------
// one onesweep pass over packed 64-bit elements (digit taken from the high word). kLast: write
// only the low word (the Gaussian index) of each element to out32.
// ------------------------------------------------------------------------------------------------
constexpr size_t pk_smem_bytes(int items) {
    return (size_t)kPkThreads * items * 8 + (size_t)kPkRadix * 8 + (size_t)kPkWarps * kPkRadix * 4 +
           (size_t)kPkRadix * 4 + 64;
}

#ifndef CUGS_OS_MINBLOCKS
#define CUGS_OS_MINBLOCKS 3  // 40 registers (12 B of spills): 3 x 512 threads per SM, sort stage 0.531 -> 0.510 ms (4 blocks: 0.544)
#endif
// kBits = digit width of the pass (6, 7, 8), 0 = run-time width; kItems = elements per thread (8 or 4)
template <bool kLast, int kBits, int kItems>
__global__ void __launch_bounds__(kPkThreads, (kItems > 8 ? 2 : CUGS_OS_MINBLOCKS))
k_onesweep_packed(int64_t n_cap, const int64_t* __restrict__ n_dev, const uint64_t* __restrict__ in,
                  uint64_t* __restrict__ out, int* __restrict__ out32, const unsigned* __restrict__ bin_base,
                  volatile unsigned* __restrict__ lookback, unsigned* __restrict__ ticket, int shift, int bits,
                  const int* __restrict__ payload_src, int* __restrict__ payload_dst) {
    // capacity-sized launch: exactly ceil(n / kTileElts) blocks pass this test and take tickets
    // 0 .. tiles-1, so the look-back chain is the same as with a grid sized on the host
    constexpr int kTileElts = kPkThreads * kItems;
    const int64_t n = n_dev ? min(*n_dev, n_cap) : n_cap;
    if ((int64_t)blockIdx.x * kTileElts >= n) return;
    extern __shared__ __align__(16) unsigned char s_raw[];
    uint64_t* s_elts = reinterpret_cast<uint64_t*>(s_raw);                   // [kTileElts]
    int64_t* s_bin_global = reinterpret_cast<int64_t*>(s_elts + kTileElts);     // global index = [d] + slot
    unsigned(*s_warp_hist)[kPkRadix] = reinterpret_cast<unsigned(*)[kPkRadix]>(s_bin_global + kPkRadix);
    unsigned* s_bin_start = &s_warp_hist[kPkWarps][0];
    unsigned* s_scan = s_bin_start + kPkRadix;
    __shared__ unsigned s_tile;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    for (int b = lane; b < kPkRadix; b += 32) s_warp_hist[warp][b] = 0;
    __syncthreads();
    const unsigned tile = s_tile;
    const int64_t tile_base = (int64_t)tile * kTileElts;
    const int valid = (int)min((int64_t)kTileElts, n - tile_base);
    const int64_t seg = tile_base + (int64_t)warp * (32 * kItems);
    const unsigned mask = (1u << bits) - 1;
    const unsigned lt_mask = (1u << lane) - 1;

    uint64_t e[kItems];
    if (tile_base + kTileElts <= n) {  // full tile: no bounds checks
#pragma unroll
        for (int i = 0; i < kItems; ++i) e[i] = __ldcs(in + seg + i * 32 + lane);
    } else {
#pragma unroll
        for (int i = 0; i < kItems; ++i) {
            const int64_t idx = seg + i * 32 + lane;
            e[i] = (idx < n) ? __ldcs(in + idx) : ~0ull;
        }
    }

    // Stable ranking inside the warp: lanes holding the same digit are matched with a ballot per bit;
    // the group's leader bumps the warp's counter with ONE shared-memory atomic whose return value is
    // the number of equal digits in the warp's earlier items (shared atomics of one warp retire in
    // program order), so the eight items are independent instruction streams for the scheduler.
    unsigned short rank[kItems];
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        const unsigned d = pk_digit(e[i], shift, mask);
        const unsigned peers = kBits > 0 ? match_digit_fixed<(kBits > 0 ? kBits : 1)>(d) : match_digit(d, bits);
        const int leader = __ffs(peers) - 1;
        unsigned old = 0;
        if (lane == leader) old = atomicAdd(&s_warp_hist[warp][d], (unsigned)__popc(peers));
        old = __shfl_sync(kFull, old, leader);
        rank[i] = (unsigned short)(old + __popc(peers & lt_mask));
    }
    __syncthreads();

    // per-bin: scan over warps -> block count; publish the tile's aggregate as early as possible
    unsigned bin_count = 0, incl = 0;
    volatile unsigned* lb = lookback + (size_t)tile * kPkRadix;
    if (tid < kPkRadix) {
        unsigned run = 0;
#pragma unroll
        for (int w = 0; w < kPkWarps; ++w) {
            const unsigned c = s_warp_hist[w][tid];
            s_warp_hist[w][tid] = run;
            run += c;
        }
        bin_count = run;
        lb[tid] = (tile == 0 ? kPkPrefix : kPkAggregate) | bin_count;
        incl = run;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned o = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) s_scan[warp] = incl;
    }
    __syncthreads();
    unsigned local_start = 0;
    if (tid < kPkRadix) {
        unsigned off = 0;
        for (int w = 0; w < warp; ++w) off += s_scan[w];
        local_start = incl - bin_count + off;
        s_bin_start[tid] = local_start;
    }
    __syncthreads();

    // scatter into the local sorted slot (needs only the local starts) ...
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        const unsigned d = pk_digit(e[i], shift, mask);
        s_elts[s_bin_start[d] + s_warp_hist[warp][d] + rank[i]] = e[i];
    }
    // ... then resolve the global bin bases with the decoupled look-back (by now the predecessors
    // have had the whole ranking + scatter time to publish)
    if (tid < kPkRadix) {
        unsigned excl = 0;
        if (tile != 0) {
            // eight predecessors per round trip (independent loads in flight), consumed in order up to
            // the first inclusive prefix or the first status that is not published yet
            int64_t j = (int64_t)tile - 1;
            bool done = false;
            while (!done) {
                unsigned st[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    unsigned v = 2u << 30;  // before tile 0: an inclusive prefix of zero
                    if (j - k >= 0) v = lookback[(size_t)(j - k) * kPkRadix + tid];
                    st[k] = v;
                }
                // branch-free consume: the flags are the two top bits, so "published" is v >= 1<<30 and
                // "inclusive prefix" is v >= 1<<31. Take entries up to and including the first prefix, or
                // up to (excluding) the first one that is not published yet.
                unsigned pubm = 0, prem = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    pubm |= (st[k] >= (1u << 30)) ? (1u << k) : 0u;
                    prem |= (st[k] >= (2u << 30)) ? (1u << k) : 0u;
                }
                const int first_unpub = __ffs(~pubm) - 1;            // 8 when all eight are published
                const int first_pref = prem ? __ffs(prem) - 1 : 8;
                done = first_pref < first_unpub;
                const int take = done ? first_pref + 1 : min(first_unpub, 8);
#pragma unroll
                for (int k = 0; k < 8; ++k) excl += (k < take) ? (st[k] & kPkValue) : 0u;
                j -= take;
            }
            lb[tid] = kPkPrefix | ((excl + bin_count) & kPkValue);
        }
        s_bin_global[tid] = (int64_t)bin_base[tid] + (int64_t)excl - (int64_t)local_start;
    }
    __syncthreads();

    // write runs of equal digits contiguously
    if (valid == kTileElts) {
#pragma unroll
        for (int i = 0; i < kItems; ++i) {
            const int slot = tid + i * kPkThreads;
            const uint64_t x = s_elts[slot];
            const int64_t g = s_bin_global[pk_digit(x, shift, mask)] + slot;
            if (kLast) out32[g] = (int)(unsigned)x;
            else out[g] = x;
            // optional: a per-element payload (indexed by the low word) gathered into sorted order by the
            // pass that knows the final position -- the depth sort's last pass delivers tiles_touched in depth
            // order, so the scan that follows reads contiguous memory
            if (payload_dst != nullptr) payload_dst[g] = payload_src[(unsigned)x];
        }
    } else {
#pragma unroll
        for (int i = 0; i < kItems; ++i) {
            const int slot = tid + i * kPkThreads;
            if (slot < valid) {
                const uint64_t x = s_elts[slot];
                const int64_t g = s_bin_global[pk_digit(x, shift, mask)] + slot;
                if (kLast) out32[g] = (int)(unsigned)x;
                else out[g] = x;
                if (payload_dst != nullptr) payload_dst[g] = payload_src[(unsigned)x];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// duplicateWithKeys in depth order (reference: k_fill_sort_pairs, rasterizer/sorting.cu:30-72).
// A warp owns 32 consecutive positions of the depth-sorted Gaussian list; their output slots are
// contiguous, so slot k is resolved to its owner lane by a shuffle binary search and every store
// of the warp is coalesced, whatever the splat size. Pair = tile_id << 32 | gaussian index; the
// reserved-but-not-emitted slots of quirk A.2 become (tile 0, Gaussian 0) = 0, as in the reference
// (whose zero-filled key 0 / value 0 sorts to the front of tile 0).
// ------------------------------------------------------------------------------------------------
constexpr int kDupBlock = 256;

// one warp: 32 consecutive positions of the depth-sorted Gaussian list starting at s0. s_tile (optional):
// the block's shared tile histogram, one shared atomic per emitted pair (fused pair histogram).
template <bool kHist>
__device__ __forceinline__ void dup_warp(int64_t s0, int64_t n, int width, int height, int ntx, int nty,
                                         const uint64_t* __restrict__ sorted_elts, const float* __restrict__ means_2d,
                                         const int* __restrict__ radii,
                                         const int* __restrict__ tiles_sorted /* tiles_touched in depth order */,
                                         const int* __restrict__ offsets, int64_t p, uint64_t* __restrict__ pairs,
                                         unsigned* __restrict__ s_tile) {
    const int lane = threadIdx.x & 31;
    const int64_t si = s0 + lane;

    int reserved = 0, emit = 0, tx0 = 0, ty0 = 0, w = 1;
    unsigned g = 0;
    int64_t off = 0;
    if (si < n) {
        g = (unsigned)sorted_elts[si];
        reserved = tiles_sorted[si];
        off = offsets[si];
        const int radius = radii[g];
        if (radius > 0 && reserved > 0) {  // sorting.cu:44-45
            const float2 m = reinterpret_cast<const float2*>(means_2d)[g];
            const TileRect r = tile_rect(m.x, m.y, radius, width, height, ntx, nty);
            const int ww = r.tx1 - r.tx0, hh = r.ty1 - r.ty0;
            if (ww > 0 && hh > 0) { emit = ww * hh; w = ww; }
            tx0 = r.tx0; ty0 = r.ty0;
        }
    }
    int incl = reserved;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(kFull, incl, d);
        if (lane >= d) incl += o;
    }
    const int wpre = incl - reserved;
    const int total = __shfl_sync(kFull, incl, 31);
    const int64_t base = __shfl_sync(kFull, off, 0);
    // j / w without a per-pair integer division: one reciprocal per Gaussian, exact for j * w < 2^32
    // (j < number of tiles <= 48 K, w <= tiles per row)
    const unsigned magic = (w > 1) ? 0xFFFFFFFFu / (unsigned)w + 1u : 0u;

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
        const unsigned o_magic = __shfl_sync(kFull, magic, owner);
        const int o_tx0 = __shfl_sync(kFull, tx0, owner);
        const int o_ty0 = __shfl_sync(kFull, ty0, owner);
        const unsigned o_g = __shfl_sync(kFull, g, owner);
        if (k < total && base + k < p) {
            uint64_t pair = 0;
            unsigned tile = 0;  // filler slots of quirk A.2: (tile 0, Gaussian 0)
            if (j < o_emit) {
                const int q = (o_w > 1) ? (int)__umulhi((unsigned)j, o_magic) : j;   // j / o_w
                const int ty = o_ty0 + q, tx = o_tx0 + (j - q * o_w);  // ty outer, tx inner (:63-64)
                tile = (unsigned)(ty * ntx + tx);
                pair = ((uint64_t)tile << 32) | (uint64_t)o_g;
            }
            __stcs(pairs + base + k, pair);
            if (kHist) atomicAdd(&s_tile[tile], 1u);
        }
    }
}

__global__ void __launch_bounds__(kDupBlock)
k_duplicate_sorted(int64_t n, int width, int height, int ntx, int nty,
                   const uint64_t* __restrict__ sorted_elts, const float* __restrict__ means_2d,
                   const int* __restrict__ radii, const int* __restrict__ tiles_touched,
                   const int* __restrict__ offsets /* in sorted order */, int64_t p_cap,
                   const int64_t* __restrict__ p_dev, uint64_t* __restrict__ pairs) {
    // pairs beyond the capacity of the caller's buffers are dropped (the frame is then flagged as
    // overflowed by the scan's P > capacity, see cugs_b200_render_forward)
    const int64_t p = p_dev ? min(*p_dev, p_cap) : p_cap;
    const int64_t s0 = ((int64_t)blockIdx.x * (kDupBlock / 32) + (threadIdx.x >> 5)) * 32;
    if (s0 >= n) return;
    dup_warp<false>(s0, n, width, height, ntx, nty, sorted_elts, means_2d, radii, tiles_touched, offsets, p, pairs,
                    nullptr);
}

// duplicateWithKeys + the pair sort's histograms in ONE pass: a persistent grid of fat blocks (one or two per
// SM) walks the depth-sorted list; every emitted pair also bumps the block's shared tile histogram, which
// is folded into the digit histograms of the pair sort's passes and flushed once per block. This removes the
// separate read of all P pairs that k_packed_histogram<true> needs (149 MB at 3 M Gaussians / 1080p).
constexpr int kDupHistThreads = 1024;

__global__ void __launch_bounds__(kDupHistThreads)
k_duplicate_sorted_hist(int64_t n, int width, int height, int ntx, int nty,
                        const uint64_t* __restrict__ sorted_elts, const float* __restrict__ means_2d,
                        const int* __restrict__ radii, const int* __restrict__ tiles_touched,
                        const int* __restrict__ offsets, int64_t p_cap, const int64_t* __restrict__ p_dev,
                        uint64_t* __restrict__ pairs, PackedPlan plan, unsigned* __restrict__ digit_hist,
                        int num_tiles, unsigned* __restrict__ tile_hist) {
    extern __shared__ unsigned s_hist[];  // [passes*256] + [num_tiles]
    unsigned* s_tile = s_hist + plan.passes * kPkRadix;
    const int total_bins = plan.passes * kPkRadix + num_tiles;
    for (int b = threadIdx.x; b < total_bins; b += kDupHistThreads) s_hist[b] = 0;
    __syncthreads();
    const int64_t p = p_dev ? min(*p_dev, p_cap) : p_cap;
    const int64_t chunk = kDupHistThreads;  // Gaussians per block iteration (32 per warp)
    for (int64_t c0 = (int64_t)blockIdx.x * chunk; c0 < n; c0 += (int64_t)gridDim.x * chunk) {
        const int64_t s0 = c0 + (int64_t)(threadIdx.x >> 5) * 32;
        if (s0 < n)
            dup_warp<true>(s0, n, width, height, ntx, nty, sorted_elts, means_2d, radii, tiles_touched, offsets, p,
                           pairs, s_tile);
    }
    __syncthreads();
    // every digit of a pair's key is a function of its tile id: fold the digit histograms out of the tile histogram
    for (int b = threadIdx.x; b < num_tiles; b += kDupHistThreads) {
        const unsigned c = s_tile[b];
        if (c) {
            for (int ps = 0; ps < plan.passes; ++ps)
                atomicAdd(&s_hist[ps * kPkRadix + (((unsigned)b >> plan.shift[ps]) & ((1u << plan.bits[ps]) - 1))], c);
            atomicAdd(&tile_hist[b], c);
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < plan.passes * kPkRadix; b += kDupHistThreads) {
        const unsigned c = s_hist[b];
        if (c) atomicAdd(&digit_hist[b], c);
    }
}

}  // namespace cugs

using namespace cugs;

// ------------------------------------------------------------------------------------------------
// host-side launchers (internal; used by api.cu)
// ------------------------------------------------------------------------------------------------
static size_t packed_temp_layout_bytes(int64_t tiles, int passes, int num_tiles) {
    return (size_t)kPkMaxPasses * kPkRadix * 4 + 64 + align_up((size_t)(num_tiles > 0 ? num_tiles : 1) * 4, 256) +
           (size_t)passes * (size_t)(tiles > 0 ? tiles : 1) * kPkRadix * 4;
}
// bytes of `temp` a sort of n elements really touches (histograms + the look-back state of its own tile size)
static size_t packed_sort_used_bytes(int64_t n, int passes, int num_tiles) {
    const int64_t tile_elts = (int64_t)kPkThreads * pk_items_for(n);
    return packed_temp_layout_bytes((n + tile_elts - 1) / tile_elts, passes, num_tiles);
}
// What a caller must provide for a CAPACITY of n elements: sized on the smallest tile any variant uses, so that the
// figure is monotone in n and a buffer sized for a capacity serves every smaller count (a count just below the
// large-tile limit has twice the tiles of a capacity just above it).
size_t cugs_packed_sort_temp_bytes(int64_t n, int passes, int num_tiles) {
    const int64_t tile_elts = (int64_t)kPkThreads * ((int64_t)CUGS_PK_SMALL_LIMIT > 0 ? kPkItemsSmall : kPkItems);
    return packed_temp_layout_bytes((n + tile_elts - 1) / tile_elts, passes, num_tiles);
}

int cugs_packed_passes(int key_bits) { return make_packed_plan(key_bits).passes; }

// Stable LSD sort of n packed elements by the low `key_bits` bits of their HIGH word.
//   a, b           : ping-pong buffers, input in a
//   out32_last     : if non-null the last pass writes only the low words there (result elements
//                    are then NOT materialised); else the result is in (passes odd ? b : a)
//   tile_ranges    : if non-null (num_tiles > 0) the full key histogram is taken in the same read
//                    as the digit histograms and turned into [start,end) ranges
//   n_dev          : if non-null, n is only the CAPACITY of a / b / out32_last and the element count is
//                    min(*n_dev, n), read on the device (no host round trip: launches are sized on n)
//   hist_done      : the digit / tile histograms in `temp` were already filled (k_duplicate_sorted_hist after
//                    cugs_packed_sort_prepare): no memset, no histogram pass
//   payload_src/dst: if non-null the LAST pass also writes payload_dst[final position] = payload_src[low word]
int cugs_packed_sort(cugs_handle_t* h, cudaStream_t s, int64_t n, int key_bits, uint64_t* a, uint64_t* b,
                     int* out32_last, int num_tiles, int* tile_ranges, void* temp, size_t temp_bytes,
                     const int64_t* n_dev, bool hist_done, const int* payload_src, int* payload_dst) {
    const PackedPlan plan = make_packed_plan(key_bits);
    if (n >= (1ll << 30))
        return set_error(h, CUGS_ERR_UNSUPPORTED, "n = %lld >= 2^30 elements is not supported", (long long)n);
    const size_t need = packed_sort_used_bytes(n, plan.passes, num_tiles);
    if (temp_bytes < need)
        return set_error(h, CUGS_ERR_WORKSPACE, "packed sort temp too small: %zu < %zu", temp_bytes, need);
    if (tile_ranges && num_tiles > 0 && n == 0) {
        CUGS_CUDA_TRY(h, cudaMemsetAsync(tile_ranges, 0, (size_t)num_tiles * 8, s));
        return CUGS_OK;
    }
    if (n == 0) return CUGS_OK;
    const int items = pk_items_for(n);
    const int64_t tile_elts = (int64_t)kPkThreads * items;
    const int64_t tiles = (n + tile_elts - 1) / tile_elts;
    unsigned* digit_hist = reinterpret_cast<unsigned*>(temp);
    unsigned* tickets = digit_hist + kPkMaxPasses * kPkRadix;
    unsigned* tile_hist = tickets + 16;
    unsigned* lookback = tile_hist + align_up((size_t)(num_tiles > 0 ? num_tiles : 1) * 4, 256) / 4;
    if (!hist_done) CUGS_CUDA_TRY(h, cudaMemsetAsync(temp, 0, need, s));

    int hist_blocks = h->sm_count;  // 148 x 8160 global atomics in the flush instead of 444 x 8160
    const int64_t hist_tile = (int64_t)kPkHistThreads * kPkHistItems;
    if ((int64_t)hist_blocks * hist_tile > n) hist_blocks = (int)((n + hist_tile - 1) / hist_tile);
    const bool want_ranges = tile_ranges != nullptr && num_tiles > 0;
    const size_t hist_smem = (size_t)plan.passes * kPkRadix * 4 + (want_ranges ? (size_t)num_tiles * 4 : 0);
    if (hist_done) {
        // nothing: filled by k_duplicate_sorted_hist
    } else if (want_ranges) {
        if (hist_smem > 200 * 1024)
            return set_error(h, CUGS_ERR_UNSUPPORTED, "%d tiles exceed the shared-memory tile histogram", num_tiles);
        CUGS_CUDA_TRY(h, cudaFuncSetAttribute(k_packed_histogram<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)hist_smem));
        k_packed_histogram<true><<<hist_blocks, kPkHistThreads, hist_smem, s>>>(n, n_dev, a, plan, digit_hist,
                                                                                num_tiles, tile_hist);
    } else {
        k_packed_histogram<false><<<hist_blocks, kPkHistThreads, hist_smem, s>>>(n, n_dev, a, plan, digit_hist, 0,
                                                                                 nullptr);
    }
    if (!hist_done) CUGS_LAUNCH_CHECK(h, "k_packed_histogram");
    k_packed_scan_bins<<<plan.passes, kPkRadix, 0, s>>>(digit_hist);
    CUGS_LAUNCH_CHECK(h, "k_packed_scan_bins");
    if (want_ranges) {
        k_tile_hist_to_ranges<<<1, 1024, 0, s>>>(num_tiles, tile_hist, tile_ranges);
        CUGS_LAUNCH_CHECK(h, "k_tile_hist_to_ranges");
    }
    uint64_t* src = a;
    uint64_t* dst = b;
    for (int ps = 0; ps < plan.passes; ++ps) {
        const bool final_pass = ps == plan.passes - 1;
        const bool last = final_pass && out32_last != nullptr;
        unsigned* lb = lookback + (size_t)ps * tiles * kPkRadix;
#define CUGS_OS_LAUNCH_I(LAST, BITS, ITEMS)                                                                     \
    do {                                                                                                        \
        CUGS_CUDA_TRY(h, cudaFuncSetAttribute(k_onesweep_packed<LAST, BITS, ITEMS>,                             \
                                              cudaFuncAttributeMaxDynamicSharedMemorySize,                      \
                                              (int)pk_smem_bytes(ITEMS)));                                      \
        k_onesweep_packed<LAST, BITS, ITEMS><<<(unsigned)tiles, kPkThreads, pk_smem_bytes(ITEMS), s>>>(         \
            n, n_dev, src, dst, LAST ? out32_last : nullptr, digit_hist + ps * kPkRadix, lb, tickets + ps,      \
            plan.shift[ps], plan.bits[ps], final_pass ? payload_src : nullptr, final_pass ? payload_dst : nullptr); \
    } while (0)
#define CUGS_OS_LAUNCH(LAST, BITS)                                                  \
    do {                                                                            \
        if (items == kPkItemsSmall) CUGS_OS_LAUNCH_I(LAST, BITS, kPkItemsSmall);    \
        else if (items == kPkItemsLarge) CUGS_OS_LAUNCH_I(LAST, BITS, kPkItemsLarge); \
        else CUGS_OS_LAUNCH_I(LAST, BITS, kPkItems);                                \
    } while (0)
#define CUGS_OS_DISPATCH(LAST)                          \
    switch (plan.bits[ps]) {                            \
        case 8: CUGS_OS_LAUNCH(LAST, 8); break;         \
        case 7: CUGS_OS_LAUNCH(LAST, 7); break;         \
        case 6: CUGS_OS_LAUNCH(LAST, 6); break;         \
        default: CUGS_OS_LAUNCH(LAST, 0); break;        \
    }
        if (last) { CUGS_OS_DISPATCH(true) } else { CUGS_OS_DISPATCH(false) }
#undef CUGS_OS_DISPATCH
#undef CUGS_OS_LAUNCH
#undef CUGS_OS_LAUNCH_I
        CUGS_LAUNCH_CHECK(h, "k_onesweep_packed");
        uint64_t* t = src; src = dst; dst = t;
    }
    return CUGS_OK;
}

// duplicateWithKeys fused with the pair sort's histograms: clears `sort_temp` (the temp of the pair sort that
// follows, which must then be called with hist_done = true), then one persistent pass
int cugs_duplicate_sorted_hist(cugs_handle_t* h, cudaStream_t s, int64_t n, int width, int height,
                               const uint64_t* sorted_elts, const float* means_2d, const int32_t* radii,
                               const int32_t* tiles_touched, const int32_t* offsets_sorted, int64_t p, uint64_t* pairs,
                               const int64_t* p_dev, int key_bits, int num_tiles, void* sort_temp,
                               size_t sort_temp_bytes) {
    const PackedPlan plan = make_packed_plan(key_bits);
    const size_t need = packed_sort_used_bytes(p, plan.passes, num_tiles);  // what the sort that follows touches
    if (sort_temp_bytes < need)
        return set_error(h, CUGS_ERR_WORKSPACE, "packed sort temp too small: %zu < %zu", sort_temp_bytes, need);
    CUGS_CUDA_TRY(h, cudaMemsetAsync(sort_temp, 0, need, s));
    if (n == 0 || p == 0) return CUGS_OK;
    unsigned* digit_hist = reinterpret_cast<unsigned*>(sort_temp);
    unsigned* tile_hist = digit_hist + kPkMaxPasses * kPkRadix + 16;
    const size_t smem = (size_t)plan.passes * kPkRadix * 4 + (size_t)num_tiles * 4;
    if (smem > 200 * 1024)
        return set_error(h, CUGS_ERR_UNSUPPORTED, "%d tiles exceed the shared-memory tile histogram", num_tiles);
    CUGS_CUDA_TRY(h, cudaFuncSetAttribute(k_duplicate_sorted_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int ntx = (width + kTile - 1) / kTile, nty = (height + kTile - 1) / kTile;
    int blocks = h->sm_count * (smem <= 100 * 1024 ? 2 : 1);
    const int64_t chunks = (n + kDupHistThreads - 1) / kDupHistThreads;
    if ((int64_t)blocks > chunks) blocks = (int)chunks;
    k_duplicate_sorted_hist<<<blocks, kDupHistThreads, smem, s>>>(n, width, height, ntx, nty, sorted_elts, means_2d,
                                                                  radii, tiles_touched, offsets_sorted, p, p_dev, pairs,
                                                                  plan, digit_hist, num_tiles, tile_hist);
    CUGS_LAUNCH_CHECK(h, "k_duplicate_sorted_hist");
    return CUGS_OK;
}

int cugs_duplicate_sorted(cugs_handle_t* h, cudaStream_t s, int64_t n, int width, int height,
                          const uint64_t* sorted_elts, const float* means_2d, const int32_t* radii,
                          const int32_t* tiles_touched, const int32_t* offsets_sorted, int64_t p, uint64_t* pairs,
                          const int64_t* p_dev) {
    if (n == 0 || p == 0) return CUGS_OK;
    const int ntx = (width + kTile - 1) / kTile, nty = (height + kTile - 1) / kTile;
    const unsigned grid = (unsigned)((n + kDupBlock - 1) / kDupBlock);
    k_duplicate_sorted<<<grid, kDupBlock, 0, s>>>(n, width, height, ntx, nty, sorted_elts, means_2d, radii,
                                                  tiles_touched, offsets_sorted, p, p_dev, pairs);
    CUGS_LAUNCH_CHECK(h, "k_duplicate_sorted");
    return CUGS_OK;
}

// ------------------------------------------------------------------------------------------------
// public stage entry points of the packed sort (declared in include/cugs_b200.h)
// ------------------------------------------------------------------------------------------------
extern "C" int cugs_b200_sort_packed_passes(int key_bits) { return cugs_packed_passes(key_bits); }

extern "C" size_t cugs_b200_sort_packed_temp_bytes(int64_t n, int key_bits, int num_tiles) {
    return cugs_packed_sort_temp_bytes(n, cugs_packed_passes(key_bits), num_tiles);
}

extern "C" int cugs_b200_sort_packed(cugs_handle_t* h, void* stream, int64_t n, int key_bits, uint64_t* elts_a,
                                     uint64_t* elts_b, int32_t* out32_last, int num_tiles, int32_t* tile_ranges,
                                     void* temp, size_t temp_bytes, const int64_t* n_dev) {
    CUGS_REQUIRE(h, h != nullptr, "handle is null");
    CUGS_REQUIRE(h, n >= 0, "n must be >= 0");
    CUGS_REQUIRE(h, key_bits >= 0 && key_bits <= 32, "key_bits must be in 0..32");
    CUGS_REQUIRE(h, num_tiles >= 0 && (tile_ranges == nullptr || num_tiles > 0), "tile_ranges needs num_tiles > 0");
    CUGS_REQUIRE(h, tile_ranges == nullptr || key_bits >= 32 || (int64_t)num_tiles <= (1ll << key_bits),
                 "num_tiles exceeds 2^key_bits");
    CUGS_REQUIRE(h, n == 0 || (elts_a && elts_b && temp), "null pointer");
    return cugs_packed_sort(h, (cudaStream_t)stream, n, key_bits, elts_a, elts_b, out32_last, num_tiles, tile_ranges,
                            temp, temp_bytes, n_dev, false, nullptr, nullptr);
}
