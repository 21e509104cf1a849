// density_control.cu — the model-resizing steps that follow the hot path on a schedule (SURVEY 8f
// rows 2 and 3): ADC clone / split / prune (optimizer/densification.cpp:94-329) and MCMC relocation
// (optimizer/mcmc_densification.cpp:56-138) as stream-compaction kernels over the five parameter
// arrays (and, optionally, the Adam moments), instead of the reference's chain of boolean-mask
// indexing + torch::cat (a dozen full-model temporaries per call).
//
// The POLICY stays the reference's: same masks, same thresholds, same output order
//   [kept originals | clones | split children (1st draw) | split children (2nd draw)]
// (densification.cpp:176, :271-286, :291-315), same relocation rule. Only the random draws differ
// (Philox here, torch's generator there); every kernel can export its normals so that tests check
// the arithmetic exactly and the draws statistically.
#include "common.cuh"

namespace cugs {

constexpr int kRowsPerBlock = 256;
constexpr unsigned kKeep = CUGS_DENSIFY_KEEP, kClone = CUGS_DENSIFY_CLONE, kSplit = CUGS_DENSIFY_SPLIT;

__device__ __forceinline__ float sigmoid_ref(float x) { return 1.0f / (1.0f + expf(-x)); }  // ATen sigmoid, f32

// flags[] stores KEEP = "passes compute_keep_mask"; an original that is split is replaced by its two
// children whatever its keep bit says (densification.cpp:296-303), so the movers see KEEP cleared.
__device__ __forceinline__ unsigned kept_view(unsigned f) { return (f & kSplit) ? (f & ~kKeep) : f; }

// ------------------------------------------------------------------------------------------------
// ADC: classification (compute_clone_mask :351-370, compute_split_mask :372-399, compute_keep_mask
// :401-444). counts = {kept originals, clones, splits}; the last block copies them to pinned memory.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRowsPerBlock)
k_densify_classify(int64_t n, const float* __restrict__ scales, const float* __restrict__ opacities,
                   const float* __restrict__ grad_accum, const float* __restrict__ grad_count,
                   const float* __restrict__ max_radii, cugs_densify_config_t cfg, uint8_t* __restrict__ flags,
                   unsigned long long* __restrict__ counts, unsigned* __restrict__ ticket,
                   int64_t* __restrict__ pinned_out) {
    __shared__ unsigned s_cnt[3];
    __shared__ bool s_last;
    if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * kRowsPerBlock + threadIdx.x;
    unsigned f = 0;
    if (i < n) {
        const float avg = __fdiv_rn(grad_accum[i], fmaxf(grad_count[i], 1.0f));
        const bool high = avg >= cfg.grad_threshold;
        // max over the axes of exp(scale) == exp(max scale): expf is monotone
        const float max_scale = expf(fmaxf(fmaxf(scales[i * 3], scales[i * 3 + 1]), scales[i * 3 + 2]));
        const bool small = max_scale < cfg.size_threshold;
        if (high && small) f |= kClone;
        if (high && !small) f |= kSplit;  // large = max_scale >= size_threshold
        bool keep = sigmoid_ref(opacities[i]) >= cfg.opacity_threshold;
        if (cfg.apply_size_pruning) {
            if (cfg.max_screen_size > 0.0f && max_radii != nullptr) keep = keep && (max_radii[i] <= cfg.max_screen_size);
            keep = keep && (max_scale <= cfg.ws_threshold);
        }
        if (keep) f |= kKeep;
        flags[i] = (uint8_t)f;
    }
    f = kept_view(f);
    const unsigned bk = __ballot_sync(kFull, f & kKeep), bc = __ballot_sync(kFull, f & kClone),
                   bs = __ballot_sync(kFull, f & kSplit);
    if ((threadIdx.x & 31) == 0) {
        if (bk) atomicAdd(&s_cnt[0], (unsigned)__popc(bk));
        if (bc) atomicAdd(&s_cnt[1], (unsigned)__popc(bc));
        if (bs) atomicAdd(&s_cnt[2], (unsigned)__popc(bs));
    }
    __syncthreads();
    if (threadIdx.x < 3 && s_cnt[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (s_last && threadIdx.x < 3) {
        __threadfence();
        pinned_out[threadIdx.x] = (int64_t)atomicAdd(&counts[threadIdx.x], 0ull);
    }
}

// counts of the three flags per block of kRowsPerBlock rows (flags may have been edited by the host
// policy -- budget caps -- between classify and apply, so apply re-derives everything from them)
__global__ void __launch_bounds__(kRowsPerBlock)
k_flag_block_counts(int64_t n, const uint8_t* __restrict__ flags, unsigned* __restrict__ block_counts, int64_t nb) {
    __shared__ unsigned s_cnt[3];
    if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * kRowsPerBlock + threadIdx.x;
    const unsigned f = (i < n) ? kept_view(flags[i]) : 0u;
    const unsigned bk = __ballot_sync(kFull, f & kKeep), bc = __ballot_sync(kFull, f & kClone),
                   bs = __ballot_sync(kFull, f & kSplit);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s_cnt[0], (unsigned)__popc(bk));
        atomicAdd(&s_cnt[1], (unsigned)__popc(bc));
        atomicAdd(&s_cnt[2], (unsigned)__popc(bs));
    }
    __syncthreads();
    if (threadIdx.x < 3) block_counts[(size_t)threadIdx.x * nb + blockIdx.x] = s_cnt[threadIdx.x];
}

// exclusive scan of `rows` independent rows of nb counters (one block per row); totals[row] = sum
__global__ void __launch_bounds__(1024)
k_scan_rows(int64_t nb, unsigned* __restrict__ data, unsigned long long* __restrict__ totals) {
    __shared__ unsigned s_warp[32];
    __shared__ unsigned long long s_carry;
    unsigned* row = data + (size_t)blockIdx.x * nb;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < nb; base += 1024) {
        const int64_t j = base + threadIdx.x;
        const unsigned c = (j < nb) ? row[j] : 0u;
        unsigned incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned o = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        unsigned off = 0;
        for (int w = 0; w < warp; ++w) off += s_warp[w];
        const unsigned long long carry = s_carry;
        if (j < nb) row[j] = (unsigned)(carry + off + incl - c);  // < 2^32: the model has < 2^31 rows
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + off + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[blockIdx.x] = s_carry;
}

struct DensifyArrays {
    const float* src[5];
    float* dst[5];
    const float* src_m[5];
    const float* src_v[5];
    float* dst_m[5];
    float* dst_v[5];
    int width[5];
};

// One block moves kRowsPerBlock source rows of all five arrays (Adam group order: positions,
// sh_coeffs, opacities, scales, rotations) to their destinations. Reads are fully coalesced; kept rows
// land contiguously (stable compaction), clones and children likewise in their own segments.
__global__ void __launch_bounds__(kRowsPerBlock)
k_densify_move(int64_t n, int64_t n_out, const uint8_t* __restrict__ flags, const unsigned* __restrict__ block_off,
               int64_t nb, const unsigned long long* __restrict__ totals, DensifyArrays a, float log_split_factor,
               unsigned seed_lo, unsigned seed_hi, float* __restrict__ split_normals /* optional [2S,3] */) {
    __shared__ unsigned s_flag[kRowsPerBlock];
    __shared__ int64_t s_dst[3][kRowsPerBlock];  // destination row as kept original / clone / first child
    __shared__ unsigned s_wcnt[3][kRowsPerBlock / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t row0 = (int64_t)blockIdx.x * kRowsPerBlock;
    const int rows = (int)min((int64_t)kRowsPerBlock, n - row0);
    const int64_t K = (int64_t)totals[0], Cn = (int64_t)totals[1], S = (int64_t)totals[2];

    const unsigned f = (tid < rows) ? kept_view(flags[row0 + tid]) : 0u;
    const unsigned lt = (1u << lane) - 1;
    const unsigned bal[3] = {__ballot_sync(kFull, f & kKeep), __ballot_sync(kFull, f & kClone),
                             __ballot_sync(kFull, f & kSplit)};
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) s_wcnt[c][warp] = (unsigned)__popc(bal[c]);
    }
    __syncthreads();
    {
        const int64_t seg_base[3] = {0, K, K + Cn};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            unsigned off = 0;
            for (int w = 0; w < warp; ++w) off += s_wcnt[c][w];
            s_dst[c][tid] = seg_base[c] + (int64_t)block_off[(size_t)c * nb + blockIdx.x] + off + __popc(bal[c] & lt);
        }
        s_flag[tid] = f;
    }
    __syncthreads();

#pragma unroll 1
    for (int g = 0; g < 5; ++g) {
        const int w = a.width[g];
        const float* __restrict__ src = a.src[g] + row0 * w;
        const float* __restrict__ sm = a.src_m[g] ? a.src_m[g] + row0 * w : nullptr;
        const float* __restrict__ sv = a.src_v[g] ? a.src_v[g] + row0 * w : nullptr;
        float* __restrict__ dst = a.dst[g];
        float* __restrict__ dm = a.dst_m[g];
        float* __restrict__ dv = a.dst_v[g];
        const int total = rows * w;
        for (int e = tid; e < total; e += kRowsPerBlock) {
            const int r = e / w, c = e - r * w;
            const unsigned fl = s_flag[r];
            if (fl == 0) continue;
            const float val = src[e];
            if (fl & kKeep) {
                const int64_t d = s_dst[0][r];
                if (d < n_out) {
                    dst[d * w + c] = val;
                    if (dm) dm[d * w + c] = sm ? sm[e] : 0.0f;   // optimizer state is carried over
                    if (dv) dv[d * w + c] = sv ? sv[e] : 0.0f;
                }
            }
            if (fl & kClone) {
                const int64_t d = s_dst[1][r];
                if (d < n_out) {
                    dst[d * w + c] = val;
                    if (dm) dm[d * w + c] = 0.0f;
                    if (dv) dv[d * w + c] = 0.0f;
                }
            }
            if (fl & kSplit) {
                const int64_t d1 = s_dst[2][r], d2 = d1 + S;
                float v1 = val, v2 = val;
                if (g == 3) {  // scales: new_scale = old - log(1.6)                      (:243-245)
                    v1 = v2 = add_rn(val, -log_split_factor);
                } else if (g == 0) {  // positions: old + randn * exp(new_scale)           (:249-253)
                    const float e_new = expf(add_rn(a.src[3][(row0 + r) * 3 + c], -log_split_factor));
                    const int64_t i = row0 + r;
                    float z[3];
#pragma unroll
                    for (int child = 0; child < 2; ++child) {
                        philox_normal3((unsigned)i, (unsigned)((uint64_t)i >> 32), (unsigned)child, 0x5eed5b17u, seed_lo,
                                       seed_hi, z[0], z[1], z[2]);
                        const float zc = (c == 0) ? z[0] : (c == 1 ? z[1] : z[2]);
                        const float child_pos = add_rn(val, mul_rn(zc, e_new));
                        if (child == 0) v1 = child_pos; else v2 = child_pos;
                        if (split_normals && d2 < n_out) split_normals[((d1 - K - Cn) + child * S) * 3 + c] = zc;
                    }
                }
                if (d2 < n_out) {
                    dst[d1 * w + c] = v1;
                    dst[d2 * w + c] = v2;
                    if (dm) { dm[d1 * w + c] = 0.0f; dm[d2 * w + c] = 0.0f; }
                    if (dv) { dv[d1 * w + c] = 0.0f; dv[d2 * w + c] = 0.0f; }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// MCMC relocation (mcmc_densification.cpp:56-138)
// ------------------------------------------------------------------------------------------------
constexpr int kRelBlock = 1024;

// per block of 1024 Gaussians: number of dead ones and the summed sampling weight of the alive ones
__global__ void __launch_bounds__(kRelBlock)
k_relocate_classify(int64_t n, const float* __restrict__ opacities, float dead_threshold,
                    unsigned* __restrict__ block_dead, double* __restrict__ block_weight,
                    double* __restrict__ warp_weight /* [blocks * 32] */) {
    __shared__ unsigned s_dead;
    __shared__ double s_w[kRelBlock / 32];
    if (threadIdx.x == 0) s_dead = 0;
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * kRelBlock + threadIdx.x;
    float w = 0.0f;
    bool dead = false;
    if (i < n) {
        w = sigmoid_ref(opacities[i]);
        dead = w < dead_threshold;
        if (dead) w = 0.0f;
    }
    const unsigned b = __ballot_sync(kFull, dead);
    double acc = (double)w;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(kFull, acc, d);
    if ((threadIdx.x & 31) == 0) {
        if (b) atomicAdd(&s_dead, (unsigned)__popc(b));
        s_w[threadIdx.x >> 5] = acc;
        warp_weight[(size_t)blockIdx.x * (kRelBlock / 32) + (threadIdx.x >> 5)] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < kRelBlock / 32; ++k) t += s_w[k];
        block_dead[blockIdx.x] = s_dead;
        block_weight[blockIdx.x] = t;
    }
}

// single block: exclusive scan of block_dead, inclusive scan (cdf) of block_weight, totals
__global__ void __launch_bounds__(1024)
k_relocate_scan(int64_t nb, unsigned* __restrict__ block_dead, double* __restrict__ block_cdf,
                unsigned long long* __restrict__ totals /* {dead} */, double* __restrict__ total_weight,
                int64_t* __restrict__ pinned_out, int64_t n, int64_t max_relocate) {
    __shared__ unsigned s_warp[32];
    __shared__ double s_wsum[32];
    __shared__ unsigned long long s_carry;
    __shared__ double s_wcarry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s_carry = 0; s_wcarry = 0.0; }
    __syncthreads();
    for (int64_t base = 0; base < nb; base += 1024) {
        const int64_t j = base + threadIdx.x;
        const unsigned c = (j < nb) ? block_dead[j] : 0u;
        const double wv = (j < nb) ? block_cdf[j] : 0.0;
        unsigned incl = c;
        double wincl = wv;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned o = __shfl_up_sync(kFull, incl, d);
            const double wo = __shfl_up_sync(kFull, wincl, d);
            if (lane >= d) { incl += o; wincl += wo; }
        }
        if (lane == 31) { s_warp[warp] = incl; s_wsum[warp] = wincl; }
        __syncthreads();
        unsigned off = 0;
        double woff = 0.0;
        for (int w = 0; w < warp; ++w) { off += s_warp[w]; woff += s_wsum[w]; }
        const unsigned long long carry = s_carry;
        const double wcarry = s_wcarry;
        if (j < nb) {
            block_dead[j] = (unsigned)(carry + off + incl - c);
            block_cdf[j] = wcarry + woff + wincl;
        }
        __syncthreads();
        if (threadIdx.x == 1023) { s_carry = carry + off + incl; s_wcarry = wcarry + woff + wincl; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        totals[0] = s_carry;
        *total_weight = s_wcarry;
        const int64_t dead = (int64_t)s_carry;
        const bool act = dead > 0 && dead < n;  // nothing to do without dead or without alive ones (:90-92)
        if (pinned_out) {
            pinned_out[0] = dead;
            pinned_out[1] = act ? min(dead, max_relocate) : 0;
        }
    }
}

// thread per Gaussian; a dead one whose rank among the dead is below the cap draws its source from the
// alive ones with probability proportional to sigmoid(opacity) (inverse cdf: block level by binary
// search over the scanned block sums, then a walk over the block's 32 warp sums and over the 32 weights
// of that warp). Only reads the
// model: the copies happen in k_relocate_copy, so that every walk sees the opacities of before the call.
__global__ void __launch_bounds__(kRelBlock)
k_relocate_select(int64_t n, const float* __restrict__ opacities, float dead_threshold, int64_t max_relocate,
                  const unsigned* __restrict__ block_dead_off, const double* __restrict__ block_cdf,
                  const double* __restrict__ warp_weight, int64_t nb,
                  const unsigned long long* __restrict__ totals, const double* __restrict__ total_weight,
                  unsigned seed_lo, unsigned seed_hi, unsigned step, int32_t* __restrict__ source) {
    __shared__ unsigned s_wcnt[kRelBlock / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * kRelBlock + threadIdx.x;
    const int64_t num_dead = (int64_t)totals[0];
    const bool dead = (i < n) && (sigmoid_ref(opacities[i]) < dead_threshold);
    const unsigned b = __ballot_sync(kFull, dead);
    if (lane == 0) s_wcnt[warp] = (unsigned)__popc(b);
    __syncthreads();
    if (i >= n) return;
    int64_t src = -1;
    if (dead && num_dead < n) {  // without alive Gaussians nothing moves (:90-92)
        unsigned off = 0;
        for (int w = 0; w < warp; ++w) off += s_wcnt[w];
        const int64_t rank = (int64_t)block_dead_off[blockIdx.x] + off + __popc(b & ((1u << lane) - 1));
        if (rank < max_relocate) {  // capped: only the first max_relocate dead ones move (:101-103)
            unsigned r[4];  // 64 random bits -> uniform in [0, 1)
            philox4x32_10((unsigned)i, (unsigned)((uint64_t)i >> 32), step, 0x50a7ce5eu, seed_lo, seed_hi, r);
            const double u01 = ((double)r[0] * 4294967296.0 + (double)r[1]) * (1.0 / 18446744073709551616.0);
            const double target = u01 * (*total_weight);
            int64_t lo = 0, hi = nb - 1;  // first block whose inclusive cdf exceeds the target
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (block_cdf[mid] > target) hi = mid; else lo = mid + 1;
            }
            double acc = lo > 0 ? block_cdf[lo - 1] : 0.0;
            const double* ww = warp_weight + lo * (kRelBlock / 32);
            int wsel = -1, wlast = -1;
            for (int k = 0; k < kRelBlock / 32; ++k) {
                const double wk = ww[k];
                if (wk <= 0.0) continue;
                wlast = k;
                if (acc + wk > target) { wsel = k; break; }
                acc += wk;
            }
            if (wsel < 0) { wsel = wlast; acc -= (wlast >= 0 ? ww[wlast] : 0.0); }  // rounding at the block's end
            int64_t last_alive = -1;
            if (wsel >= 0) {
                const int64_t j0 = lo * kRelBlock + (int64_t)wsel * 32, j1 = min(n, j0 + 32);
                for (int64_t j = j0; j < j1; ++j) {
                    const float w = sigmoid_ref(opacities[j]);
                    if (w < dead_threshold) continue;
                    last_alive = j;
                    acc += (double)w;
                    if (acc > target) { src = j; break; }
                }
            }
            if (src < 0) src = last_alive;  // summation-order rounding at the end of the warp
        }
    }
    source[i] = (int32_t)src;
}

// sources are alive (never written), targets are dead (each written by its own threads): no hazards.
// One warp per Gaussian row so that the SH row (3C floats) moves coalesced.
__global__ void __launch_bounds__(256)
k_relocate_copy(int64_t n, int num_coeffs, float* __restrict__ positions, float* __restrict__ sh,
                float* __restrict__ opacities, float* __restrict__ scales, float* __restrict__ rotations,
                const int32_t* __restrict__ source, float scene_extent, float log_shrink, float low_opacity,
                unsigned seed_lo, unsigned seed_hi, unsigned step, float* __restrict__ normals_out /* optional [N,3] */) {
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    const int64_t src = source[i];
    if (src < 0) return;
    const int row = 3 * num_coeffs;
    for (int k = lane; k < row; k += 32) sh[i * row + k] = sh[src * row + k];
    if (lane < 4) rotations[i * 4 + lane] = rotations[src * 4 + lane];
    if (lane < 3) {
        float z[3];
        philox_normal3((unsigned)i, (unsigned)((uint64_t)i >> 32), step, 0x7e10ca7eu, seed_lo, seed_hi, z[0], z[1], z[2]);
        const float zk = (lane == 0) ? z[0] : (lane == 1 ? z[1] : z[2]);
        // source_pos + randn * scene_extent * 0.01f                                     (:120-122)
        positions[i * 3 + lane] = add_rn(positions[src * 3 + lane], mul_rn(mul_rn(zk, scene_extent), 0.01f));
        scales[i * 3 + lane] = add_rn(scales[src * 3 + lane], -log_shrink);  // 10x smaller (:125-127)
        if (normals_out) normals_out[i * 3 + lane] = zk;
    }
    if (lane == 0) opacities[i] = low_opacity;  // inverse_sigmoid(0.01) (:130-133)
}

}  // namespace cugs

using namespace cugs;

static inline size_t au(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" size_t cugs_b200_densify_temp_bytes(int64_t n) {
    const size_t nb = (size_t)((n + kRowsPerBlock - 1) / kRowsPerBlock) + 1;
    return 256 + au(3 * nb * sizeof(unsigned));
}

extern "C" int cugs_b200_densify_classify(cugs_handle_t* h, void* stream, int64_t n, const float* scales,
                                          const float* opacities, const float* grad_accum, const float* grad_count,
                                          const float* max_radii, const cugs_densify_config_t* cfg, uint8_t* flags,
                                          int64_t* counts_host, void* temp, size_t temp_bytes) {
    CUGS_REQUIRE(h, h != nullptr, "handle is null");
    CUGS_REQUIRE(h, n >= 0 && n < (int64_t)1 << 31, "n out of range");
    CUGS_REQUIRE(h, cfg != nullptr && counts_host != nullptr, "null config / counts");
    counts_host[0] = counts_host[1] = counts_host[2] = 0;
    if (n == 0) return CUGS_OK;
    CUGS_REQUIRE(h, scales && opacities && grad_accum && grad_count && flags && temp, "null pointer");
    if (temp_bytes < cugs_b200_densify_temp_bytes(n))
        return set_error(h, CUGS_ERR_WORKSPACE, "densify temp too small: %zu < %zu", temp_bytes,
                         cugs_b200_densify_temp_bytes(n));
    cudaStream_t s = (cudaStream_t)stream;
    CUGS_CUDA_TRY(h, cudaMemsetAsync(temp, 0, 256, s));
    unsigned long long* counts = reinterpret_cast<unsigned long long*>(temp);
    unsigned* ticket = reinterpret_cast<unsigned*>(counts + 4);
    const unsigned grid = (unsigned)((n + kRowsPerBlock - 1) / kRowsPerBlock);
    k_densify_classify<<<grid, kRowsPerBlock, 0, s>>>(n, scales, opacities, grad_accum, grad_count, max_radii, *cfg,
                                                      flags, counts, ticket, h->pinned + 4);
    CUGS_LAUNCH_CHECK(h, "k_densify_classify");
    CUGS_CUDA_TRY(h, cudaStreamSynchronize(s));  // the reference reads the three counts with .item() (:119, :186, :318)
    for (int k = 0; k < 3; ++k) counts_host[k] = h->pinned[4 + k];
    return CUGS_OK;
}

extern "C" int cugs_b200_densify_apply(cugs_handle_t* h, void* stream, int64_t n, int64_t n_out, int num_coeffs,
                                       const uint8_t* flags, const float* const* src, float* const* dst,
                                       const float* const* src_m, const float* const* src_v, float* const* dst_m,
                                       float* const* dst_v, uint64_t seed, float* split_normals_out, void* temp,
                                       size_t temp_bytes) {
    CUGS_REQUIRE(h, h != nullptr, "handle is null");
    CUGS_REQUIRE(h, n >= 0 && n < (int64_t)1 << 31 && n_out >= 0, "n out of range");
    CUGS_REQUIRE(h, num_coeffs >= 1, "num_coeffs must be >= 1");
    if (n == 0) return CUGS_OK;
    CUGS_REQUIRE(h, flags && src && dst && temp, "null pointer");
    CUGS_REQUIRE(h, (dst_m == nullptr) == (dst_v == nullptr), "dst_m and dst_v go together");
    CUGS_REQUIRE(h, (src_m == nullptr) == (src_v == nullptr), "src_m and src_v go together");
    if (temp_bytes < cugs_b200_densify_temp_bytes(n))
        return set_error(h, CUGS_ERR_WORKSPACE, "densify temp too small: %zu < %zu", temp_bytes,
                         cugs_b200_densify_temp_bytes(n));
    DensifyArrays a{};
    const int widths[5] = {3, 3 * num_coeffs, 1, 3, 4};
    for (int g = 0; g < 5; ++g) {
        CUGS_REQUIRE(h, src[g] && (n_out == 0 || dst[g]), "null parameter array");
        CUGS_REQUIRE(h, src[g] != dst[g], "densify_apply is out of place");
        a.src[g] = src[g];
        a.dst[g] = dst[g];
        a.src_m[g] = src_m ? src_m[g] : nullptr;
        a.src_v[g] = src_v ? src_v[g] : nullptr;
        a.dst_m[g] = dst_m ? dst_m[g] : nullptr;
        a.dst_v[g] = dst_v ? dst_v[g] : nullptr;
        a.width[g] = widths[g];
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t nb = (n + kRowsPerBlock - 1) / kRowsPerBlock;
    unsigned long long* totals = reinterpret_cast<unsigned long long*>(temp) + 8;  // after classify's counters
    unsigned* block_counts = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(temp) + 256);
    k_flag_block_counts<<<(unsigned)nb, kRowsPerBlock, 0, s>>>(n, flags, block_counts, nb);
    CUGS_LAUNCH_CHECK(h, "k_flag_block_counts");
    k_scan_rows<<<3, 1024, 0, s>>>(nb, block_counts, totals);
    CUGS_LAUNCH_CHECK(h, "k_scan_rows");
    const float log_split = std::log(1.6f);  // densification.cpp:244
    k_densify_move<<<(unsigned)nb, kRowsPerBlock, 0, s>>>(n, n_out, flags, block_counts, nb, totals, a, log_split,
                                                          (unsigned)seed, (unsigned)(seed >> 32), split_normals_out);
    CUGS_LAUNCH_CHECK(h, "k_densify_move");
    // n_out is the caller's claim (computed from counts it may have edited together with the flags): verify it
    // against the totals the scan just derived from the flags themselves. Densification runs on a 100-step
    // schedule and the classification before it blocks anyway, so one more stream synchronisation is free.
    unsigned long long tot[3] = {0, 0, 0};
    CUGS_CUDA_TRY(h, cudaMemcpyAsync(tot, totals, sizeof(tot), cudaMemcpyDeviceToHost, s));
    CUGS_CUDA_TRY(h, cudaStreamSynchronize(s));
    const long long expect = (long long)tot[0] + (long long)tot[1] + 2 * (long long)tot[2];
    if (expect != (long long)n_out)
        return set_error(h, CUGS_ERR_INVALID_ARG,
                         "n_out = %lld does not match the flags: kept %llu + clones %llu + 2 x splits %llu = %lld "
                         "(the destination was only written up to min(n_out, that))",
                         (long long)n_out, tot[0], tot[1], tot[2], expect);
    return CUGS_OK;
}

extern "C" size_t cugs_b200_mcmc_relocate_temp_bytes(int64_t n) {
    const size_t nb = (size_t)((n + kRelBlock - 1) / kRelBlock) + 1;
    return 256 + au(nb * sizeof(unsigned)) + au(nb * sizeof(double)) + au(nb * (kRelBlock / 32) * sizeof(double)) +
           au((size_t)(n > 0 ? n : 1) * sizeof(int32_t));
}

extern "C" int cugs_b200_mcmc_relocate(cugs_handle_t* h, void* stream, int64_t n, int num_coeffs, float* positions,
                                       float* sh_coeffs, float* opacities, float* scales, float* rotations,
                                       float dead_threshold, int64_t max_relocate, float scene_extent, uint64_t seed,
                                       uint32_t step, int32_t* source_out, float* normals_out, int64_t* counts_host,
                                       void* temp, size_t temp_bytes) {
    CUGS_REQUIRE(h, h != nullptr, "handle is null");
    CUGS_REQUIRE(h, n >= 0 && n < (int64_t)1 << 31, "n out of range");
    CUGS_REQUIRE(h, num_coeffs >= 1 && max_relocate >= 0, "bad argument");
    if (counts_host) counts_host[0] = counts_host[1] = 0;
    if (n == 0) return CUGS_OK;
    CUGS_REQUIRE(h, positions && sh_coeffs && opacities && scales && rotations && temp, "null pointer");
    if (temp_bytes < cugs_b200_mcmc_relocate_temp_bytes(n))
        return set_error(h, CUGS_ERR_WORKSPACE, "relocate temp too small: %zu < %zu", temp_bytes,
                         cugs_b200_mcmc_relocate_temp_bytes(n));
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t nb = (n + kRelBlock - 1) / kRelBlock;
    char* base = reinterpret_cast<char*>(temp);
    unsigned long long* totals = reinterpret_cast<unsigned long long*>(base);
    double* total_weight = reinterpret_cast<double*>(base + 64);
    unsigned* block_dead = reinterpret_cast<unsigned*>(base + 256);
    double* block_cdf = reinterpret_cast<double*>(base + 256 + au((size_t)(nb + 1) * sizeof(unsigned)));
    double* warp_weight = reinterpret_cast<double*>(reinterpret_cast<char*>(block_cdf) +
                                                    au((size_t)(nb + 1) * sizeof(double)));
    int32_t* source = source_out ? source_out
                                 : reinterpret_cast<int32_t*>(reinterpret_cast<char*>(warp_weight) +
                                                              au((size_t)(nb + 1) * (kRelBlock / 32) * sizeof(double)));
    k_relocate_classify<<<(unsigned)nb, kRelBlock, 0, s>>>(n, opacities, dead_threshold, block_dead, block_cdf,
                                                           warp_weight);
    CUGS_LAUNCH_CHECK(h, "k_relocate_classify");
    k_relocate_scan<<<1, 1024, 0, s>>>(nb, block_dead, block_cdf, totals, total_weight,
                                       counts_host ? h->pinned + 4 : nullptr, n, max_relocate);
    CUGS_LAUNCH_CHECK(h, "k_relocate_scan");
    k_relocate_select<<<(unsigned)nb, kRelBlock, 0, s>>>(n, opacities, dead_threshold, max_relocate, block_dead,
                                                         block_cdf, warp_weight, nb, totals, total_weight, (unsigned)seed,
                                                         (unsigned)(seed >> 32), step, source);
    CUGS_LAUNCH_CHECK(h, "k_relocate_select");
    const float log_shrink = std::log(10.0f);            // mcmc_densification.cpp:126
    const float low_opacity = std::log(0.01f / 0.99f);   // :130
    k_relocate_copy<<<(unsigned)((n * 32 + 255) / 256), 256, 0, s>>>(n, num_coeffs, positions, sh_coeffs, opacities,
                                                                    scales, rotations, source, scene_extent, log_shrink,
                                                                    low_opacity, (unsigned)seed, (unsigned)(seed >> 32),
                                                                    step, normals_out);
    CUGS_LAUNCH_CHECK(h, "k_relocate_copy");
    if (counts_host) {  // the reference reads num_dead with .item() (:86); pass NULL to stay asynchronous
        CUGS_CUDA_TRY(h, cudaStreamSynchronize(s));
        counts_host[0] = h->pinned[4];
        counts_host[1] = h->pinned[5];
    }
    return CUGS_OK;
}
