// grad_exchange.cu — row compaction of the parameter gradients around the view-parallel all-reduce.
//
// A view only produces gradients for the Gaussians its pixels actually blended (everything behind a
// saturated pixel, outside the frustum or culled gets exactly zero), so the dense 59-floats-per-
// Gaussian gradient arena that the ranks sum every step is mostly zeros (3 M / 1080p: ~75 %).
// preprocess_bwd records which Gaussians were touched (touch mask); the ranks MAX-reduce the masks
// (4 B/Gaussian), every rank compacts the rows of the union into one dense buffer with the kernels
// below, ONE all-reduce(sum) runs on the compact buffer, and the result is scattered back. Rows
// outside the union are zero on every rank and stay zero. The reference has no multi-GPU path at
// all; the dense variant (one all-reduce of the whole arena) remains available.
//
// Compact layout for M touched Gaussians (group-major, the Adam group order of
// optimizer/fused_adam.cu:94-97): positions [M,3] | sh_coeffs [M,3C] | opacities [M] | scales [M,3]
// | rotations [M,4], each block starting at a multiple of 4 floats (cugs_b200_compact_grad_floats).
#include "common.cuh"

namespace cugs {

struct GradGroups {
    float* g[5];  // positions, sh_coeffs, opacities, scales, rotations
};

// index list of the touched Gaussians: idx[offsets[i]] = i
__global__ void __launch_bounds__(256)
k_build_touch_index(int64_t n, const int* __restrict__ touch, const int* __restrict__ offsets, int* __restrict__ idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && touch[i] != 0) idx[offsets[i]] = (int)i;
}

// One thread per float of the compact buffer (fully coalesced on the compact side, w-float contiguous
// runs on the dense side: 192-byte runs for the SH group, which is 81 % of the bytes).
// blockIdx.y = parameter group. Two other work splits were measured in round 2 (2 x B200, 531 k rows = 125 MB
// each way, profiles/r02/NOTES.md) and were SLOWER than this one (0.105 ms): one warp per row with the index
// broadcast (0.17 ms: one dependent load chain per warp) and one thread per 16-byte chunk of a row with the
// groups interleaved inside a warp (0.13 ms: mixed float4 / scalar accesses in one warp).
__host__ __device__ __forceinline__ int64_t align4(int64_t x) { return (x + 3) & ~(int64_t)3; }

// m_dev != NULL: `m` is the row CAPACITY the compact layout was sized for (known to the host from an earlier
// step) and the real row count is read on the device; a gather zero-fills the rows between the two so that
// the all-reduce sums defined values, and rows beyond the capacity are dropped (status[1] flags it).
template <bool kGather>
__global__ void __launch_bounds__(256)
k_move_grad_rows(int C, const int* __restrict__ idx, int64_t m, const int64_t* __restrict__ m_dev, GradGroups dense,
                 float* __restrict__ compact) {
    const int grp = blockIdx.y;
    const int shw = 3 * C;
    const int64_t m_real = m_dev ? min(*m_dev, m) : m;
    // group bases, each rounded up to 4 floats so that the SH block can be moved as float4
    const int64_t b_sh = align4(3 * m), b_op = b_sh + align4((int64_t)shw * m), b_sc = b_op + align4(m),
                  b_ro = b_sc + align4(3 * m);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (grp == 1 && (shw & 3) == 0) {  // SH: 81 % of the bytes, rows are 16-byte aligned on both sides
        const int w4 = shw >> 2;
        const float4* __restrict__ d4 = reinterpret_cast<const float4*>(dense.g[1]);
        float4* __restrict__ d4w = reinterpret_cast<float4*>(dense.g[1]);
        float4* __restrict__ c4 = reinterpret_cast<float4*>(compact + b_sh);
        for (int64_t e = t0; e < m * w4; e += stride) {
            const int64_t j = e / w4;
            if (j >= m_real) {
                if (kGather) c4[e] = make_float4(0.f, 0.f, 0.f, 0.f);
                continue;
            }
            const int64_t src = (int64_t)idx[j] * w4 + (e - j * w4);
            if (kGather) c4[e] = d4[src];
            else d4w[src] = c4[e];
        }
        return;
    }
    const int w = (grp == 0) ? 3 : (grp == 1) ? shw : (grp == 2) ? 1 : (grp == 3) ? 3 : 4;
    const int64_t base = (grp == 0) ? 0 : (grp == 1) ? b_sh : (grp == 2) ? b_op : (grp == 3) ? b_sc : b_ro;
    float* __restrict__ d = dense.g[grp];
    for (int64_t e = t0; e < m * w; e += stride) {
        const int64_t j = e / w;
        if (j >= m_real) {
            if (kGather) compact[base + e] = 0.f;
            continue;
        }
        const int64_t src = (int64_t)idx[j] * w + (e - j * w);
        if (kGather) compact[base + e] = d[src];
        else d[src] = compact[base + e];
    }
}

// idx[offsets[i]] = i for touched rows below the capacity; status = {M, M > capacity}
__global__ void __launch_bounds__(256)
k_build_touch_index_cap(int64_t n, const int* __restrict__ touch, const int* __restrict__ offsets, int64_t cap,
                        const int64_t* __restrict__ m_dev, int* __restrict__ idx, int64_t* __restrict__ status) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && status != nullptr) {
        status[0] = *m_dev;
        status[1] = *m_dev > cap ? 1 : 0;
    }
    if (i < n && touch[i] != 0 && offsets[i] < cap) idx[offsets[i]] = (int)i;
}

}  // namespace cugs

using namespace cugs;

// idx_temp: m ints of scratch (the caller passes the tail of its offsets/compact scratch); here the
// index list is rebuilt on every call into `idx`.
static int move_rows(cugs_handle_t* h, void* stream, int64_t n, int num_coeffs, const int32_t* touch,
                     const int32_t* offsets, int64_t m, float* const grads[5], float* compact, int32_t* idx,
                     bool gather, const int64_t* m_dev, int64_t* status_dev) {
    CUGS_REQUIRE(h, h != nullptr, "handle is null");
    CUGS_REQUIRE(h, n >= 0 && m >= 0 && m <= n, "bad n / m");
    CUGS_REQUIRE(h, num_coeffs >= 1 && num_coeffs <= 64, "bad num_coeffs");
    if (n == 0 || m == 0) return CUGS_OK;
    CUGS_REQUIRE(h, grads && compact && idx, "null pointer");
    CUGS_REQUIRE(h, (touch != nullptr) == (offsets != nullptr), "touch and offsets go together");
    CUGS_REQUIRE(h, touch != nullptr || !gather, "gather needs the touch mask and its scan");
    GradGroups G;
    for (int k = 0; k < 5; ++k) {
        CUGS_REQUIRE(h, grads[k] != nullptr, "null gradient group");
        G.g[k] = grads[k];
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (touch != nullptr) {  // (a scatter right after the gather of the same exchange passes touch = NULL: idx is still valid)
        if (m_dev)
            k_build_touch_index_cap<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, touch, offsets, m, m_dev, idx,
                                                                                status_dev);
        else
            k_build_touch_index<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, touch, offsets, idx);
        CUGS_LAUNCH_CHECK(h, "k_build_touch_index");
    }
    int64_t bx = (m * 3 * num_coeffs + 255) / 256;
    const int64_t cap = (int64_t)h->sm_count * 16;
    if (bx > cap) bx = cap;
    const dim3 grid((unsigned)bx, 5);
    if (gather) k_move_grad_rows<true><<<grid, 256, 0, s>>>(num_coeffs, idx, m, m_dev, G, compact);
    else k_move_grad_rows<false><<<grid, 256, 0, s>>>(num_coeffs, idx, m, m_dev, G, compact);
    CUGS_LAUNCH_CHECK(h, "k_move_grad_rows");
    return CUGS_OK;
}

extern "C" int64_t cugs_b200_compact_grad_floats(int64_t m, int num_coeffs) {
    return align4(3 * m) + align4((int64_t)3 * num_coeffs * m) + align4(m) + align4(3 * m) + align4(4 * m);
}

extern "C" int cugs_b200_gather_grad_rows(cugs_handle_t* h, void* stream, int64_t n, int num_coeffs,
                                          const int32_t* touch, const int32_t* offsets, int64_t m,
                                          const float* const grads[5], float* compact, int32_t* idx_scratch,
                                          const int64_t* m_dev, int64_t* status_dev) {
    return move_rows(h, stream, n, num_coeffs, touch, offsets, m, const_cast<float* const*>(grads), compact,
                     idx_scratch, true, m_dev, status_dev);
}

extern "C" int cugs_b200_scatter_grad_rows(cugs_handle_t* h, void* stream, int64_t n, int num_coeffs,
                                           const int32_t* touch, const int32_t* offsets, int64_t m,
                                           const float* compact, float* const grads[5], int32_t* idx_scratch,
                                           const int64_t* m_dev) {
    return move_rows(h, stream, n, num_coeffs, touch, offsets, m, grads, const_cast<float*>(compact), idx_scratch,
                     false, m_dev, nullptr);
}

// ================================================================================================
// Peer-to-peer exchange over NVLink / NVSwitch (no NCCL, no compaction buffers).
//
// The gradient arena, the [touch mask | max_radii] buffer and the additive statistics of every rank live in
// SYMMETRIC memory (same layout on every GPU, each GPU maps all the others': torch symmetric memory = CUDA
// IPC / fabric handles), so a kernel can sum a row across the ranks directly: the rank that OWNS a slice of the
// work reads the row from all `world` arenas (independent 16-byte peer loads, all in flight together), adds them
// in rank order and writes the sum back into all `world` arenas. That is a reduce-scatter and an all-gather in one
// pass with the bytes of ONE all-reduce (2 (R-1)/R of the data cross NVLink per GPU), it needs no gather into a
// compact buffer before and no scatter after (0.28 ms of local copies at 8 GPUs), the number of touched rows M
// never has to reach the host (the slice bounds are computed on the device), and every rank receives the SAME
// bits because each sum is computed exactly once. The caller brackets the kernels with the symmetric-memory
// barrier (all writes of the previous phase visible, nobody still reading what the next phase overwrites).
// ================================================================================================
namespace cugs {

constexpr int kMaxPeers = 8;  // one NVSwitch box

struct PeerInts { int32_t* p[kMaxPeers]; };
struct PeerFloats { float* p[kMaxPeers]; };
struct PeerGroups { float* g[kMaxPeers][5]; };

// phase 1: [mask | max_radii bits] MAX (non-negative floats order like their bit patterns) and the two additive
// statistics SUM, element-wise over the rank's slice of [0, n)
__global__ void __launch_bounds__(256)
k_xchg_masks(int64_t n, int world, int rank, PeerInts maxbuf, PeerFloats accum, PeerFloats count, bool with_stats) {
    const int64_t per = (n + world - 1) / world;
    const int64_t i0 = (int64_t)rank * per, i1 = min(n, i0 + per);
    for (int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (int64_t)gridDim.x * blockDim.x) {
        int m = 0, r = 0;
        float a = 0.f, c = 0.f;
        for (int p = 0; p < world; ++p) {
            m = max(m, maxbuf.p[p][i]);
            r = max(r, maxbuf.p[p][n + i]);
            if (with_stats) { a += accum.p[p][i]; c += count.p[p][i]; }
        }
        for (int p = 0; p < world; ++p) {
            maxbuf.p[p][i] = m;
            maxbuf.p[p][n + i] = r;
            if (with_stats) { accum.p[p][i] = a; count.p[p][i] = c; }
        }
    }
}

// phase 2: the rows of the union mask. idx[j] = Gaussian of union row j (identical on every rank), M on the device.
// blockIdx.y = parameter group, as in k_move_grad_rows; the rank owns rows [rank M / R, (rank + 1) M / R).
// kWorld is a template parameter so that the R peer loads of an element are R independent instructions issued back
// to back (no predication, exact register arrays), and every thread keeps kUnroll elements in flight: the first
// version (run-time world, one element per thread) moved 63 MB each way in 0.31 ms at 2 GPUs (~200 GB/s per
// direction), latency bound on the remote reads.
template <int kWorld, int kUnroll>
__global__ void __launch_bounds__(256)
k_xchg_rows(int C, const int* __restrict__ idx, const int64_t* __restrict__ m_dev, int rank, PeerGroups peers) {
    const int grp = blockIdx.y;
    const int64_t m = *m_dev;
    const int64_t per = (m + kWorld - 1) / kWorld;
    const int64_t j0 = (int64_t)rank * per, j1 = min(m, j0 + per);
    if (j0 >= j1) return;
    const int shw = 3 * C;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int w = (grp == 0) ? 3 : (grp == 1) ? shw : (grp == 2) ? 1 : (grp == 3) ? 3 : 4;
    if ((w & 3) == 0) {  // SH (3C % 4 == 0) and rotations: 16-byte accesses
        const unsigned w4 = (unsigned)(w >> 2);
        const int64_t total = (j1 - j0) * w4;
        for (int64_t e0 = t0; e0 < total; e0 += stride * kUnroll) {
            int64_t off[kUnroll];
            float4 v[kUnroll][kWorld];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int64_t e = e0 + (int64_t)u * stride;
                off[u] = -1;
                if (e < total) {
                    const unsigned jj = (unsigned)((uint64_t)e / w4);   // (e < 2^32 for any model that fits a GPU)
                    off[u] = (int64_t)idx[j0 + jj] * w4 + (e - (int64_t)jj * w4);
                }
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u)
#pragma unroll
                for (int p = 0; p < kWorld; ++p)
                    if (off[u] >= 0) v[u][p] = reinterpret_cast<const float4*>(peers.g[p][grp])[off[u]];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                if (off[u] < 0) continue;
                float4 s = v[u][0];
#pragma unroll
                for (int p = 1; p < kWorld; ++p) { s.x += v[u][p].x; s.y += v[u][p].y; s.z += v[u][p].z; s.w += v[u][p].w; }
#pragma unroll
                for (int p = 0; p < kWorld; ++p) reinterpret_cast<float4*>(peers.g[p][grp])[off[u]] = s;
            }
        }
    } else {
        const int64_t total = (j1 - j0) * w;
        for (int64_t e = t0; e < total; e += stride) {
            const int64_t jj = e / w;
            const int64_t off = (int64_t)idx[j0 + jj] * w + (e - jj * w);
            float v[kWorld];
#pragma unroll
            for (int p = 0; p < kWorld; ++p) v[p] = peers.g[p][grp][off];
            float s = v[0];
#pragma unroll
            for (int p = 1; p < kWorld; ++p) s += v[p];
#pragma unroll
            for (int p = 0; p < kWorld; ++p) peers.g[p][grp][off] = s;
        }
    }
}

// ---- NVLS variant: the same two phases on the MULTICAST mapping of the symmetric buffers. multimem.ld_reduce
// lets the NVSwitch fetch the element from every GPU, reduce it in the switch and return ONE value; multimem.st
// stores once and the switch replicates it to every GPU. Per GPU and direction that is (R-1)/R + 1/R of the data
// on the links instead of 2 (R-1)/R for unicast loads + stores: measured (profiles/r02/NOTES.md) the unicast
// kernel is bound by NVLink at ~200-300 GB/s per direction on this pod, like NCCL's ring.
struct McGroups { float* g[5]; };

__device__ __forceinline__ float4 mc_ld_add_f32x4(const float* addr) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st_f32x4(float* addr, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float mc_ld_add_f32(const float* addr) {
    float v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f32 %0, [%1];" : "=f"(v) : "l"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st_f32(float* addr, float v) {
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" :: "l"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ int mc_ld_max_s32(const int* addr) {
    int v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.max.s32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st_s32(int* addr, int v) {
    asm volatile("multimem.st.relaxed.sys.global.s32 [%0], %1;" :: "l"(addr), "r"(v) : "memory");
}

__global__ void __launch_bounds__(256)
k_xchg_masks_mc(int64_t n, int world, int rank, int* __restrict__ maxbuf_mc, float* __restrict__ accum_mc,
                float* __restrict__ count_mc, bool with_stats) {
    const int64_t per = (n + world - 1) / world;
    const int64_t i0 = (int64_t)rank * per, i1 = min(n, i0 + per);
    for (int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (int64_t)gridDim.x * blockDim.x) {
        const int m = mc_ld_max_s32(maxbuf_mc + i);
        const int r = mc_ld_max_s32(maxbuf_mc + n + i);
        mc_st_s32(maxbuf_mc + i, m);
        mc_st_s32(maxbuf_mc + n + i, r);
        if (with_stats) {
            const float a = mc_ld_add_f32(accum_mc + i), c = mc_ld_add_f32(count_mc + i);
            mc_st_f32(accum_mc + i, a);
            mc_st_f32(count_mc + i, c);
        }
    }
}

__global__ void __launch_bounds__(256)
k_xchg_rows_mc(int C, const int* __restrict__ idx, const int64_t* __restrict__ m_dev, int world, int rank, McGroups mc) {
    const int grp = blockIdx.y;
    const int64_t m = *m_dev;
    const int64_t per = (m + world - 1) / world;
    const int64_t j0 = (int64_t)rank * per, j1 = min(m, j0 + per);
    if (j0 >= j1) return;
    const int shw = 3 * C;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int w = (grp == 0) ? 3 : (grp == 1) ? shw : (grp == 2) ? 1 : (grp == 3) ? 3 : 4;
    float* base = mc.g[grp];
    if ((w & 3) == 0) {
        const unsigned w4 = (unsigned)(w >> 2);
        const int64_t total = (j1 - j0) * w4;
        for (int64_t e0 = t0; e0 < total; e0 += stride * 4) {
            int64_t off[4];
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t e = e0 + (int64_t)u * stride;
                off[u] = -1;
                if (e < total) {
                    const unsigned jj = (unsigned)((uint64_t)e / w4);
                    off[u] = ((int64_t)idx[j0 + jj] * w4 + (e - (int64_t)jj * w4)) * 4;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (off[u] >= 0) v[u] = mc_ld_add_f32x4(base + off[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (off[u] >= 0) mc_st_f32x4(base + off[u], v[u]);
        }
    } else {
        const int64_t total = (j1 - j0) * w;
        for (int64_t e = t0; e < total; e += stride) {
            const int64_t jj = e / w;
            const int64_t off = (int64_t)idx[j0 + jj] * w + (e - jj * w);
            mc_st_f32(base + off, mc_ld_add_f32(base + off));
        }
    }
}

}  // namespace cugs

extern "C" int cugs_b200_build_touch_index(cugs_handle_t* h, void* stream, int64_t n, const int32_t* touch,
                                           const int32_t* offsets, int32_t* idx) {
    CUGS_REQUIRE(h, h != nullptr, "handle is null");
    CUGS_REQUIRE(h, n >= 0, "n must be >= 0");
    if (n == 0) return CUGS_OK;
    CUGS_REQUIRE(h, touch && offsets && idx, "null pointer");
    k_build_touch_index<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, touch, offsets, idx);
    CUGS_LAUNCH_CHECK(h, "k_build_touch_index");
    return CUGS_OK;
}

extern "C" int cugs_b200_p2p_reduce_masks(cugs_handle_t* h, void* stream, int64_t n, int world, int rank,
                                          int32_t* const* max_buf_peers, float* const* grad_accum_peers,
                                          float* const* grad_count_peers, int32_t* max_buf_mc, float* grad_accum_mc,
                                          float* grad_count_mc) {
    CUGS_REQUIRE(h, h != nullptr, "handle is null");
    CUGS_REQUIRE(h, n >= 0 && world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "bad n / world / rank");
    if (n == 0) return CUGS_OK;
    if (max_buf_mc != nullptr) {  // NVLS: multicast mapping
        const bool stats = grad_accum_mc != nullptr;
        CUGS_REQUIRE(h, stats == (grad_count_mc != nullptr), "grad_accum and grad_count go together");
        const int64_t per_mc = (n + world - 1) / world;
        int64_t blocks_mc = (per_mc + 255) / 256;
        if (blocks_mc > (int64_t)h->sm_count * 8) blocks_mc = (int64_t)h->sm_count * 8;
        k_xchg_masks_mc<<<(unsigned)blocks_mc, 256, 0, (cudaStream_t)stream>>>(n, world, rank, max_buf_mc, grad_accum_mc,
                                                                               grad_count_mc, stats);
        CUGS_LAUNCH_CHECK(h, "k_xchg_masks_mc");
        return CUGS_OK;
    }
    CUGS_REQUIRE(h, max_buf_peers != nullptr, "null pointer");
    const bool with_stats = grad_accum_peers != nullptr;
    CUGS_REQUIRE(h, with_stats == (grad_count_peers != nullptr), "grad_accum and grad_count go together");
    PeerInts mb{};
    PeerFloats ga{}, gc{};
    for (int p = 0; p < world; ++p) {
        CUGS_REQUIRE(h, max_buf_peers[p] != nullptr, "null peer pointer");
        mb.p[p] = max_buf_peers[p];
        if (with_stats) {
            CUGS_REQUIRE(h, grad_accum_peers[p] && grad_count_peers[p], "null peer pointer");
            ga.p[p] = grad_accum_peers[p];
            gc.p[p] = grad_count_peers[p];
        }
    }
    const int64_t per = (n + world - 1) / world;
    int64_t blocks = (per + 255) / 256;
    if (blocks > (int64_t)h->sm_count * 8) blocks = (int64_t)h->sm_count * 8;
    k_xchg_masks<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(n, world, rank, mb, ga, gc, with_stats);
    CUGS_LAUNCH_CHECK(h, "k_xchg_masks");
    return CUGS_OK;
}

extern "C" int cugs_b200_p2p_reduce_rows(cugs_handle_t* h, void* stream, int64_t n, int num_coeffs, int world, int rank,
                                         const int32_t* idx, const int64_t* m_dev, float* const* grads_peers,
                                         float* const* grads_mc) {
    CUGS_REQUIRE(h, h != nullptr, "handle is null");
    CUGS_REQUIRE(h, n >= 0 && world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "bad n / world / rank");
    CUGS_REQUIRE(h, num_coeffs >= 1 && num_coeffs <= 64, "bad num_coeffs");
    if (n == 0) return CUGS_OK;
    CUGS_REQUIRE(h, idx && m_dev && (grads_peers || grads_mc), "null pointer");
    if (grads_mc != nullptr) {  // NVLS: multicast mapping
        McGroups mc{};
        for (int k = 0; k < 5; ++k) {
            CUGS_REQUIRE(h, grads_mc[k] != nullptr, "null multicast gradient pointer");
            mc.g[k] = grads_mc[k];
        }
        int64_t bxm = ((n / world + 1) * (3 * num_coeffs / 4 + 1) / 4 + 255) / 256;
        const int64_t capm = (int64_t)h->sm_count * 8;
        if (bxm > capm) bxm = capm;
        if (bxm < 1) bxm = 1;
        k_xchg_rows_mc<<<dim3((unsigned)bxm, 5), 256, 0, (cudaStream_t)stream>>>(num_coeffs, idx, m_dev, world, rank, mc);
        CUGS_LAUNCH_CHECK(h, "k_xchg_rows_mc");
        return CUGS_OK;
    }
    PeerGroups pg{};
    for (int p = 0; p < world; ++p)
        for (int k = 0; k < 5; ++k) {
            CUGS_REQUIRE(h, grads_peers[p * 5 + k] != nullptr, "null peer gradient pointer");
            pg.g[p][k] = grads_peers[p * 5 + k];
        }
    // sized for the worst case (every row touched) of this rank's slice; the kernel reads M on the device
    int64_t bx = ((n / world + 1) * (3 * num_coeffs / 4 + 1) + 255) / 256;
    const int64_t cap = (int64_t)h->sm_count * 8;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    const dim3 grid((unsigned)bx, 5);
    cudaStream_t s = (cudaStream_t)stream;
    switch (world) {
        case 1: k_xchg_rows<1, 4><<<grid, 256, 0, s>>>(num_coeffs, idx, m_dev, rank, pg); break;
        case 2: k_xchg_rows<2, 4><<<grid, 256, 0, s>>>(num_coeffs, idx, m_dev, rank, pg); break;
        case 3: k_xchg_rows<3, 2><<<grid, 256, 0, s>>>(num_coeffs, idx, m_dev, rank, pg); break;
        case 4: k_xchg_rows<4, 2><<<grid, 256, 0, s>>>(num_coeffs, idx, m_dev, rank, pg); break;
        case 5: k_xchg_rows<5, 2><<<grid, 256, 0, s>>>(num_coeffs, idx, m_dev, rank, pg); break;
        case 6: k_xchg_rows<6, 2><<<grid, 256, 0, s>>>(num_coeffs, idx, m_dev, rank, pg); break;
        case 7: k_xchg_rows<7, 2><<<grid, 256, 0, s>>>(num_coeffs, idx, m_dev, rank, pg); break;
        default: k_xchg_rows<8, 2><<<grid, 256, 0, s>>>(num_coeffs, idx, m_dev, rank, pg); break;
    }
    CUGS_LAUNCH_CHECK(h, "k_xchg_rows");
    return CUGS_OK;
}
