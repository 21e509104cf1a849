// blend.cu — per-16x16-tile alpha blending, forward and backward, for sm_100a.
//
// Forward  replaces k_rasterize_forward  (reference rasterizer/forward.cu:48-174).
// Backward replaces k_rasterize_backward (reference rasterizer/backward.cu:31-233).
//
// Both kernels are instruction-issue bound (ncu: >80 % issue-active, <5 % DRAM), so the design
// minimises instructions per (pixel, Gaussian) evaluation:
//  * one CTA of 64 threads (2 warps) per tile; a thread owns FOUR pixels of one row (columns c, c+4,
//    c+8, c+12), a warp a 16x8 patch. The Gaussian record is read from shared memory once per
//    thread (2 broadcast LDS.128) for four evaluations, the dy-terms of `power` are shared by the
//    four pixels, `power` is evaluated two pixels per instruction with packed FP32x2 arithmetic
//    (FADD2/FMUL2/FFMA2, bit-identical to the scalar sequence), and the four independent
//    dependency chains give the ILP a 2-warp CTA needs;
//  * Gaussians are staged in batches of 128 through a 2-deep shared-memory ring of 48-byte packed
//    records {x,y,a,b | c,thr,op,r | g,b,hx,hy} (written by preprocess) with 16-byte cp.async
//    (LDGSTS) copies; the gather of batch k+1 overlaps the blending of batch k;
//  * `power` is computed with the reference's exact rounding and compared against a per-Gaussian
//    conservative bound thr = -log(255*op) - 1e-4: the accurate expf (11 instructions + MUFU) only
//    runs for evaluations that can reach alpha >= 1/255. Those take the exact path (same expf,
//    same comparisons, same rounding of T as the reference), so n_contrib / final_T are
//    bit-identical to the reference;
//  * per-warp list compaction: at the start of a batch every warp tests the 128 staged Gaussians
//    against ITS 16x8 pixel patch (conservative footprint box {hx,hy} carried in the record) with
//    four ballots and keeps an ordered list of the survivors in shared memory; the inner loop
//    only visits those (44 % of the (patch, Gaussian) pairs are dropped on the synthetic scenes)
//    — dropping a Gaussian that cannot pass the cheap reject anywhere in the patch changes nothing;
//  * early termination: per-warp vote every 8 Gaussians, per-CTA __syncthreads_and per batch;
//  * backward: a thread first sums the nine per-Gaussian gradient terms over its own four pixels
//    in registers, then the warp reduces them with a multi-value butterfly (8 values in 7
//    shuffle steps + 1 value in 5, instead of 9 x 5), and 9 lanes issue ONE coalesced
//    red.global.add.f32 to the 9 consecutive floats of the Gaussian's packed gradient record —
//    instead of the reference's nine scalar atomics per (pixel, Gaussian);
//  * backward, per contributing (pixel, Gaussian): the colour accumulated behind the Gaussian is carried as ONE
//    scalar per pixel, G = sum_c dL/dC_c * S_c (dL/dC is constant along the walk), and the mean / conic gradients
//    come from three per-thread moments (sum dLp, sum dLp dx, sum dLp dx^2; dy is common to a thread's pixels):
//    40 instructions per pixel slot instead of 55, 76 registers without spills instead of 80 with
//    (profiles/r02/NOTES.md: 0.909 -> 0.761 ms at workload B).
#include "common.cuh"

namespace cugs {

constexpr int kBlendThreads = 64;
#ifndef CUGS_FWD_MINBLOCKS
#define CUGS_FWD_MINBLOCKS 16
#endif
#ifndef CUGS_BWD_MINBLOCKS
#define CUGS_BWD_MINBLOCKS 12
#endif
constexpr int kPix = 4;      // pixels per thread (one row segment)
// Staging granularity. Default: the CTA's two warps share one ring of 128-Gaussian batches (one barrier per batch).
// -DCUGS_BLEND_WP ("warp private"): every warp stages its OWN ring of 64-Gaussian batches and walks the tile's
// list on its own -- no block barrier anywhere, a warp whose 16x8 patch is finished leaves without waiting for its
// sibling -- at the price of staging every record twice (L2 -> shared, the 48-byte records are L2 resident).
#ifdef CUGS_BLEND_WP
constexpr int kBatch = 64;
#define CUGS_ST_T (threadIdx.x & 31)          /* staging slot of the thread inside the batch */
#define CUGS_ST_N 32                          /* threads that stage one batch */
#define CUGS_SG(buf) s_g[threadIdx.x >> 5][buf]
#define CUGS_SID(buf) s_idx[threadIdx.x >> 5][buf]
#define CUGS_BATCH_SYNC() __syncwarp()
#define CUGS_BATCH_DONE(pred) __all_sync(kFull, (pred))
#define CUGS_RING_DIMS [2][2][kBatch]
#else
constexpr int kBatch = 128;  // Gaussians per staged batch (two per thread)
#define CUGS_ST_T threadIdx.x
#define CUGS_ST_N kBlendThreads
#define CUGS_SG(buf) s_g[buf]
#define CUGS_SID(buf) s_idx[buf]
#define CUGS_BATCH_SYNC() __syncthreads()
#define CUGS_BATCH_DONE(pred) __syncthreads_and(pred)
#define CUGS_RING_DIMS [2][kBatch]
#endif
constexpr float kAlphaMin = 1.0f / 255.0f;
constexpr float kTMin = 1.0f / 255.0f;  // forward.cuh:25-31 kTransmittanceThreshold

struct __align__(16) StagedGaussian {
    float4 q0;  // x, y, a, b
    float4 q1;  // c, thr, opacity, r
    float4 q2;  // g, b, hx, hy (conservative footprint half-extents, common.cuh blend_extents)
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ float rcp_fast(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// ---- experiment (-DCUGS_BLEND_TMA): stage each 48-byte record with ONE 1-D TMA bulk copy
// (cp.async.bulk.shared::cluster.global, completion counted in bytes on an mbarrier) instead of three 16-byte
// LDGSTS. Measured on B200 (profiles/r02/NOTES.md): no gain -- the staging is one batch ahead of its use and
// costs ~1 % of the kernels' instructions either way -- so the default build keeps LDGSTS.
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t"
        "}" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_copy48(void* smem, const void* gmem, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 48, [%2];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem)),
                 "l"(gmem), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}

// Stage one Gaussian of the batch into shared memory.
template <bool kPacked>
__device__ __forceinline__ void stage_gaussian(StagedGaussian* dst, int g, const float4* __restrict__ packed,
                                               const float* __restrict__ means_2d,
                                               const float* __restrict__ conic,
                                               const float* __restrict__ rgb,
                                               const float* __restrict__ opa) {
    if (kPacked) {
        const float4* src = packed + (int64_t)g * 3;
        cp_async16(&dst->q0, src);
        cp_async16(&dst->q1, src + 1);
        cp_async16(&dst->q2, src + 2);
    } else {
        const float2 m = reinterpret_cast<const float2*>(means_2d)[g];
        const float a = conic[(int64_t)g * 3], b = conic[(int64_t)g * 3 + 1], c = conic[(int64_t)g * 3 + 2];
        const float o = opa[g];
        const float thr = blend_reject_threshold(o);
        const float inv_det = 1.0f / (a * c - b * b);  // Sigma' = conic^-1: Sigma'_xx = c/det, Sigma'_yy = a/det
        float hx, hy;
        blend_extents(thr, c * inv_det, a * inv_det, hx, hy);
        dst->q0 = make_float4(m.x, m.y, a, b);
        dst->q1 = make_float4(c, thr, o, rgb[(int64_t)g * 3]);
        dst->q2 = make_float4(rgb[(int64_t)g * 3 + 1], rgb[(int64_t)g * 3 + 2], hx, hy);
    }
}

// thread -> its four pixels: lane = (row-in-warp << 2) | c; warp w covers rows 8w..8w+7; pixel slot
// k of the thread is column 4k + c, so slot k of a warp is the 4-column band [4k, 4k+3] x 8 rows:
// a Gaussian that does not reach a band fails the cheap reject in ALL lanes of that slot and the
// slot's exact path is skipped by a uniform branch.
constexpr int kPixStride = 4;
__device__ __forceinline__ void pixels_of_thread(int tile_x, int tile_y, int& px0, int& py) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    px0 = tile_x * kTile + (lane & 3);
    py = tile_y * kTile + warp * 8 + (lane >> 2);
}

// Ordered list of the batch's Gaussians whose footprint box can reach this warp's 16x8 patch.
__device__ __forceinline__ int build_warp_list(const void* sg_void, int bc, float x0, float y0,
                                               unsigned char* list) {
    struct Rec { float4 q0, q1, q2; };
    const Rec* sg = reinterpret_cast<const Rec*>(sg_void);
    const int lane = threadIdx.x & 31;
    const float x1 = x0 + 15.0f, y1 = y0 + 7.0f;  // pixel centres of the patch corners
    int cnt = 0;
#pragma unroll
    for (int u = 0; u < kBatch / 32; ++u) {
        const int jj = u * 32 + lane;
        bool keep = false;
        if (jj < bc) {
            const float2 xy = *reinterpret_cast<const float2*>(&sg[jj].q0);
            const float2 h = *reinterpret_cast<const float2*>(&sg[jj].q2.z);
            keep = !((x0 - xy.x > h.x) || (xy.x - x1 > h.x) || (y0 - xy.y > h.y) || (xy.y - y1 > h.y));
        }
        const unsigned m = __ballot_sync(kFull, keep);
        if (keep) list[cnt + __popc(m & ((1u << lane) - 1))] = (unsigned char)jj;
        cnt += __popc(m);
    }
    __syncwarp();
    return cnt;
}

// ---- packed FP32x2 arithmetic (sm_100: FADD2 / FMUL2 / FFMA2, IEEE round-to-nearest per element, so
// the results are the scalar ones bit for bit while the issue slots are halved) ------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpk2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// `power` of two pixels of one row at once, with the reference's roundings (see blend_power4):
//   dx = px - x; s1 = fma(dx, a, dy*b); s2 = fma(dx, b, dy*c); power = fma(dx, s1, dy*s2) * -0.5
__device__ __forceinline__ f32x2 blend_power_x2(f32x2 px, float neg_x, float dy, float a, float b, float dyb,
                                                float dyc, f32x2& dx, f32x2& s1, f32x2& s2) {
    dx = add2(px, pk2(neg_x, neg_x));  // px + (-x) == px - x exactly
    s1 = fma2(dx, pk2(a, a), pk2(dyb, dyb));
    s2 = fma2(dx, pk2(b, b), pk2(dyc, dyc));
    return mul2(fma2(dx, s1, mul2(pk2(dy, dy), s2)), pk2(-0.5f, -0.5f));
}

// The reference's `power`, rounding for rounding (forward.cu:131-132 and
// backward.cu:132-133 compile to the same sequence):
//   s1 = fma(dx, a, dy*b); s2 = fma(dx, b, dy*c); power = (fma(dx, s1, dy*s2)) * -0.5
// dyb = rn(dy*b) and dyc = rn(dy*c) are shared by the four pixels of a thread (same row).
__device__ __forceinline__ float blend_power4(float dx, float dy, float a, float b, float dyb, float dyc,
                                              float& s1, float& s2) {
    s1 = fma_rn(dx, a, dyb);
    s2 = fma_rn(dx, b, dyc);
    return mul_rn(fma_rn(dx, s1, mul_rn(dy, s2)), -0.5f);
}

// ================================================================================================
// forward
// ================================================================================================
template <bool kPacked>
__global__ void __launch_bounds__(kBlendThreads, CUGS_FWD_MINBLOCKS)
k_blend_fwd(int ntx, int width, int height, float bg_r, float bg_g, float bg_b,
            const int* __restrict__ tile_ranges, const int* __restrict__ gaussian_idx,
            const float4* __restrict__ packed, const float* __restrict__ means_2d,
            const float* __restrict__ conic, const float* __restrict__ rgb,
            const float* __restrict__ opa, float* __restrict__ out_color,
            float* __restrict__ out_T, int* __restrict__ out_n) {
    __shared__ StagedGaussian s_g CUGS_RING_DIMS;
    __shared__ unsigned char s_list[2][kBatch];
#ifdef CUGS_BLEND_TMA
    __shared__ __align__(8) uint64_t s_bar[2];
    if (kPacked) {
        if (threadIdx.x == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncthreads();
    }
#define CUGS_STAGE(dstp, gi, slot)                                                              \
    do {                                                                                        \
        if (kPacked) bulk_copy48((dstp), packed + (int64_t)(gi) * 3, &s_bar[(slot)]);           \
        else stage_gaussian<kPacked>((dstp), (gi), packed, means_2d, conic, rgb, opa);          \
    } while (0)
#define CUGS_EXPECT(slot, cnt)                                                                  \
    do {                                                                                        \
        if (kPacked && threadIdx.x == 0) mbar_expect_tx(&s_bar[(slot)], (unsigned)(cnt) * 48u); \
    } while (0)
#define CUGS_WAIT_BATCH(slot, parity, pending)                                                  \
    do {                                                                                        \
        if (kPacked) mbar_wait(&s_bar[(slot)], (parity));                                       \
        else if (pending) cp_async_wait<1>();                                                   \
        else cp_async_wait<0>();                                                                \
    } while (0)
#else
#define CUGS_STAGE(dstp, gi, slot) stage_gaussian<kPacked>((dstp), (gi), packed, means_2d, conic, rgb, opa)
#define CUGS_EXPECT(slot, cnt) do {} while (0)
#define CUGS_WAIT_BATCH(slot, parity, pending)                                                  \
    do {                                                                                        \
        if (pending) cp_async_wait<1>();                                                        \
        else cp_async_wait<0>();                                                                \
    } while (0)
#endif

    const int tile = blockIdx.x;
    const int tile_x = tile % ntx, tile_y = tile / ntx;
    int px0, py;
    pixels_of_thread(tile_x, tile_y, px0, py);
    const float pyf = (float)py + 0.5f;  // forward.cu:72-73
    const int warp = threadIdx.x >> 5;
    const float patch_x0 = (float)(tile_x * kTile) + 0.5f, patch_y0 = (float)(tile_y * kTile + warp * 8) + 0.5f;
    // lim[k] = 0 while pixel k is live, -inf once it is done (or outside the image): the reference's
    // `power > 0 -> skip` test becomes `power > lim[k]`, which also rejects everything for a done
    // pixel at no extra cost (forward.cu:135, :153)
    float pxf[kPix], lim[kPix];
#pragma unroll
    for (int k = 0; k < kPix; ++k) {
        pxf[k] = (float)(px0 + k * kPixStride) + 0.5f;
        lim[k] = ((px0 + k * kPixStride < width) && (py < height)) ? 0.0f : -INFINITY;
    }
    const f32x2 px01 = pk2(pxf[0], pxf[1]), px23 = pk2(pxf[2], pxf[3]);
#define CUGS_ALL_DONE (lim[0] < 0.0f && lim[1] < 0.0f && lim[2] < 0.0f && lim[3] < 0.0f)

    const int2 range = reinterpret_cast<const int2*>(tile_ranges)[tile];
    const int count = range.y - range.x;
    const int nb = (count + kBatch - 1) / kBatch;

    float T[kPix], C0[kPix], C1[kPix], C2[kPix];
    int contrib[kPix];
#pragma unroll
    for (int k = 0; k < kPix; ++k) { T[k] = 1.0f; C0[k] = C1[k] = C2[k] = 0.0f; contrib[k] = 0; }

    // prologue: gather batch 0, prefetch the indices of batch 1
    int next_idx[2] = {-1, -1};
    if (nb > 0) {
        CUGS_EXPECT(0, min(kBatch, count));
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int li = range.x + CUGS_ST_T + u * CUGS_ST_N;
            if (li < range.y) CUGS_STAGE(&CUGS_SG(0)[CUGS_ST_T + u * CUGS_ST_N], gaussian_idx[li], 0);
            const int li1 = li + kBatch;
            if (li1 < range.y) next_idx[u] = gaussian_idx[li1];
        }
        cp_async_commit();
    }

    for (int b = 0; b < nb; ++b) {
        // issue the gather of batch b+1 into the other buffer (its previous reader, batch b-1,
        // finished before the __syncthreads_and at the end of the previous iteration)
        if (b + 1 < nb) {
            CUGS_EXPECT((b + 1) & 1, min(kBatch, count - (b + 1) * kBatch));
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (next_idx[u] >= 0) CUGS_STAGE(&CUGS_SG((b + 1) & 1)[CUGS_ST_T + u * CUGS_ST_N], next_idx[u], (b + 1) & 1);
                const int li2 = range.x + (b + 2) * kBatch + CUGS_ST_T + u * CUGS_ST_N;
                next_idx[u] = (li2 < range.y) ? gaussian_idx[li2] : -1;
            }
            cp_async_commit();
        }
        CUGS_WAIT_BATCH(b & 1, (b >> 1) & 1, b + 1 < nb);
        CUGS_BATCH_SYNC();

        const StagedGaussian* sg = CUGS_SG(b & 1);
        const int bc = min(kBatch, count - b * kBatch);
        const unsigned char* list = s_list[warp];
        const int cnt = build_warp_list(sg, bc, patch_x0, patch_y0, s_list[warp]);
        for (int i0 = 0; i0 < cnt; i0 += 8) {
            if (__all_sync(kFull, CUGS_ALL_DONE)) break;
            const int i1 = min(i0 + 8, cnt);
            for (int i = i0; i < i1; ++i) {
                const int j = list[i];
                const float4 q0 = sg[j].q0;
                const float2 ct = *reinterpret_cast<const float2*>(&sg[j].q1);
                const float dy = pyf - q0.y;
                const float dyb = mul_rn(dy, q0.w), dyc = mul_rn(dy, ct.x);
                float power[kPix];
                bool pass[kPix];
                {
                    f32x2 dxa, s1a, s2a, dxb, s1b, s2b;
                    const f32x2 pa = blend_power_x2(px01, -q0.x, dy, q0.z, q0.w, dyb, dyc, dxa, s1a, s2a);
                    const f32x2 pb = blend_power_x2(px23, -q0.x, dy, q0.z, q0.w, dyb, dyc, dxb, s1b, s2b);
                    unpk2(pa, power[0], power[1]);
                    unpk2(pb, power[2], power[3]);
                }
                bool any = false;
#pragma unroll
                for (int k = 0; k < kPix; ++k) {
                    // cheap reject: cannot reach alpha >= 1/255 (NaNs fall through to the exact path)
                    pass[k] = !(power[k] < ct.y || power[k] > lim[k]);
                    any |= pass[k];
                }
                if (!any) continue;
                const float2 orr = *reinterpret_cast<const float2*>(&sg[j].q1.z);  // opacity, r
                const float2 gb = *reinterpret_cast<const float2*>(&sg[j].q2);
#pragma unroll
                for (int k = 0; k < kPix; ++k) {
                    if (!pass[k]) continue;
                    const float alpha = fminf(mul_rn(orr.x, expf(power[k])), 0.99f);  // forward.cu:137-140
                    if (alpha < kAlphaMin) continue;
                    const float w = mul_rn(alpha, T[k]);
                    C0[k] = fma_rn(w, orr.y, C0[k]);
                    C1[k] = fma_rn(w, gb.x, C1[k]);
                    C2[k] = fma_rn(w, gb.y, C2[k]);
                    T[k] = mul_rn(T[k], add_rn(1.0f, -alpha));
                    ++contrib[k];
                    if (T[k] < kTMin) lim[k] = -INFINITY;  // the crossing Gaussian is composited (forward.cu:153)
                }
            }
        }
        if (CUGS_BATCH_DONE(CUGS_ALL_DONE)) break;
    }
    cp_async_wait<0>();
#undef CUGS_ALL_DONE

    if (py < height) {
#pragma unroll
        for (int k = 0; k < kPix; ++k) {
            if (px0 + k * kPixStride >= width) continue;
            const int64_t pi = (int64_t)py * width + px0 + k * kPixStride;
            out_color[pi * 3 + 0] = fma_rn(T[k], bg_r, C0[k]);  // forward.cu:166-168
            out_color[pi * 3 + 1] = fma_rn(T[k], bg_g, C1[k]);
            out_color[pi * 3 + 2] = fma_rn(T[k], bg_b, C2[k]);
            out_T[pi] = T[k];
            out_n[pi] = contrib[k];
        }
    }
}

// ================================================================================================
// backward
// ================================================================================================
// Sum v[0..8] over the 32 lanes of the warp. On return lane 4*i (i = 0..7) holds the total of
// v[i] in `r`, and every lane holds the total of v[8] in `r8`.
__device__ __forceinline__ void warp_reduce9(const float (&v)[9], int lane, float& r, float& r8) {
    // 8 values: after the xor-16 / xor-8 / xor-4 exchanges a lane keeps 4, 2, 1 partial sums
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    float w4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float keep = b4 ? v[i + 4] : v[i];
        const float send = b4 ? v[i] : v[i + 4];
        w4[i] = keep + __shfl_xor_sync(kFull, send, 16);
    }
    float w2[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float keep = b3 ? w4[i + 2] : w4[i];
        const float send = b3 ? w4[i] : w4[i + 2];
        w2[i] = keep + __shfl_xor_sync(kFull, send, 8);
    }
    {
        const float keep = b2 ? w2[1] : w2[0];
        const float send = b2 ? w2[0] : w2[1];
        r = keep + __shfl_xor_sync(kFull, send, 4);
    }
    r += __shfl_xor_sync(kFull, r, 2);
    r += __shfl_xor_sync(kFull, r, 1);
    // now lanes with (b4,b3,b2) = bits hold value index 4*b4 + 2*b3 + b2
    r8 = v[8];
    r8 += __shfl_xor_sync(kFull, r8, 16);
    r8 += __shfl_xor_sync(kFull, r8, 8);
    r8 += __shfl_xor_sync(kFull, r8, 4);
    r8 += __shfl_xor_sync(kFull, r8, 2);
    r8 += __shfl_xor_sync(kFull, r8, 1);
}

template <bool kPacked>
__global__ void __launch_bounds__(kBlendThreads, CUGS_BWD_MINBLOCKS)
k_blend_bwd(int ntx, int width, int height, float bg_r, float bg_g, float bg_b,
            const int* __restrict__ tile_ranges, const int* __restrict__ gaussian_idx,
            const float4* __restrict__ packed, const float* __restrict__ means_2d,
            const float* __restrict__ conic, const float* __restrict__ rgb,
            const float* __restrict__ opa, const float* __restrict__ dL_dcolor,
            const float* __restrict__ final_T, const int* __restrict__ n_contrib,
            float* __restrict__ grad_acc /* [N,12] */) {
    __shared__ StagedGaussian s_g CUGS_RING_DIMS;
    __shared__ int s_idx CUGS_RING_DIMS;
    __shared__ unsigned char s_list[2][kBatch];
#ifdef CUGS_BLEND_TMA
    __shared__ __align__(8) uint64_t s_bar[2];
    if (kPacked) {
        if (threadIdx.x == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncthreads();
    }
#endif

    const int tile = blockIdx.x;
    const int tile_x = tile % ntx, tile_y = tile / ntx;
    int px0, py;
    pixels_of_thread(tile_x, tile_y, px0, py);
    const float pyf = (float)py + 0.5f;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float patch_x0 = (float)(tile_x * kTile) + 0.5f, patch_y0 = (float)(tile_y * kTile + warp * 8) + 0.5f;
    // after warp_reduce9 lanes 0,4,...,28 hold v0..v7 and lane 1 takes v8: nine lanes, nine consecutive floats
    const int red_slot = (lane == 1) ? 8
                         : ((lane & 3) == 0 ? (((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)) : -1);

    const int2 range = reinterpret_cast<const int2*>(tile_ranges)[tile];
    const int count = range.y - range.x;
    const int nb = (count + kBatch - 1) / kBatch;

    // per-pixel forward outputs (backward.cu:65-87)
    float pxf[kPix], T[kPix], g0[kPix], g1[kPix], g2[kPix];
    float G[kPix];    // sum_c dL/dC_c * (colour accumulated behind the current Gaussian, background included)
    float lim[kPix];  // 0 while the pixel still has contributors to process, -inf afterwards (see forward)
    int left[kPix];   // contributors still to process; <= 0 means the pixel is done
#pragma unroll
    for (int k = 0; k < kPix; ++k) {
        pxf[k] = (float)(px0 + k * kPixStride) + 0.5f;
        T[k] = 0.0f; g0[k] = g1[k] = g2[k] = 0.0f; left[k] = 0;
        if (px0 + k * kPixStride < width && py < height) {
            const int64_t pi = (int64_t)py * width + px0 + k * kPixStride;
            T[k] = final_T[pi];
            left[k] = n_contrib[pi];
            g0[k] = dL_dcolor[pi * 3 + 0];
            g1[k] = dL_dcolor[pi * 3 + 1];
            g2[k] = dL_dcolor[pi * 3 + 2];
        }
        G[k] = T[k] * (g0[k] * bg_r + g1[k] * bg_g + g2[k] * bg_b);
        lim[k] = (left[k] > 0) ? 0.0f : -INFINITY;
    }
    const f32x2 px01 = pk2(pxf[0], pxf[1]), px23 = pk2(pxf[2], pxf[3]);
#define CUGS_ALL_DONE (lim[0] < 0.0f && lim[1] < 0.0f && lim[2] < 0.0f && lim[3] < 0.0f)

    // batches are visited last to first; batch b covers [range.x + b*kBatch, ...)
    int next_idx[2] = {-1, -1};
    if (nb > 0) {
        CUGS_EXPECT((nb - 1) & 1, count - (nb - 1) * kBatch);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int t = CUGS_ST_T + u * CUGS_ST_N;
            const int li = range.x + (nb - 1) * kBatch + t;
            if (li < range.y) {
                const int g = gaussian_idx[li];
                CUGS_SID((nb - 1) & 1)[t] = g;
                CUGS_STAGE(&CUGS_SG((nb - 1) & 1)[t], g, (nb - 1) & 1);
            }
            if (nb > 1) next_idx[u] = gaussian_idx[li - kBatch];  // batch nb-2 is always full
        }
        cp_async_commit();
    }

    for (int b = nb - 1; b >= 0; --b) {
        if (b > 0) {
            CUGS_EXPECT((b - 1) & 1, kBatch);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int t = CUGS_ST_T + u * CUGS_ST_N;
                CUGS_SID((b - 1) & 1)[t] = next_idx[u];
                CUGS_STAGE(&CUGS_SG((b - 1) & 1)[t], next_idx[u], (b - 1) & 1);
                if (b > 1) next_idx[u] = gaussian_idx[range.x + (b - 2) * kBatch + t];
            }
            cp_async_commit();
        }
        CUGS_WAIT_BATCH(b & 1, ((nb - 1 - b) >> 1) & 1, b > 0);
        CUGS_BATCH_SYNC();

        const StagedGaussian* sg = CUGS_SG(b & 1);
        const int* sid = CUGS_SID(b & 1);
        const int bc = min(kBatch, count - b * kBatch);
        const unsigned char* list = s_list[warp];
        const int cnt = build_warp_list(sg, bc, patch_x0, patch_y0, s_list[warp]);
        for (int i = cnt - 1; i >= 0; --i) {
            if ((i & 7) == 7 && __all_sync(kFull, CUGS_ALL_DONE)) break;
            const int j = list[i];
            const float4 q0 = sg[j].q0;
            const float4 q1 = sg[j].q1;
            const float a = q0.z, bq = q0.w, c = q1.x;
            const float dy = pyf - q0.y;
            const float dyb = mul_rn(dy, bq), dyc = mul_rn(dy, c);
            float power[kPix], dx[kPix];
            bool pass[kPix];
            {
                f32x2 dxa, s1a, s2a, dxb, s1b, s2b;
                const f32x2 pa = blend_power_x2(px01, -q0.x, dy, a, bq, dyb, dyc, dxa, s1a, s2a);
                const f32x2 pb = blend_power_x2(px23, -q0.x, dy, a, bq, dyb, dyc, dxb, s1b, s2b);
                unpk2(pa, power[0], power[1]); unpk2(pb, power[2], power[3]);
                unpk2(dxa, dx[0], dx[1]); unpk2(dxb, dx[2], dx[3]);
            }
            bool any = false;
#pragma unroll
            for (int k = 0; k < kPix; ++k) {
                pass[k] = !(power[k] < q1.y || power[k] > lim[k]);
                any |= pass[k];
            }
            if (!__any_sync(kFull, any)) continue;

            float v[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) v[q] = 0.0f;
            if (any) {
                const float2 gb = *reinterpret_cast<const float2*>(&sg[j].q2);
                const float cr = q1.w, cg = gb.x, cb = gb.y;
                float m_dx = 0.f, m_xx = 0.f;  // sum dLp*dx, sum dLp*dx*dx ; dy is common to the 4 pixels
                float m_p = 0.f;               // sum dLp
#pragma unroll
                for (int k = 0; k < kPix; ++k) {
                    if (!pass[k]) continue;
                    // the forward's own expf: same alpha >= 1/255 and alpha-clamp decisions bit for bit. (Round 1
                    // used ex2.approx plus a guard band around the two thresholds; the band's two compares and
                    // divergent branch cost more issue slots than expf's range reduction: 0.909 -> 0.858 ms.)
                    const float ex = expf(power[k]);
                    const float oe = mul_rn(q1.z, ex);
                    const float alpha = fminf(oe, 0.99f);
                    if (alpha < kAlphaMin) continue;       // backward.cu:137-139
                    if (--left[k] <= 0) lim[k] = -INFINITY;  // found++ ; found > n_contrib -> stop (:141-145)
                    const float oma = fmaxf(1.0f - alpha, 1e-5f);  // backward.cu:150-151
                    const float inv = rcp_fast(oma);  // oma >= 1e-5: no denormal handling needed
                    T[k] = T[k] * inv;
                    const float w = alpha * T[k];
                    v[0] = fmaf(g0[k], w, v[0]);
                    v[1] = fmaf(g1[k], w, v[1]);
                    v[2] = fmaf(g2[k], w, v[2]);
                    // dL/dalpha = sum_c dL/dC_c * (T*rgb_c - S_c/oma)   (backward.cu:166-168)
                    const float gc = g0[k] * cr + g1[k] * cg + g2[k] * cb;
                    const bool clamped = (oe >= 0.99f);      // backward.cu:178-190
                    // The pixel's dL/dC is constant along the walk, so the colour accumulated behind the Gaussian
                    // (backward.cu's accum_rec) enters only through G = sum_c dL/dC_c * S_c: one G += w * gc replaces
                    // three S updates and a dot product, and three registers per pixel.
                    const float dLa = T[k] * gc - inv * G[k];
                    G[k] = fmaf(w, gc, G[k]);
                    const float dLa_c = clamped ? 0.0f : dLa;
                    const float dLp = dLa_c * alpha;
                    v[3] = fmaf(dLa_c, ex, v[3]);
                    const float pdx = dLp * dx[k];
                    m_p += dLp;
                    m_dx += pdx;
                    m_xx = fmaf(pdx, dx[k], m_xx);
                }
                // dL/dmean: sum dLp * (a dx + b dy) and sum dLp * (b dx + c dy), from the moments (dy is common to the
                // thread's four pixels) instead of two FMAs per pixel
                v[4] = fmaf(a, m_dx, dyb * m_p);
                v[5] = fmaf(bq, m_dx, dyc * m_p);
                v[6] = -0.5f * m_xx;            // sum dLp * (-0.5 dx^2)
                v[7] = -(m_dx * dy);            // sum dLp * (-dx dy)
                v[8] = -0.5f * (m_p * dy * dy); // sum dLp * (-0.5 dy^2)
            }
            float r, r8;
            warp_reduce9(v, lane, r, r8);
            if (red_slot >= 0) atomicAdd(grad_acc + (int64_t)sid[j] * 12 + red_slot, (red_slot == 8) ? r8 : r);
        }
        if (CUGS_BATCH_DONE(CUGS_ALL_DONE)) break;
    }
    cp_async_wait<0>();
#undef CUGS_ALL_DONE
}

// ================================================================================================
// work counter (measurement only, never inside a timed region): the number of (pixel, Gaussian)
// evaluations the REFERENCE traversal performs, split into alpha-rejected and contributing ones --
// E_fwd by forward.cu:121-157 (a pixel walks its tile's list front to back until T < 1/255), E_bwd by
// backward.cu:117-145 (back to front until more than n_contrib alpha-passing Gaussians were met).
// These are the work units of SURVEY 8(d): 15 FLOP + 1 EX2 per rejected evaluation, 24 FLOP + 1 EX2 per
// contributing forward evaluation, 55 FLOP + 1 EX2 + 1 RCP per contributing backward evaluation.
// One thread per pixel, one 256-thread CTA per tile, plain reference arithmetic.
// ================================================================================================
__global__ void __launch_bounds__(256)
k_count_evaluations(int ntx, int width, int height, const int* __restrict__ tile_ranges,
                    const int* __restrict__ gaussian_idx, const float* __restrict__ means_2d,
                    const float* __restrict__ conic, const float* __restrict__ opa,
                    const int* __restrict__ n_contrib, unsigned long long* __restrict__ counts /* [4] */) {
    __shared__ float s_x[256], s_y[256], s_a[256], s_b[256], s_c[256], s_o[256];
    __shared__ unsigned long long s_cnt[4];
    const int tile = blockIdx.x;
    const int tile_x = tile % ntx, tile_y = tile / ntx;
    const int px = tile_x * kTile + (threadIdx.x & 15), py = tile_y * kTile + (threadIdx.x >> 4);
    const bool inside = px < width && py < height;
    const float pxf = (float)px + 0.5f, pyf = (float)py + 0.5f;
    const int2 range = reinterpret_cast<const int2*>(tile_ranges)[tile];
    const int count = range.y - range.x;
    const int nb = (count + 255) / 256;
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0ull;
    unsigned long long c_rej = 0, c_con = 0, b_rej = 0, b_con = 0;

    auto stage = [&](int b) {
        const int li = range.x + b * 256 + threadIdx.x;
        if (li < range.y) {
            const int g = gaussian_idx[li];
            s_x[threadIdx.x] = means_2d[(int64_t)g * 2];
            s_y[threadIdx.x] = means_2d[(int64_t)g * 2 + 1];
            s_a[threadIdx.x] = conic[(int64_t)g * 3];
            s_b[threadIdx.x] = conic[(int64_t)g * 3 + 1];
            s_c[threadIdx.x] = conic[(int64_t)g * 3 + 2];
            s_o[threadIdx.x] = opa[g];
        }
    };
    auto alpha_of = [&](int j, bool& rejected) {
        const float dx = pxf - s_x[j], dy = pyf - s_y[j];
        float s1, s2;
        const float power = blend_power4(dx, dy, s_a[j], s_b[j], mul_rn(dy, s_b[j]), mul_rn(dy, s_c[j]), s1, s2);
        if (power > 0.0f) { rejected = true; return 0.0f; }
        const float alpha = fminf(mul_rn(s_o[j], expf(power)), 0.99f);
        rejected = alpha < kAlphaMin;
        return alpha;
    };

    // forward traversal
    float T = 1.0f;
    bool done = !inside;
    for (int b = 0; b < nb; ++b) {
        __syncthreads();
        stage(b);
        __syncthreads();
        if (!done) {
            const int bc = min(256, count - b * 256);
            for (int j = 0; j < bc; ++j) {
                bool rej;
                const float alpha = alpha_of(j, rej);
                if (rej) { ++c_rej; continue; }
                ++c_con;
                T = mul_rn(T, add_rn(1.0f, -alpha));
                if (T < kTMin) { done = true; break; }
            }
        }
    }
    // backward traversal
    const int max_contrib = inside ? n_contrib[(int64_t)py * width + px] : 0;
    int found = 0;
    done = !inside;
    for (int b = nb - 1; b >= 0; --b) {
        __syncthreads();
        stage(b);
        __syncthreads();
        if (!done) {
            const int bc = min(256, count - b * 256);
            for (int j = bc - 1; j >= 0; --j) {
                bool rej;
                (void)alpha_of(j, rej);
                if (rej) { ++b_rej; continue; }
                if (++found > max_contrib) { ++b_rej; done = true; break; }  // evaluated, then dropped (:142-145)
                ++b_con;
            }
        }
    }
    atomicAdd(&s_cnt[0], c_rej); atomicAdd(&s_cnt[1], c_con); atomicAdd(&s_cnt[2], b_rej); atomicAdd(&s_cnt[3], b_con);
    __syncthreads();
    if (threadIdx.x < 4 && s_cnt[threadIdx.x]) atomicAdd(&counts[threadIdx.x], s_cnt[threadIdx.x]);
}

#undef CUGS_STAGE
#undef CUGS_EXPECT
#undef CUGS_WAIT_BATCH

// grad_acc [N,12] -> the four public arrays of RasterizeBackwardOutput (backward.hpp:13-18)
__global__ void __launch_bounds__(256)
k_unpack_grads(int64_t n, const float4* __restrict__ acc, float* __restrict__ dL_drgb,
               float* __restrict__ dL_dopa, float* __restrict__ dL_dmean, float* __restrict__ dL_dconic) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = acc[i * 3], b = acc[i * 3 + 1], c = acc[i * 3 + 2];
    if (dL_drgb) { dL_drgb[i * 3] = a.x; dL_drgb[i * 3 + 1] = a.y; dL_drgb[i * 3 + 2] = a.z; }
    if (dL_dopa) dL_dopa[i] = a.w;
    if (dL_dmean) reinterpret_cast<float2*>(dL_dmean)[i] = make_float2(b.x, b.y);
    if (dL_dconic) { dL_dconic[i * 3] = b.z; dL_dconic[i * 3 + 1] = b.w; dL_dconic[i * 3 + 2] = c.x; }
}

}  // namespace cugs

using namespace cugs;

extern "C" int cugs_b200_blend_fwd(cugs_handle_t* h, void* stream, const cugs_view_t* v,
                                   const int32_t* tile_ranges, const int32_t* gaussian_idx,
                                   const float* means_2d, const float* cov_2d_inv, const float* rgb,
                                   const float* opacities_act, const float* packed, float* color,
                                   float* final_T, int32_t* n_contrib) {
    CUGS_REQUIRE(h, h != nullptr, "handle is null");
    CUGS_REQUIRE(h, v != nullptr && v->width > 0 && v->height > 0, "bad view");
    CUGS_REQUIRE(h, tile_ranges && color && final_T && n_contrib, "null pointer");
    CUGS_REQUIRE(h, packed || (means_2d && cov_2d_inv && rgb && opacities_act), "null Gaussian arrays");
    const int ntx = (v->width + kTile - 1) / kTile, nty = (v->height + kTile - 1) / kTile;
    cudaStream_t s = (cudaStream_t)stream;
    if (packed)
        k_blend_fwd<true><<<ntx * nty, kBlendThreads, 0, s>>>(
            ntx, v->width, v->height, v->bg[0], v->bg[1], v->bg[2], tile_ranges, gaussian_idx,
            reinterpret_cast<const float4*>(packed), means_2d, cov_2d_inv, rgb, opacities_act, color,
            final_T, n_contrib);
    else
        k_blend_fwd<false><<<ntx * nty, kBlendThreads, 0, s>>>(
            ntx, v->width, v->height, v->bg[0], v->bg[1], v->bg[2], tile_ranges, gaussian_idx, nullptr,
            means_2d, cov_2d_inv, rgb, opacities_act, color, final_T, n_contrib);
    CUGS_LAUNCH_CHECK(h, "k_blend_fwd");
    return CUGS_OK;
}

// internal: accumulate into grad_acc only (used by render_backward)
int cugs_blend_bwd_accumulate(cugs_handle_t* h, cudaStream_t s, int64_t n, const cugs_view_t* v,
                              const int32_t* tile_ranges, const int32_t* gaussian_idx,
                              const float* means_2d, const float* cov_2d_inv, const float* rgb,
                              const float* opacities_act, const float* packed, const float* dL_dcolor,
                              const float* final_T, const int32_t* n_contrib, float* grad_acc) {
    const int ntx = (v->width + kTile - 1) / kTile, nty = (v->height + kTile - 1) / kTile;
    CUGS_CUDA_TRY(h, cudaMemsetAsync(grad_acc, 0, (size_t)n * 12 * sizeof(float), s));
    if (packed)
        k_blend_bwd<true><<<ntx * nty, kBlendThreads, 0, s>>>(
            ntx, v->width, v->height, v->bg[0], v->bg[1], v->bg[2], tile_ranges, gaussian_idx,
            reinterpret_cast<const float4*>(packed), means_2d, cov_2d_inv, rgb, opacities_act, dL_dcolor,
            final_T, n_contrib, grad_acc);
    else
        k_blend_bwd<false><<<ntx * nty, kBlendThreads, 0, s>>>(
            ntx, v->width, v->height, v->bg[0], v->bg[1], v->bg[2], tile_ranges, gaussian_idx, nullptr,
            means_2d, cov_2d_inv, rgb, opacities_act, dL_dcolor, final_T, n_contrib, grad_acc);
    CUGS_LAUNCH_CHECK(h, "k_blend_bwd");
    return CUGS_OK;
}

extern "C" int cugs_b200_count_evaluations(cugs_handle_t* h, void* stream, const cugs_view_t* v,
                                           const int32_t* tile_ranges, const int32_t* gaussian_idx,
                                           const float* means_2d, const float* cov_2d_inv,
                                           const float* opacities_act, const int32_t* n_contrib,
                                           uint64_t* counts4_dev) {
    CUGS_REQUIRE(h, h != nullptr, "handle is null");
    CUGS_REQUIRE(h, v != nullptr && v->width > 0 && v->height > 0, "bad view");
    CUGS_REQUIRE(h, tile_ranges && means_2d && cov_2d_inv && opacities_act && n_contrib && counts4_dev, "null pointer");
    const int ntx = (v->width + kTile - 1) / kTile, nty = (v->height + kTile - 1) / kTile;
    cudaStream_t s = (cudaStream_t)stream;
    CUGS_CUDA_TRY(h, cudaMemsetAsync(counts4_dev, 0, 4 * sizeof(uint64_t), s));
    k_count_evaluations<<<ntx * nty, 256, 0, s>>>(ntx, v->width, v->height, tile_ranges, gaussian_idx, means_2d,
                                                  cov_2d_inv, opacities_act, n_contrib,
                                                  reinterpret_cast<unsigned long long*>(counts4_dev));
    CUGS_LAUNCH_CHECK(h, "k_count_evaluations");
    return CUGS_OK;
}

extern "C" int cugs_b200_blend_bwd(cugs_handle_t* h, void* stream, int64_t n, const cugs_view_t* v,
                                   const int32_t* tile_ranges, const int32_t* gaussian_idx,
                                   const float* means_2d, const float* cov_2d_inv, const float* rgb,
                                   const float* opacities_act, const float* packed,
                                   const float* dL_dcolor, const float* final_T,
                                   const int32_t* n_contrib, float* dL_drgb, float* dL_dopacity_act,
                                   float* dL_dmeans_2d, float* dL_dcov_2d_inv, float* grad_acc) {
    CUGS_REQUIRE(h, h != nullptr, "handle is null");
    CUGS_REQUIRE(h, v != nullptr && v->width > 0 && v->height > 0, "bad view");
    CUGS_REQUIRE(h, n >= 0, "n must be >= 0");
    if (n == 0) return CUGS_OK;
    CUGS_REQUIRE(h, tile_ranges && dL_dcolor && final_T && n_contrib && grad_acc, "null pointer");
    CUGS_REQUIRE(h, packed || (means_2d && cov_2d_inv && rgb && opacities_act), "null Gaussian arrays");
    cudaStream_t s = (cudaStream_t)stream;
    if (int e = cugs_blend_bwd_accumulate(h, s, n, v, tile_ranges, gaussian_idx, means_2d, cov_2d_inv, rgb,
                                          opacities_act, packed, dL_dcolor, final_T, n_contrib, grad_acc))
        return e;
    k_unpack_grads<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(
        n, reinterpret_cast<const float4*>(grad_acc), dL_drgb, dL_dopacity_act, dL_dmeans_2d, dL_dcov_2d_inv);
    CUGS_LAUNCH_CHECK(h, "k_unpack_grads");
    return CUGS_OK;
}
