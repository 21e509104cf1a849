// blend.cu — per-16x16-tile alpha blending, forward and backward, for sm_100a.
//
// Forward  replaces k_rasterize_forward  (reference rasterizer/forward.cu:48-174).
// Backward replaces k_rasterize_backward (reference rasterizer/backward.cu:31-233).
//
// Design (B200):
//  * one CTA (256 threads) per tile; a warp owns a compact 8x4 pixel patch (not 2 rows of 16), so
//    that the lanes of a warp reject / accept the same Gaussians and finish together;
//  * Gaussians are staged in batches of 256 through a 2-deep shared-memory ring. Each thread
//    gathers ONE 48-byte packed record {x,y,a,b | c,thr,op,r | g,b,-,-} (written by preprocess)
//    with three 16-byte cp.async (LDGSTS) copies — no register staging, one or two sectors per
//    Gaussian instead of nine scalar loads from four arrays — and the gather of batch k+1 overlaps
//    the blending of batch k;
//  * per evaluation the hot path is 2 broadcast LDS + 9 FP32 ops + 2 compares: `power` is computed
//    with the reference's exact rounding, and compared against a per-Gaussian conservative bound
//    thr = -log(255*op) - 1e-4 so that the accurate expf (11 instructions + MUFU) only runs for
//    evaluations that can contribute. Those take the exact path: same expf, same comparisons, same
//    rounding of T as the reference, so n_contrib and final_T are bit-identical;
//  * early termination: per-warp vote on `done`, per-CTA __syncthreads_and between batches;
//  * backward: the warp walks the range back to front in lock step, sums the nine per-Gaussian
//    gradient terms across its 32 pixels with shuffles, and lane 0 issues two 128-bit vector
//    reductions + one scalar reduction (red.global.add.v4.f32) per (warp, Gaussian) — instead of
//    nine scalar atomics per (pixel, Gaussian).
#include "common.cuh"

namespace cugs {

constexpr int kBlendThreads = 256;
constexpr int kBatch = 256;  // Gaussians per staged batch (one per thread)
constexpr float kAlphaMin = 1.0f / 255.0f;
constexpr float kTMin = 1.0f / 255.0f;  // forward.cuh:25-31 kTransmittanceThreshold

struct __align__(16) StagedGaussian {
    float4 q0;  // x, y, a, b
    float4 q1;  // c, thr, opacity, r
    float4 q2;  // g, b, -, -
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "f"(a), "f"(b), "f"(c),
                 "f"(d)
                 : "memory");
}

// The reference's `power`, rounding for rounding (forward.cu:131-132 and
// backward.cu:132-133 compile to the same sequence).
__device__ __forceinline__ float blend_power(float dx, float dy, float a, float b, float c) {
    const float s1 = fma_rn(dx, a, mul_rn(dy, b));
    const float s2 = fma_rn(dx, b, mul_rn(dy, c));
    return mul_rn(fma_rn(dx, s1, mul_rn(dy, s2)), -0.5f);
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
        dst->q0 = make_float4(m.x, m.y, a, b);
        dst->q1 = make_float4(c, blend_reject_threshold(o), o, rgb[(int64_t)g * 3]);
        dst->q2 = make_float4(rgb[(int64_t)g * 3 + 1], rgb[(int64_t)g * 3 + 2], 0.f, 0.f);
    }
}

__device__ __forceinline__ void pixel_of_thread(int tile_x, int tile_y, int& px, int& py) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    px = tile_x * kTile + (warp & 1) * 8 + (lane & 7);
    py = tile_y * kTile + (warp >> 1) * 4 + (lane >> 3);
}

// ================================================================================================
// forward
// ================================================================================================
template <bool kPacked>
__global__ void __launch_bounds__(kBlendThreads)
k_blend_fwd(int ntx, int width, int height, float bg_r, float bg_g, float bg_b,
            const int* __restrict__ tile_ranges, const int* __restrict__ gaussian_idx,
            const float4* __restrict__ packed, const float* __restrict__ means_2d,
            const float* __restrict__ conic, const float* __restrict__ rgb,
            const float* __restrict__ opa, float* __restrict__ out_color,
            float* __restrict__ out_T, int* __restrict__ out_n) {
    __shared__ StagedGaussian s_g[2][kBatch];

    const int tile = blockIdx.x;
    const int tile_x = tile % ntx, tile_y = tile / ntx;
    int px, py;
    pixel_of_thread(tile_x, tile_y, px, py);
    const bool inside = (px < width) && (py < height);
    const float pxf = (float)px + 0.5f, pyf = (float)py + 0.5f;  // forward.cu:72-73

    const int2 range = reinterpret_cast<const int2*>(tile_ranges)[tile];
    const int count = range.y - range.x;
    const int nb = (count + kBatch - 1) / kBatch;

    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
    int contrib = 0;
    bool done = !inside;

    // prologue: gather batch 0, prefetch the index of batch 1
    int next_idx = -1;
    if (nb > 0) {
        const int li = range.x + threadIdx.x;
        if (li < range.y) stage_gaussian<kPacked>(&s_g[0][threadIdx.x], gaussian_idx[li], packed, means_2d, conic, rgb, opa);
        cp_async_commit();
        const int li1 = li + kBatch;
        if (li1 < range.y) next_idx = gaussian_idx[li1];
    }

    for (int b = 0; b < nb; ++b) {
        // issue the gather of batch b+1 into the other buffer (its previous reader, batch b-1,
        // finished before the __syncthreads_and at the end of the previous iteration)
        if (b + 1 < nb) {
            if (next_idx >= 0) stage_gaussian<kPacked>(&s_g[(b + 1) & 1][threadIdx.x], next_idx, packed, means_2d, conic, rgb, opa);
            cp_async_commit();
            const int li2 = range.x + (b + 2) * kBatch + threadIdx.x;
            next_idx = (li2 < range.y) ? gaussian_idx[li2] : -1;
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();

        const StagedGaussian* sg = s_g[b & 1];
        const int bc = min(kBatch, count - b * kBatch);
        for (int j0 = 0; j0 < bc; j0 += 16) {
            if (__all_sync(kFull, done)) break;
            const int j1 = min(j0 + 16, bc);
            for (int j = j0; j < j1; ++j) {
                const float4 q0 = sg[j].q0;
                const float2 ct = *reinterpret_cast<const float2*>(&sg[j].q1);
                const float dx = pxf - q0.x, dy = pyf - q0.y;
                const float power = blend_power(dx, dy, q0.z, q0.w, ct.x);
                // cheap reject: cannot reach alpha >= 1/255 (NaNs fall through to the exact path)
                if (done || power < ct.y || power > 0.0f) continue;
                const float2 orr = *reinterpret_cast<const float2*>(&sg[j].q1.z);  // opacity, r
                const float alpha = fminf(mul_rn(orr.x, expf(power)), 0.99f);       // forward.cu:137-140
                if (alpha < kAlphaMin) continue;
                const float2 gb = *reinterpret_cast<const float2*>(&sg[j].q2);
                const float w = mul_rn(alpha, T);
                C0 = fma_rn(w, orr.y, C0);
                C1 = fma_rn(w, gb.x, C1);
                C2 = fma_rn(w, gb.y, C2);
                T = mul_rn(T, add_rn(1.0f, -alpha));
                ++contrib;
                if (T < kTMin) done = true;  // the crossing Gaussian is composited (forward.cu:153)
            }
        }
        if (__syncthreads_and(done)) break;
    }
    cp_async_wait<0>();

    if (inside) {
        const int64_t pi = (int64_t)py * width + px;
        out_color[pi * 3 + 0] = fma_rn(T, bg_r, C0);  // forward.cu:166-168
        out_color[pi * 3 + 1] = fma_rn(T, bg_g, C1);
        out_color[pi * 3 + 2] = fma_rn(T, bg_b, C2);
        out_T[pi] = T;
        out_n[pi] = contrib;
    }
}

// ================================================================================================
// backward
// ================================================================================================
__device__ __forceinline__ float warp_sum(float v) {
    v += __shfl_xor_sync(kFull, v, 16);
    v += __shfl_xor_sync(kFull, v, 8);
    v += __shfl_xor_sync(kFull, v, 4);
    v += __shfl_xor_sync(kFull, v, 2);
    v += __shfl_xor_sync(kFull, v, 1);
    return v;
}

template <bool kPacked>
__global__ void __launch_bounds__(kBlendThreads)
k_blend_bwd(int ntx, int width, int height, float bg_r, float bg_g, float bg_b,
            const int* __restrict__ tile_ranges, const int* __restrict__ gaussian_idx,
            const float4* __restrict__ packed, const float* __restrict__ means_2d,
            const float* __restrict__ conic, const float* __restrict__ rgb,
            const float* __restrict__ opa, const float* __restrict__ dL_dcolor,
            const float* __restrict__ final_T, const int* __restrict__ n_contrib,
            float* __restrict__ grad_acc /* [N,12] */) {
    __shared__ StagedGaussian s_g[2][kBatch];
    __shared__ int s_idx[2][kBatch];

    const int tile = blockIdx.x;
    const int tile_x = tile % ntx, tile_y = tile / ntx;
    int px, py;
    pixel_of_thread(tile_x, tile_y, px, py);
    const bool inside = (px < width) && (py < height);
    const float pxf = (float)px + 0.5f, pyf = (float)py + 0.5f;
    const int lane = threadIdx.x & 31;

    const int2 range = reinterpret_cast<const int2*>(tile_ranges)[tile];
    const int count = range.y - range.x;
    const int nb = (count + kBatch - 1) / kBatch;

    // per-pixel forward outputs (backward.cu:65-87)
    float T = 0.0f, g0 = 0.f, g1 = 0.f, g2 = 0.f;
    int maxc = 0;
    if (inside) {
        const int64_t pi = (int64_t)py * width + px;
        T = final_T[pi];
        maxc = n_contrib[pi];
        g0 = dL_dcolor[pi * 3 + 0];
        g1 = dL_dcolor[pi * 3 + 1];
        g2 = dL_dcolor[pi * 3 + 2];
    }
    float S0 = T * bg_r, S1 = T * bg_g, S2 = T * bg_b;
    int found = 0;
    bool done = !inside || maxc <= 0;

    // batches are visited last to first; batch b covers [range.x + b*kBatch, ...)
    int next_idx = -1;
    if (nb > 0) {
        const int li = range.x + (nb - 1) * kBatch + threadIdx.x;
        if (li < range.y) {
            const int g = gaussian_idx[li];
            s_idx[(nb - 1) & 1][threadIdx.x] = g;
            stage_gaussian<kPacked>(&s_g[(nb - 1) & 1][threadIdx.x], g, packed, means_2d, conic, rgb, opa);
        }
        cp_async_commit();
        if (nb > 1) next_idx = gaussian_idx[li - kBatch];  // batch nb-2 is always full
    }

    for (int b = nb - 1; b >= 0; --b) {
        if (b > 0) {
            s_idx[(b - 1) & 1][threadIdx.x] = next_idx;
            stage_gaussian<kPacked>(&s_g[(b - 1) & 1][threadIdx.x], next_idx, packed, means_2d, conic, rgb, opa);
            cp_async_commit();
            if (b > 1) next_idx = gaussian_idx[range.x + (b - 2) * kBatch + threadIdx.x];
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();

        const StagedGaussian* sg = s_g[b & 1];
        const int* sid = s_idx[b & 1];
        const int bc = min(kBatch, count - b * kBatch);
        for (int j = bc - 1; j >= 0; --j) {
            if ((j & 7) == 7 && __all_sync(kFull, done)) break;
            const float4 q0 = sg[j].q0;
            const float4 q1 = sg[j].q1;
            const float dx = pxf - q0.x, dy = pyf - q0.y;
            const float a = q0.z, bq = q0.w, c = q1.x;
            const float power = blend_power(dx, dy, a, bq, c);
            bool hit = false;
            float ex = 0.f, alpha = 0.f;
            if (!(done || power < q1.y || power > 0.0f)) {
                ex = expf(power);
                alpha = fminf(mul_rn(q1.z, ex), 0.99f);
                hit = !(alpha < kAlphaMin);
            }
            if (hit) {
                ++found;
                if (found > maxc) { done = true; hit = false; }  // backward.cu:141-145
            }
            if (!__any_sync(kFull, hit)) continue;

            float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f, v5 = 0.f, v6 = 0.f, v7 = 0.f, v8 = 0.f;
            if (hit) {
                const float2 gb = *reinterpret_cast<const float2*>(&sg[j].q2);
                const float cr = q1.w, cg = gb.x, cb = gb.y;
                const float oma = fmaxf(1.0f - alpha, 1e-5f);  // backward.cu:150-151
                T = T / oma;
                const float w = alpha * T;
                v0 = g0 * w; v1 = g1 * w; v2 = g2 * w;
                float dLa = 0.0f;
                dLa += g0 * (T * cr - S0 / oma);
                dLa += g1 * (T * cg - S1 / oma);
                dLa += g2 * (T * cb - S2 / oma);
                S0 += w * cr; S1 += w * cg; S2 += w * cb;
                const bool clamped = (q1.z * ex >= 0.99f);
                const float dLp = clamped ? 0.0f : dLa * alpha;
                v3 = clamped ? 0.0f : dLa * ex;
                v4 = dLp * (a * dx + bq * dy);
                v5 = dLp * (bq * dx + c * dy);
                v6 = dLp * (-0.5f * dx * dx);
                v7 = dLp * (-dx * dy);
                v8 = dLp * (-0.5f * dy * dy);
                if (found == maxc) done = true;  // nothing left for this pixel
            }
            v0 = warp_sum(v0); v1 = warp_sum(v1); v2 = warp_sum(v2);
            v3 = warp_sum(v3); v4 = warp_sum(v4); v5 = warp_sum(v5);
            v6 = warp_sum(v6); v7 = warp_sum(v7); v8 = warp_sum(v8);
            if (lane == 0) {
                float* acc = grad_acc + (int64_t)sid[j] * 12;
                red_add_v4(acc, v0, v1, v2, v3);
                red_add_v4(acc + 4, v4, v5, v6, v7);
                atomicAdd(acc + 8, v8);
            }
        }
        if (__syncthreads_and(done)) break;
    }
    cp_async_wait<0>();
}

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
