// preprocess.cu — fused per-Gaussian kernels (forward and backward) for sm_100a.
//
// Forward  = k_project_gaussians (reference rasterizer/projection.cu:55-189) + view directions
//            (:273-280) + k_evaluate_sh (core/sh.cu:19-79) + clamp_min(0) (:284), ONE launch.
// Backward = k_project_backward (rasterizer/projection_backward.cu:26-247) + directions (:332-338)
//            + k_evaluate_sh_backward (core/sh_backward.cu:29-112) [+ accumulate_gradients,
//            optimizer/densification.cpp:59-88], ONE launch.
//
// Design (B200): one warp owns 32 consecutive Gaussians.
//   phase A: lane = Gaussian. Projection math with every rounding spelled out (bit-exact radii,
//            tile counts, depth bits and means_2d versus the reference's nvcc-contracted code).
//            The lane also evaluates the SH basis Y[16] for its Gaussian into shared memory.
//   phase B: the warp streams its 32 x 192 B = 6 KB contiguous block of SH coefficients (or SH
//            gradients) as 12 fully coalesced 512-byte float4 rows; float4 q of the block belongs
//            to Gaussian q/12, channel (q%12)/4, coefficients 4*(q%4)..+3, so a 2-step xor-shuffle
//            over groups of 4 lanes finishes one (Gaussian, channel) dot product and the result
//            index is simply q/4 — no strided 192-byte-per-thread access anywhere.
// Summation order of the SH dot product: the generic path (any C) sums k ascending like core/sh.cu:70-76; the
// C = 16 fast path sums four 4-coefficient partials and combines them with two shuffles, so its rgb differs from
// the reference by rounding (<= 2e-6 absolute, tests/test_gpu_parity.py) and the backward's ReLU gate `rgb > 0`
// can flip only for colours inside that noise of the clamp (test_sh_colours_at_the_clamp_boundary_vs_reference).
// HBM traffic is the compulsory 284 B/Gaussian forward (+48 B for the packed blend record) and
// 336 B/Gaussian backward (the ReLU gate comes from the forward rgb, not from re-reading SH).
#include "common.cuh"

namespace cugs {

constexpr int kPreBlock = 256;
constexpr int kPreWarps = kPreBlock / 32;
constexpr int kYStride = 20;  // floats per Gaussian row of the Y table (16 + 4 pad: conflict-free LDS.128)

struct ProjFwd {
    float t[3];
    float xs, ys;
    float op;
    float s[3];
    float R[9];
    float M[9];
    float cov3[6];
    float Tm[6];
    float c2[3];      // Sigma' (a, b, c) incl. low-pass
    float det;
    float inv_norm;
};

// ---- pieces of the forward math; every FMA is explicit (see common.cuh) ------------------------
__device__ __forceinline__ void cam_transform(const ViewParams& vp, float px, float py, float pz,
                                              float t[3]) {
    // projection.cu:97-99  W0*px + W1*py + W2*pz + t  ->  fadd(fma(pz,W2,fma(px,W0,py*W1)), t)
    t[0] = add_rn(fma_rn(pz, vp.W[2], fma_rn(px, vp.W[0], mul_rn(py, vp.W[1]))), vp.t[0]);
    t[1] = add_rn(fma_rn(pz, vp.W[5], fma_rn(px, vp.W[3], mul_rn(py, vp.W[4]))), vp.t[1]);
    t[2] = add_rn(fma_rn(pz, vp.W[8], fma_rn(px, vp.W[6], mul_rn(py, vp.W[7]))), vp.t[2]);
}

__device__ __forceinline__ void quat_rotation(float w, float x, float y, float z, float R[9],
                                              float& inv_norm) {
    // projection.cuh:29-49
    const float n2 = add_rn(fma_rn(z, z, fma_rn(y, y, fma_rn(w, w, mul_rn(x, x)))), 1e-12f);
    inv_norm = rsqrtf(n2);
    w = mul_rn(w, inv_norm); x = mul_rn(x, inv_norm); y = mul_rn(y, inv_norm); z = mul_rn(z, inv_norm);
    // nvcc shares y*y and z*z with R[4] / R[8]: in R[0] BOTH products are rounded (sm_100 SASS of
    // k_project_gaussians: FMUL yy, FMUL zz, FADD), in R[4] / R[8] the x*x product is fused.
    R[0] = fma_rn(-2.0f, add_rn(mul_rn(y, y), mul_rn(z, z)), 1.0f);
    R[1] = mul_rn(2.0f, fma_rn(x, y, -mul_rn(w, z)));
    R[2] = mul_rn(2.0f, dot2c(x, z, w, y));
    R[3] = mul_rn(2.0f, dot2c(x, y, w, z));
    R[4] = fma_rn(-2.0f, dot2c(x, x, z, z), 1.0f);
    R[5] = mul_rn(2.0f, fma_rn(y, z, -mul_rn(w, x)));
    R[6] = mul_rn(2.0f, fma_rn(x, z, -mul_rn(w, y)));
    R[7] = mul_rn(2.0f, dot2c(y, z, w, x));
    R[8] = fma_rn(-2.0f, dot2c(x, x, y, y), 1.0f);
}

__device__ __forceinline__ void covariance_chain(const ViewParams& vp, const float ls[3],
                                                 const float q[4], ProjFwd& f) {
    // projection.cuh:66-90: Sigma = M M^T with M = R diag(exp(log_scale))
    f.s[0] = expf(ls[0]); f.s[1] = expf(ls[1]); f.s[2] = expf(ls[2]);
    quat_rotation(q[0], q[1], q[2], q[3], f.R, f.inv_norm);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) f.M[i * 3 + j] = mul_rn(f.R[i * 3 + j], f.s[j]);
    const float* M = f.M;
    // which product of each 3-term sum is rounded separately is read off the reference's SASS
    f.cov3[0] = dot3c(M[1], M[1], M[0], M[0], M[2], M[2]);
    f.cov3[1] = dot3c(M[0], M[3], M[1], M[4], M[2], M[5]);
    f.cov3[2] = dot3c(M[1], M[7], M[0], M[6], M[2], M[8]);
    f.cov3[3] = dot3c(M[3], M[3], M[4], M[4], M[5], M[5]);
    f.cov3[4] = dot3c(M[3], M[6], M[4], M[7], M[5], M[8]);
    f.cov3[5] = dot3c(M[6], M[6], M[7], M[7], M[8], M[8]);

    // projection.cuh:114-165: Sigma' = (J W) Sigma (J W)^T + 0.3 I. The reference keeps the
    // J[1] = J[3] = 0 products, which only pins which product is rounded separately.
    const float tz_inv = 1.0f / add_rn(f.t[2], 1e-6f);
    const float tz_inv2 = mul_rn(tz_inv, tz_inv);
    const float J0 = mul_rn(vp.fx, tz_inv);
    const float J2 = mul_rn(mul_rn(-vp.fx, f.t[0]), tz_inv2);
    const float J4 = mul_rn(vp.fy, tz_inv);
    const float J5 = mul_rn(mul_rn(-vp.fy, f.t[1]), tz_inv2);
    const float* W = vp.W;
    float* T = f.Tm;
    T[0] = fma_rn(J2, W[6], fma_rn(J0, W[0], mul_rn(0.0f, W[3])));
    T[1] = fma_rn(J2, W[7], fma_rn(J0, W[1], mul_rn(0.0f, W[4])));
    T[2] = fma_rn(J2, W[8], fma_rn(J0, W[2], mul_rn(0.0f, W[5])));
    T[3] = fma_rn(J5, W[6], fma_rn(0.0f, W[0], mul_rn(J4, W[3])));
    T[4] = fma_rn(J5, W[7], fma_rn(0.0f, W[1], mul_rn(J4, W[4])));
    T[5] = fma_rn(J5, W[8], fma_rn(0.0f, W[2], mul_rn(J4, W[5])));
    const float* S = f.cov3;
    float TS[6];
    TS[0] = dot3c(T[0], S[0], T[1], S[1], T[2], S[2]);
    TS[1] = dot3c(T[0], S[1], T[1], S[3], T[2], S[4]);
    TS[2] = dot3c(T[0], S[2], T[1], S[4], T[2], S[5]);
    TS[3] = dot3c(T[3], S[0], T[4], S[1], T[5], S[2]);
    TS[4] = dot3c(T[3], S[1], T[4], S[3], T[5], S[4]);
    TS[5] = dot3c(T[3], S[2], T[4], S[4], T[5], S[5]);
    f.c2[0] = add_rn(dot3c(TS[0], T[0], TS[1], T[1], TS[2], T[2]), 0.3f);
    f.c2[1] = dot3c(TS[0], T[3], TS[1], T[4], TS[2], T[5]);
    f.c2[2] = add_rn(dot3c(TS[3], T[3], TS[4], T[4], TS[5], T[5]), 0.3f);
    // projection.cuh:209-211 det = a*c - b*b
    f.det = fma_rn(f.c2[0], f.c2[2], -mul_rn(f.c2[1], f.c2[1]));
}

__device__ __forceinline__ int radius_from_cov(const float c2[3], float det) {
    // projection.cuh:179-195
    const float trace = add_rn(c2[0], c2[2]);
    const float disc = fmaxf(fma_rn(trace, trace, -mul_rn(4.0f, det)), 0.0f);
    const float lambda_max = mul_rn(0.5f, add_rn(trace, sqrtf(disc)));
    if (lambda_max <= 0.0f) return 0;
    return (int)ceilf(mul_rn(3.0f, sqrtf(lambda_max)));
}

// ================================================================================================
// Forward
// ================================================================================================
// kFull = false: render-only frame, the four backward-only arrays (depths, cov_2d_inv, rgb, opa_act) are
// not written (a compile-time switch: a run-time null test costs the training path 11 registers).
#ifndef CUGS_PRE_MINBLOCKS
#define CUGS_PRE_MINBLOCKS 4  // 64 registers (52 B of spills), 4 x 256 threads per SM: measured 0.196 ms vs 0.214 ms at 3 blocks
#endif
// kDeg3: the active SH degree is 3 (the steady state of training): basis and dot products without the
// run-time degree tests.
template <bool kVecSH, bool kFull, bool kDeg3>
__global__ void __launch_bounds__(kPreBlock, CUGS_PRE_MINBLOCKS)
k_preprocess_fwd(int64_t n, ViewParams vp, const float* __restrict__ positions,
                 const float* __restrict__ rotations, const float* __restrict__ scales,
                 const float* __restrict__ opacities, const float* __restrict__ sh,
                 float* __restrict__ means_2d, float* __restrict__ depths,
                 float* __restrict__ cov_2d_inv, int* __restrict__ radii,
                 int* __restrict__ tiles_touched, float* __restrict__ rgb,
                 float* __restrict__ opa_act, float4* __restrict__ packed,
                 unsigned* __restrict__ depth_minmax, uint64_t* __restrict__ gsort) {
    __shared__ __align__(16) float sY[kPreWarps][32 * kYStride];
    __shared__ float sRGB[kPreWarps][96];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t g0 = ((int64_t)blockIdx.x * kPreWarps + warp) * 32;
    if (g0 >= n) return;  // warp-uniform
    const int64_t i = g0 + lane;
    const bool valid = i < n;

    float px = 0.f, py = 0.f, pz = 0.f;
    if (valid) {
        px = positions[i * 3 + 0];
        py = positions[i * 3 + 1];
        pz = positions[i * 3 + 2];
    }

    // ---- SH basis for this lane's Gaussian (all N, culled or not: core/sh.cu:27-28) ----
    float Y[16];
    {
        float dx, dy, dz;
        view_dir(px, py, pz, vp.cam, dx, dy, dz);
        sh_basis(kDeg3 ? 3 : vp.deg, dx, dy, dz, Y);
    }
    const int na = kDeg3 ? 16 : (vp.deg + 1) * (vp.deg + 1);

    if (kVecSH) {
        float4* row = reinterpret_cast<float4*>(&sY[warp][lane * kYStride]);
        row[0] = make_float4(Y[0], Y[1], Y[2], Y[3]);
        row[1] = make_float4(Y[4], Y[5], Y[6], Y[7]);
        row[2] = make_float4(Y[8], Y[9], Y[10], Y[11]);
        row[3] = make_float4(Y[12], Y[13], Y[14], Y[15]);
    }

    // ---- phase A: projection ----
    float o_x = 0.f, o_y = 0.f, o_depth = 0.f, o_op = 0.f, o_a = 0.f, o_b = 0.f, o_c = 0.f;
    float o_sxx = 0.f, o_syy = 0.f;  // Sigma'_xx, Sigma'_yy (footprint extents for the blend kernels)
    int o_radius = 0, o_tiles = 0;
    bool quirk = false;
    if (valid) {
        ProjFwd f;
        cam_transform(vp, px, py, pz, f.t);
        if (!(f.t[2] <= 0.2f)) {  // projection.cu:104
            o_x = add_rn(mul_rn(vp.fx, f.t[0]) / f.t[2], vp.cx);  // :109-110
            o_y = add_rn(mul_rn(vp.fy, f.t[1]) / f.t[2], vp.cy);
            o_depth = f.t[2];
            o_op = 1.0f / add_rn(1.0f, expf(-opacities[i]));  // :119-121

            const float lsm = logf(add_rn(vp.scale_mod, 1e-8f));  // :126-130
            const float ls[3] = {add_rn(scales[i * 3 + 0], lsm), add_rn(scales[i * 3 + 1], lsm),
                                 add_rn(scales[i * 3 + 2], lsm)};
            const float4 q4 = reinterpret_cast<const float4*>(rotations)[i];
            const float q[4] = {q4.x, q4.y, q4.z, q4.w};
            covariance_chain(vp, ls, q, f);
            if (!(f.det <= 0.0f)) {  // projection.cu:151-152
                const float inv_det = 1.0f / f.det;  // projection.cuh:220-223
                o_a = mul_rn(f.c2[2], inv_det);
                o_b = mul_rn(-f.c2[1], inv_det);
                o_c = mul_rn(f.c2[0], inv_det);
                o_sxx = f.c2[0];
                o_syy = f.c2[2];
                int radius = radius_from_cov(f.c2, f.det);
                if (radius > 0) {
                    radius = min(radius, max(vp.width, vp.height));  // :165-166
                    o_radius = radius;
                    const TileRect r = tile_rect(o_x, o_y, radius, vp.width, vp.height, vp.ntx, vp.nty);
                    // A.1-11: the product of two negative extents is kept (reference quirk)
                    o_tiles = max((r.tx1 - r.tx0) * (r.ty1 - r.ty0), 0);
                    quirk = (o_tiles > 0) && (r.tx1 - r.tx0 < 0);  // slots will hold key 0 (A.2)
                }
            }
        }
        reinterpret_cast<float2*>(means_2d)[i] = make_float2(o_x, o_y);
        if (kFull) {
            depths[i] = o_depth;
            cov_2d_inv[i * 3 + 0] = o_a;
            cov_2d_inv[i * 3 + 1] = o_b;
            cov_2d_inv[i * 3 + 2] = o_c;
            opa_act[i] = o_op;
        }
        radii[i] = o_radius;
        tiles_touched[i] = o_tiles;
        // element of the depth sort (tile_binning.cu): depth bits << 32 | index; key 0 for the
        // filler entries of quirk A.2, 0xffffffff (sorts last) when the Gaussian emits no pair
        if (gsort != nullptr) {
            const unsigned key = (o_tiles > 0) ? (quirk ? 0u : __float_as_uint(o_depth)) : 0xffffffffu;
            gsort[i] = ((uint64_t)key << 32) | (uint64_t)(unsigned)i;
        }
    }

    if (depth_minmax != nullptr) {
        unsigned dmin = (valid && o_tiles > 0) ? (quirk ? 0u : __float_as_uint(o_depth)) : 0xffffffffu;
        unsigned dmax = (valid && o_tiles > 0) ? __float_as_uint(o_depth) : 0u;
        dmin = __reduce_min_sync(kFull, dmin);
        dmax = __reduce_max_sync(kFull, dmax);
        if (lane == 0 && dmin <= dmax) {
            atomicMin(&depth_minmax[0], dmin);
            atomicMax(&depth_minmax[1], dmax);
        }
    }

    // ---- phase B: SH colour ----
    float c_r, c_g, c_b;
    if (kVecSH) {
        __syncwarp();
        const float4* sh4 = reinterpret_cast<const float4*>(sh) + g0 * 12;
        const int64_t lim = (n - g0) * 12;  // float4s available from this warp's base
        float4 cv[12];
#pragma unroll
        for (int it = 0; it < 12; ++it) {
            const int q = lane + 32 * it;
            const int k0 = (q & 3) * 4;
            cv[it] = (q < lim && k0 < na) ? __ldcs(sh4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int it = 0; it < 12; ++it) {
            const int q = lane + 32 * it;
            const int g = q / 12;
            const int k0 = (q & 3) * 4;
            const float4 y4 = *reinterpret_cast<const float4*>(&sY[warp][g * kYStride + k0]);
            float v = 0.0f;
            if (k0 + 0 < na) v = cv[it].x * y4.x;
            if (k0 + 1 < na) v = fmaf(cv[it].y, y4.y, v);
            if (k0 + 2 < na) v = fmaf(cv[it].z, y4.z, v);
            if (k0 + 3 < na) v = fmaf(cv[it].w, y4.w, v);
            v += __shfl_xor_sync(kFull, v, 1);
            v += __shfl_xor_sync(kFull, v, 2);
            if ((lane & 3) == 0) {
                const float col = fmaxf(v + 0.5f, 0.0f);  // +0.5 (sh.cu:77), clamp_min(0) (projection.cu:284)
                const int o = q >> 2;                     // = g*3 + channel
                sRGB[warp][o] = col;
                if (kFull && q < lim) rgb[g0 * 3 + o] = col;
            }
        }
        __syncwarp();
        c_r = sRGB[warp][lane * 3 + 0];
        c_g = sRGB[warp][lane * 3 + 1];
        c_b = sRGB[warp][lane * 3 + 2];
    } else {
        // generic coefficient count (C = 1, 4, 9, ...): per-lane, reference summation order
        float col[3] = {0.f, 0.f, 0.f};
        if (valid) {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const float* c = sh + (i * 3 + ch) * vp.C;
                float acc = 0.0f;
                for (int k = 0; k < na; ++k) acc += c[k] * Y[k];
                col[ch] = fmaxf(acc + 0.5f, 0.0f);
                if (kFull) rgb[i * 3 + ch] = col[ch];
            }
        }
        c_r = col[0]; c_g = col[1]; c_b = col[2];
    }

    // ---- packed blend record {x,y,a,b | c,thr,op,r | g,b,hx,hy} ----
    if (packed != nullptr && valid) {
        float4* rec = packed + i * 3;
        const float thr = blend_reject_threshold(o_op);
        float hx, hy;
        blend_extents(thr, o_sxx, o_syy, hx, hy);
        rec[0] = make_float4(o_x, o_y, o_a, o_b);
        rec[1] = make_float4(o_c, thr, o_op, c_r);
        rec[2] = make_float4(c_g, c_b, hx, hy);
    }
}

// ================================================================================================
// Backward
// ================================================================================================
#ifndef CUGS_PREBWD_MINBLOCKS
#define CUGS_PREBWD_MINBLOCKS 4
#endif
template <bool kVecSH, bool kDeg3>
__global__ void __launch_bounds__(kPreBlock, CUGS_PREBWD_MINBLOCKS)
k_preprocess_bwd(int64_t n, ViewParams vp, const float* __restrict__ positions,
                 const float* __restrict__ rotations, const float* __restrict__ scales,
                 const float* __restrict__ opacities, const float* __restrict__ sh,
                 const int* __restrict__ radii, const float* __restrict__ rgb,
                 const float* __restrict__ dL_dmeans_2d, const float* __restrict__ dL_dconic,
                 const float* __restrict__ dL_drgb, const float* __restrict__ dL_dopa_act,
                 float* __restrict__ dL_dpos, float* __restrict__ dL_drot,
                 float* __restrict__ dL_dscl, float* __restrict__ dL_dopa,
                 float* __restrict__ dL_dsh, float* __restrict__ grad_accum,
                 float* __restrict__ grad_count, float* __restrict__ max_radii,
                 const float4* __restrict__ gacc /* [N,3] packed blend gradients or null */,
                 float* __restrict__ dL_dmeans_2d_out /* written when gacc != null */,
                 bool accumulate /* add to the five parameter-gradient outputs instead of overwriting */,
                 int* __restrict__ touch_mask /* optional: 1 where the incoming 2-D gradient is non-zero */,
                 bool sparse_rows /* with touch_mask: only move gradient rows that can be non-zero */,
                 const int* __restrict__ list /* optional: compacted indices of the Gaussians to process */,
                 const int* __restrict__ list_count /* device count of `list` */) {
    __shared__ __align__(16) float sY[kPreWarps][32 * kYStride];
    __shared__ float sG[kPreWarps][96];  // gated dL/drgb per (Gaussian, channel)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t g0 = ((int64_t)blockIdx.x * kPreWarps + warp) * 32;
    // list mode (second phase of the sparse backward, see k_bwd_classify): the warp owns 32 consecutive
    // entries of the list of touched Gaussians instead of 32 consecutive Gaussians
    const int64_t limit = list ? (int64_t)*list_count : n;
    if (g0 >= limit) return;
    const bool valid = g0 + lane < limit;
    const int64_t i = list ? (valid ? (int64_t)list[g0 + lane] : 0) : g0 + lane;

    float px = 0.f, py = 0.f, pz = 0.f;
    int radius = 0;
    float gate_g[3] = {0.f, 0.f, 0.f};
    float in_m0 = 0.f, in_m1 = 0.f, in_da = 0.f, in_db = 0.f, in_dc = 0.f, in_dop = 0.f;
    // sparse_rows: the caller guarantees that every gradient row whose mask entry is 0 on entry is all
    // zero (true after allocation with zeros, after this kernel, and after the sparse exchange). A
    // row is then written only if it is touched now or may hold an old value (overwrite mode), and
    // read-modified-written only if it is touched now (accumulate mode): ~80 % of the rows of a
    // view are skipped entirely.
    bool write_row = valid;
    // sparse_rows: a Gaussian whose nine incoming 2-D gradients are all exactly zero has all-zero parameter
    // gradients; the whole chain rule (a full re-projection) is skipped for it (~80 % of a view's Gaussians)
    bool skip_chain = false;
    if (valid) {
        px = positions[i * 3 + 0]; py = positions[i * 3 + 1]; pz = positions[i * 3 + 2];
        radius = radii[i];
        float dr[3];
        if (gacc != nullptr) {  // {drgb.xyz, dop | dmean.xy, da, db | dc, -, -, -}
            const float4 a = gacc[i * 3], b = gacc[i * 3 + 1], c = gacc[i * 3 + 2];
            dr[0] = a.x; dr[1] = a.y; dr[2] = a.z; in_dop = a.w;
            in_m0 = b.x; in_m1 = b.y; in_da = b.z; in_db = b.w; in_dc = c.x;
            if (dL_dmeans_2d_out != nullptr) reinterpret_cast<float2*>(dL_dmeans_2d_out)[i] = make_float2(in_m0, in_m1);
            if (touch_mask != nullptr) {
                const bool t = (a.x != 0.f) | (a.y != 0.f) | (a.z != 0.f) | (a.w != 0.f) | (b.x != 0.f) | (b.y != 0.f) |
                               (b.z != 0.f) | (b.w != 0.f) | (c.x != 0.f);
                const int old = (accumulate || sparse_rows) ? touch_mask[i] : 0;
                touch_mask[i] = accumulate ? (old | (int)t) : (int)t;
                if (sparse_rows) {
                    write_row = accumulate ? t : (t || old != 0);
                    skip_chain = !t;
                }
            }
        } else {
            dr[0] = dL_drgb[i * 3]; dr[1] = dL_drgb[i * 3 + 1]; dr[2] = dL_drgb[i * 3 + 2];
            in_dop = dL_dopa_act[i];
            const float2 dm = reinterpret_cast<const float2*>(dL_dmeans_2d)[i];
            in_m0 = dm.x; in_m1 = dm.y;
            in_da = dL_dconic[i * 3 + 0]; in_db = dL_dconic[i * 3 + 1]; in_dc = dL_dconic[i * 3 + 2];
        }
        // ReLU gate of projection.cu:284: forward rgb > 0  <=>  raw SH colour + 0.5 > 0
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) gate_g[ch] = (rgb[i * 3 + ch] > 0.0f) ? dr[ch] : 0.0f;
    }
    float Y[16];
    {
        float dx, dy, dz;
        view_dir(px, py, pz, vp.cam, dx, dy, dz);
        sh_basis(kDeg3 ? 3 : vp.deg, dx, dy, dz, Y);
    }
    const int na = kDeg3 ? 16 : (vp.deg + 1) * (vp.deg + 1);
    if (kVecSH) {
        float4* row = reinterpret_cast<float4*>(&sY[warp][lane * kYStride]);
        row[0] = make_float4(Y[0], Y[1], Y[2], Y[3]);
        row[1] = make_float4(Y[4], Y[5], Y[6], Y[7]);
        row[2] = make_float4(Y[8], Y[9], Y[10], Y[11]);
        row[3] = make_float4(Y[12], Y[13], Y[14], Y[15]);
        sG[warp][lane * 3 + 0] = gate_g[0];
        sG[warp][lane * 3 + 1] = gate_g[1];
        sG[warp][lane * 3 + 2] = gate_g[2];
    }

    // ---- phase A: chain rule to position / rotation / log-scale / logit-opacity ----
    float g_pos[3] = {0.f, 0.f, 0.f}, g_rot[4] = {0.f, 0.f, 0.f, 0.f}, g_scl[3] = {0.f, 0.f, 0.f};
    float g_opa = 0.f;
    const float m0 = in_m0, m1 = in_m1;
    if (valid && radius > 0 && !skip_chain) {  // projection_backward.cu:48
        ProjFwd f;
        cam_transform(vp, px, py, pz, f.t);
        const float lsm = logf(add_rn(vp.scale_mod, 1e-8f));
        const float ls[3] = {add_rn(scales[i * 3 + 0], lsm), add_rn(scales[i * 3 + 1], lsm),
                             add_rn(scales[i * 3 + 2], lsm)};
        const float4 q4 = reinterpret_cast<const float4*>(rotations)[i];
        const float q[4] = {q4.x, q4.y, q4.z, q4.w};
        covariance_chain(vp, ls, q, f);
        if (!(f.det <= 0.0f)) {  // :95
            const float inv_det = 1.0f / f.det;
            const float ia = f.c2[2] * inv_det, ib = -f.c2[1] * inv_det, ic = f.c2[0] * inv_det;
            const float* W = vp.W;
            const float tx = f.t[0], ty = f.t[1], tz = f.t[2];
            const float tz_inv = 1.0f / (tz + 1e-6f), tz_inv2 = tz_inv * tz_inv;
            const float J0 = vp.fx * tz_inv, J2 = -vp.fx * tx * tz_inv2;
            const float J4 = vp.fy * tz_inv, J5 = -vp.fy * ty * tz_inv2;
            // projection_backward.cu:109-115 (T without the zero products)
            const float T[6] = {J0 * W[0] + J2 * W[6], J0 * W[1] + J2 * W[7], J0 * W[2] + J2 * W[8],
                                J4 * W[3] + J5 * W[6], J4 * W[4] + J5 * W[7], J4 * W[5] + J5 * W[8]};

            // backward.cuh:37-64: dSigma' = -Sinv dSinv Sinv, off-diagonal halved first
            const float da = in_da, db = in_db * 0.5f, dc = in_dc;
            const float t00 = ia * da + ib * db, t01 = ia * db + ib * dc;
            const float t10 = ib * da + ic * db, t11 = ib * db + ic * dc;
            const float d2a = -(t00 * ia + t01 * ib);
            const float d2b = -(t00 * ib + t01 * ic);
            const float d2c = -(t10 * ib + t11 * ic);

            // backward.cuh:82-107: dSigma = T^T dSigma' T
            const float u0 = T[0] * d2a + T[3] * d2b, u1 = T[0] * d2b + T[3] * d2c;
            const float u2 = T[1] * d2a + T[4] * d2b, u3 = T[1] * d2b + T[4] * d2c;
            const float u4 = T[2] * d2a + T[5] * d2b, u5 = T[2] * d2b + T[5] * d2c;
            const float d00 = u0 * T[0] + u1 * T[3], d01 = u0 * T[1] + u1 * T[4];
            const float d02 = u0 * T[2] + u1 * T[5], d11 = u2 * T[1] + u3 * T[4];
            const float d12 = u2 * T[2] + u3 * T[5], d22 = u4 * T[2] + u5 * T[5];

            // backward.cuh:123-153: dM = 2 dSigma_full M
            const float* M = f.M;
            float dM[9];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                dM[0 + j] = 2.0f * (d00 * M[j] + d01 * M[3 + j] + d02 * M[6 + j]);
                dM[3 + j] = 2.0f * (d01 * M[j] + d11 * M[3 + j] + d12 * M[6 + j]);
                dM[6 + j] = 2.0f * (d02 * M[j] + d12 * M[3 + j] + d22 * M[6 + j]);
            }
            // projection_backward.cu:170-184
            float dR[9];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int j = 0; j < 3; ++j) dR[r * 3 + j] = dM[r * 3 + j] * f.s[j];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const float ds = dM[j] * f.R[j] + dM[3 + j] * f.R[3 + j] + dM[6 + j] * f.R[6 + j];
                g_scl[j] = ds * f.s[j];
            }
            // backward.cuh:168-227
            {
                const float w = q[0] * f.inv_norm, x = q[1] * f.inv_norm, y = q[2] * f.inv_norm,
                            z = q[3] * f.inv_norm;
                const float dw = 2.0f * (-z * dR[1] + y * dR[2] + z * dR[3] - x * dR[5] + -y * dR[6] + x * dR[7]);
                const float dx = 2.0f * (y * dR[1] + z * dR[2] + y * dR[3] - 2.0f * x * dR[4] - w * dR[5] +
                                         z * dR[6] + w * dR[7] - 2.0f * x * dR[8]);
                const float dy = 2.0f * (-2.0f * y * dR[0] + x * dR[1] + w * dR[2] + x * dR[3] + z * dR[5] +
                                         -w * dR[6] + z * dR[7] - 2.0f * y * dR[8]);
                const float dz = 2.0f * (-2.0f * z * dR[0] - w * dR[1] + x * dR[2] + w * dR[3] -
                                         2.0f * z * dR[4] + y * dR[5] + x * dR[6] + y * dR[7]);
                const float dot = dw * w + dx * x + dy * y + dz * z;
                g_rot[0] = f.inv_norm * (dw - w * dot);
                g_rot[1] = f.inv_norm * (dx - x * dot);
                g_rot[2] = f.inv_norm * (dy - y * dot);
                g_rot[3] = f.inv_norm * (dz - z * dot);
            }
            // projection_backward.cu:196-205: through means_2d
            float dt0 = m0 * vp.fx * tz_inv;
            float dt1 = m1 * vp.fy * tz_inv;
            float dt2 = m0 * (-vp.fx * tx * tz_inv2) + m1 * (-vp.fy * ty * tz_inv2);
            // backward.cuh:248-346: through J's dependence on t_cam
            {
                const float* S = f.cov3;
                const float TS0 = T[0] * S[0] + T[1] * S[1] + T[2] * S[2];
                const float TS1 = T[0] * S[1] + T[1] * S[3] + T[2] * S[4];
                const float TS2 = T[0] * S[2] + T[1] * S[4] + T[2] * S[5];
                const float TS3 = T[3] * S[0] + T[4] * S[1] + T[5] * S[2];
                const float TS4 = T[3] * S[1] + T[4] * S[3] + T[5] * S[4];
                const float TS5 = T[3] * S[2] + T[4] * S[4] + T[5] * S[5];
                const float dT0 = 2.0f * (d2a * TS0 + d2b * TS3), dT1 = 2.0f * (d2a * TS1 + d2b * TS4);
                const float dT2 = 2.0f * (d2a * TS2 + d2b * TS5), dT3 = 2.0f * (d2b * TS0 + d2c * TS3);
                const float dT4 = 2.0f * (d2b * TS1 + d2c * TS4), dT5 = 2.0f * (d2b * TS2 + d2c * TS5);
                const float dJ0 = dT0 * W[0] + dT1 * W[1] + dT2 * W[2];
                const float dJ2 = dT0 * W[6] + dT1 * W[7] + dT2 * W[8];
                const float dJ4 = dT3 * W[3] + dT4 * W[4] + dT5 * W[5];
                const float dJ5 = dT3 * W[6] + dT4 * W[7] + dT5 * W[8];
                const float tz_inv3 = tz_inv2 * tz_inv;
                dt0 += dJ2 * (-vp.fx * tz_inv2);
                dt1 += dJ5 * (-vp.fy * tz_inv2);
                dt2 += dJ0 * (-vp.fx * tz_inv2) + dJ2 * (2.0f * vp.fx * tx * tz_inv3) +
                       dJ4 * (-vp.fy * tz_inv2) + dJ5 * (2.0f * vp.fy * ty * tz_inv3);
            }
            // :217-219
            g_pos[0] = W[0] * dt0 + W[3] * dt1 + W[6] * dt2;
            g_pos[1] = W[1] * dt0 + W[4] * dt1 + W[7] * dt2;
            g_pos[2] = W[2] * dt0 + W[5] * dt1 + W[8] * dt2;
            // :226-228
            const float sg = 1.0f / (1.0f + expf(-opacities[i]));
            g_opa = in_dop * sg * (1.0f - sg);
        }
    }
    if (write_row) {
        if (accumulate) {
            g_pos[0] += dL_dpos[i * 3 + 0]; g_pos[1] += dL_dpos[i * 3 + 1]; g_pos[2] += dL_dpos[i * 3 + 2];
            const float4 r4 = reinterpret_cast<const float4*>(dL_drot)[i];
            g_rot[0] += r4.x; g_rot[1] += r4.y; g_rot[2] += r4.z; g_rot[3] += r4.w;
            g_scl[0] += dL_dscl[i * 3 + 0]; g_scl[1] += dL_dscl[i * 3 + 1]; g_scl[2] += dL_dscl[i * 3 + 2];
            g_opa += dL_dopa[i];
        }
        dL_dpos[i * 3 + 0] = g_pos[0]; dL_dpos[i * 3 + 1] = g_pos[1]; dL_dpos[i * 3 + 2] = g_pos[2];
        reinterpret_cast<float4*>(dL_drot)[i] = make_float4(g_rot[0], g_rot[1], g_rot[2], g_rot[3]);
        dL_dscl[i * 3 + 0] = g_scl[0]; dL_dscl[i * 3 + 1] = g_scl[1]; dL_dscl[i * 3 + 2] = g_scl[2];
        dL_dopa[i] = g_opa;
    }
    if (valid) {
        if (grad_accum != nullptr) {  // optimizer/densification.cpp:59-88
            if (radius > 0) {
                grad_accum[i] += sqrtf(m0 * m0 + m1 * m1);
                grad_count[i] += 1.0f;
            }
            max_radii[i] = fmaxf(max_radii[i], (float)radius);
        }
    }

    // ---- phase B: dL/dSH = gate * dL/drgb * Y_k, explicit zeros for inactive coefficients ----
    const unsigned row_mask = __ballot_sync(kFull, write_row);  // bit g: the warp's Gaussian g moves its rows
    if (kVecSH) {
        __syncwarp();
        float4* const sh_out4 = reinterpret_cast<float4*>(dL_dsh);
#pragma unroll
        for (int it = 0; it < 12; ++it) {
            const int q = lane + 32 * it;
            const int g = q / 12;
            // row of Gaussian i_g = 12 float4; q - 12 g = float4 inside the row (contiguous mode: i_g = g0 + g,
            // so the warp still writes its 6 KB block as 12 fully coalesced 512-byte rows)
            const int64_t ig = __shfl_sync(kFull, i, g);
            if (!((row_mask >> g) & 1u)) continue;  // also drops rows beyond the end (write_row = false there)
            float4* out4 = sh_out4 + ig * 12 - (int64_t)g * 12;
            const int k0 = (q & 3) * 4;
            const float gd = sG[warp][q >> 2];
            const float4 y4 = *reinterpret_cast<const float4*>(&sY[warp][g * kYStride + k0]);
            float4 o;
            o.x = (k0 + 0 < na) ? gd * y4.x : 0.0f;
            o.y = (k0 + 1 < na) ? gd * y4.y : 0.0f;
            o.z = (k0 + 2 < na) ? gd * y4.z : 0.0f;
            o.w = (k0 + 3 < na) ? gd * y4.w : 0.0f;
            if (accumulate) {
                const float4 old = __ldcs(out4 + q);
                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
            }
            __stcs(out4 + q, o);
        }
    } else if (write_row) {
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            float* o = dL_dsh + (i * 3 + ch) * vp.C;
            if (accumulate) {
                for (int k = 0; k < na; ++k) o[k] += gate_g[ch] * Y[k];
            } else {
                for (int k = 0; k < na; ++k) o[k] = gate_g[ch] * Y[k];
                for (int k = na; k < vp.C; ++k) o[k] = 0.0f;
            }
        }
    }
}

// ================================================================================================
// Sparse backward, phase 1 (cugs_b200_render_backward with CUGS_BWD_SPARSE_ROWS): one pass over the N
// packed 2-D gradient records that does everything that is per-Gaussian bookkeeping -- the touch mask,
// dL/dmeans_2d, the fused densification statistics (optimizer/densification.cpp:59-88), zeroing of rows
// that held a value from an earlier step but are not touched now -- and appends the touched Gaussians
// (some gradient non-zero: ~20 % of a view) to a compact list. Phase 2 is k_preprocess_bwd in list mode:
// the chain rule (a full re-projection, projection_backward.cu:26-247) and the 236-byte gradient row only
// for the listed Gaussians, with full warps, instead of 32-lane warps in which ~6 lanes have work.
// ================================================================================================
constexpr int kClassifyThreads = 512;
constexpr int kClassifyChunk = 6144;  // Gaussians per block iteration = capacity of the block's shared index list

// A persistent grid (a few blocks per SM). Every block walks chunks of kClassifyChunk consecutive Gaussians,
// collects the touched ones in a SHARED list (warp-aggregated shared atomics) and reserves their place in the
// global list with ONE global atomic per chunk -- one atomic per warp on a single address (94 k of them at 3 M
// Gaussians) serialised in L2 and made the first version of this pass take 121 us instead of ~40.
__global__ void __launch_bounds__(kClassifyThreads)
k_bwd_classify(int64_t n, int num_coeffs, const float4* __restrict__ gacc, const int* __restrict__ radii,
               int* __restrict__ touch_mask, bool accumulate, float* __restrict__ dL_dmeans_2d_out,
               float* __restrict__ grad_accum, float* __restrict__ grad_count, float* __restrict__ max_radii,
               float* __restrict__ dL_dpos, float* __restrict__ dL_drot, float* __restrict__ dL_dscl,
               float* __restrict__ dL_dopa, float* __restrict__ dL_dsh, int* __restrict__ list,
               int* __restrict__ list_count) {
    __shared__ int s_list[kClassifyChunk];
    __shared__ int s_count, s_base;
    const int lane = threadIdx.x & 31;
    for (int64_t c0 = (int64_t)blockIdx.x * kClassifyChunk; c0 < n; c0 += (int64_t)gridDim.x * kClassifyChunk) {
        if (threadIdx.x == 0) s_count = 0;
        __syncthreads();
        const int64_t c1 = min(n, c0 + kClassifyChunk);
        for (int64_t i0 = c0; i0 < c1; i0 += kClassifyThreads) {  // block-uniform trip count
            const int64_t i = i0 + threadIdx.x;
            bool t = false;
            if (i < c1) {
                const float4 a = __ldcs(gacc + i * 3), b = __ldcs(gacc + i * 3 + 1), c = __ldcs(gacc + i * 3 + 2);
                t = (a.x != 0.f) | (a.y != 0.f) | (a.z != 0.f) | (a.w != 0.f) | (b.x != 0.f) | (b.y != 0.f) |
                    (b.z != 0.f) | (b.w != 0.f) | (c.x != 0.f);
                reinterpret_cast<float2*>(dL_dmeans_2d_out)[i] = make_float2(b.x, b.y);
                const int old = touch_mask[i];
                touch_mask[i] = accumulate ? (old | (int)t) : (int)t;
                if (grad_accum != nullptr) {
                    const int radius = radii[i];
                    if (radius > 0) {
                        grad_accum[i] += sqrtf(b.x * b.x + b.y * b.y);
                        grad_count[i] += 1.0f;
                    }
                    max_radii[i] = fmaxf(max_radii[i], (float)radius);
                }
                if (!accumulate && !t && old != 0) {  // stale row of an earlier step: back to zero
                    dL_dpos[i * 3 + 0] = 0.f; dL_dpos[i * 3 + 1] = 0.f; dL_dpos[i * 3 + 2] = 0.f;
                    reinterpret_cast<float4*>(dL_drot)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    dL_dscl[i * 3 + 0] = 0.f; dL_dscl[i * 3 + 1] = 0.f; dL_dscl[i * 3 + 2] = 0.f;
                    dL_dopa[i] = 0.f;
                    float* o = dL_dsh + i * 3 * num_coeffs;
                    for (int k = 0; k < 3 * num_coeffs; ++k) o[k] = 0.f;
                }
            }
            const unsigned m = __ballot_sync(kFull, t);
            if (m != 0u) {
                int base = 0;
                if (lane == __ffs(m) - 1) base = atomicAdd(&s_count, __popc(m));
                base = __shfl_sync(kFull, base, __ffs(m) - 1);
                if (t) s_list[base + __popc(m & ((1u << lane) - 1))] = (int)i;
            }
        }
        __syncthreads();
        const int cnt = s_count;
        if (threadIdx.x == 0) s_base = cnt ? atomicAdd(list_count, cnt) : 0;
        __syncthreads();
        const int base = s_base;
        for (int k = threadIdx.x; k < cnt; k += kClassifyThreads) list[base + k] = s_list[k];
        __syncthreads();  // s_list / s_count are reused by the next chunk
    }
}

// ================================================================================================
// Stage functions: evaluate_sh_cuda / evaluate_sh_backward_cuda with explicit directions
// (core/sh.cu:81-123, core/sh_backward.cu:114-156). Reference summation order, no clamp.
// ================================================================================================
__global__ void __launch_bounds__(256)
k_sh_forward(int64_t n, int deg, int C, const float* __restrict__ sh, const float* __restrict__ dirs,
             float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float Y[16];
    sh_basis(deg, dirs[i * 3], dirs[i * 3 + 1], dirs[i * 3 + 2], Y);
    const int na = (deg + 1) * (deg + 1);
    for (int ch = 0; ch < 3; ++ch) {
        const float* c = sh + (i * 3 + ch) * C;
        float acc = 0.0f;
        for (int k = 0; k < na; ++k) acc += c[k] * Y[k];
        out[i * 3 + ch] = acc + 0.5f;
    }
}

__global__ void __launch_bounds__(256)
k_sh_backward(int64_t n, int deg, int C, const float* __restrict__ sh, const float* __restrict__ dirs,
              const float* __restrict__ dL_drgb, float* __restrict__ dL_dsh) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float Y[16];
    sh_basis(deg, dirs[i * 3], dirs[i * 3 + 1], dirs[i * 3 + 2], Y);
    const int na = (deg + 1) * (deg + 1);
    for (int ch = 0; ch < 3; ++ch) {
        const float* c = sh + (i * 3 + ch) * C;
        float* o = dL_dsh + (i * 3 + ch) * C;
        float raw = 0.0f;
        for (int k = 0; k < na; ++k) raw += c[k] * Y[k];
        raw += 0.5f;
        const float g = dL_drgb[i * 3 + ch] * ((raw > 0.0f) ? 1.0f : 0.0f);
        for (int k = 0; k < na; ++k) o[k] = g * Y[k];
        for (int k = na; k < C; ++k) o[k] = 0.0f;
    }
}

}  // namespace cugs

// ================================================================================================
// C ABI
// ================================================================================================
using namespace cugs;

static int check_view(cugs_handle_t* h, const cugs_view_t* v) {
    CUGS_REQUIRE(h, v != nullptr, "view is null");
    CUGS_REQUIRE(h, v->width > 0 && v->height > 0, "image size must be positive");
    CUGS_REQUIRE(h, v->active_sh_degree >= 0 && v->active_sh_degree <= 3, "SH degree must be 0..3");
    const int need = (v->active_sh_degree + 1) * (v->active_sh_degree + 1);
    CUGS_REQUIRE(h, v->num_coeffs >= need, "num_coeffs < (degree+1)^2");
    return CUGS_OK;
}

int cugs_preprocess_fwd_launch(cugs_handle_t* h, void* stream, int64_t n, const cugs_view_t* v,
                               const float* positions, const float* rotations, const float* scales,
                               const float* opacities, const float* sh_coeffs, float* means_2d, float* depths,
                               float* cov_2d_inv, int32_t* radii, int32_t* tiles_touched, float* rgb,
                               float* opacities_act, float* packed, uint32_t* depth_minmax, uint64_t* gsort);

extern "C" int cugs_b200_preprocess_fwd(cugs_handle_t* h, void* stream, int64_t n,
                                        const cugs_view_t* v, const float* positions,
                                        const float* rotations, const float* scales,
                                        const float* opacities, const float* sh_coeffs,
                                        float* means_2d, float* depths, float* cov_2d_inv,
                                        int32_t* radii, int32_t* tiles_touched, float* rgb,
                                        float* opacities_act, float* packed, uint32_t* depth_minmax) {
    CUGS_REQUIRE(h, h != nullptr, "handle is null");
    CUGS_REQUIRE(h, n == 0 || (depths && cov_2d_inv && rgb && opacities_act), "null output");
    return cugs_preprocess_fwd_launch(h, stream, n, v, positions, rotations, scales, opacities, sh_coeffs, means_2d,
                                      depths, cov_2d_inv, radii, tiles_touched, rgb, opacities_act, packed,
                                      depth_minmax, nullptr);
}

int cugs_preprocess_fwd_launch(cugs_handle_t* h, void* stream, int64_t n, const cugs_view_t* v,
                               const float* positions, const float* rotations, const float* scales,
                               const float* opacities, const float* sh_coeffs, float* means_2d, float* depths,
                               float* cov_2d_inv, int32_t* radii, int32_t* tiles_touched, float* rgb,
                               float* opacities_act, float* packed, uint32_t* depth_minmax, uint64_t* gsort) {
    CUGS_REQUIRE(h, h != nullptr, "handle is null");
    CUGS_REQUIRE(h, n >= 0, "n must be >= 0");
    if (int e = check_view(h, v)) return e;
    if (n == 0) return CUGS_OK;
    CUGS_REQUIRE(h, positions && rotations && scales && opacities && sh_coeffs, "null input");
    CUGS_REQUIRE(h, means_2d && radii && tiles_touched, "null output");
    // render-only frames (cugs_b200_render_plan): the four backward-only arrays are all null
    const bool all = depths && cov_2d_inv && rgb && opacities_act;
    const bool none = !depths && !cov_2d_inv && !rgb && !opacities_act;
    CUGS_REQUIRE(h, all || (none && packed != nullptr), "null output");
    const ViewParams vp = make_view_params(v);
    const unsigned grid = (unsigned)((n + kPreBlock - 1) / kPreBlock);
    cudaStream_t s = (cudaStream_t)stream;
#define CUGS_LAUNCH_PRE(VEC, FULL, DEG3)                                                                   \
    k_preprocess_fwd<VEC, FULL, DEG3><<<grid, kPreBlock, 0, s>>>(                                          \
        n, vp, positions, rotations, scales, opacities, sh_coeffs, means_2d, depths, cov_2d_inv, radii,    \
        tiles_touched, rgb, opacities_act, reinterpret_cast<float4*>(packed), depth_minmax, gsort)
    if (v->num_coeffs == 16 && v->active_sh_degree == 3) {
        if (all) CUGS_LAUNCH_PRE(true, true, true); else CUGS_LAUNCH_PRE(true, false, true);
    } else if (v->num_coeffs == 16) {
        if (all) CUGS_LAUNCH_PRE(true, true, false); else CUGS_LAUNCH_PRE(true, false, false);
    } else {
        if (all) CUGS_LAUNCH_PRE(false, true, false); else CUGS_LAUNCH_PRE(false, false, false);
    }
#undef CUGS_LAUNCH_PRE
    CUGS_LAUNCH_CHECK(h, "k_preprocess_fwd");
    return CUGS_OK;
}

// internal launcher shared by the stage entry point and render_backward (packed gradients)
int cugs_preprocess_bwd_launch(cugs_handle_t* h, cudaStream_t s, int64_t n, const cugs_view_t* v,
                               const float* positions, const float* rotations, const float* scales,
                               const float* opacities, const float* sh_coeffs, const int32_t* radii,
                               const float* rgb, const float* dL_dmeans_2d, const float* dL_dcov_2d_inv,
                               const float* dL_drgb, const float* dL_dopacity_act, float* dL_dpositions,
                               float* dL_drotations, float* dL_dscales, float* dL_dopacities,
                               float* dL_dsh_coeffs, float* grad_accum, float* grad_count,
                               float* max_radii, const float* grad_acc, float* dL_dmeans_2d_out,
                               bool accumulate, int32_t* touch_mask, bool sparse_rows, int32_t* list,
                               int32_t* list_count, int phase /* 0 = all, 1 = classify only, 2 = chain only (list mode) */) {
    const ViewParams vp = make_view_params(v);
    const unsigned grid = (unsigned)((n + kPreBlock - 1) / kPreBlock);
    if (list != nullptr) {
        // two-phase sparse backward: classify all N, then the chain rule on the compact list only (the second
        // launch is sized for N; its warps beyond the device-side count leave at once)
        if (phase != 2) {
            CUGS_CUDA_TRY(h, cudaMemsetAsync(list_count, 0, sizeof(int32_t), s));
            int64_t cb = (n + kClassifyChunk - 1) / kClassifyChunk;
            if (cb > (int64_t)h->sm_count * 4) cb = (int64_t)h->sm_count * 4;
            k_bwd_classify<<<(unsigned)cb, kClassifyThreads, 0, s>>>(
                n, v->num_coeffs, reinterpret_cast<const float4*>(grad_acc), radii, touch_mask, accumulate,
                dL_dmeans_2d_out, grad_accum, grad_count, max_radii, dL_dpositions, dL_drotations, dL_dscales,
                dL_dopacities, dL_dsh_coeffs, list, list_count);
            CUGS_LAUNCH_CHECK(h, "k_bwd_classify");
        }
        if (phase == 1) return CUGS_OK;
        touch_mask = nullptr; grad_accum = grad_count = max_radii = nullptr; dL_dmeans_2d_out = nullptr;
        sparse_rows = false;
    }
    if (v->num_coeffs == 16 && v->active_sh_degree == 3)
        k_preprocess_bwd<true, true><<<grid, kPreBlock, 0, s>>>(
            n, vp, positions, rotations, scales, opacities, sh_coeffs, radii, rgb, dL_dmeans_2d,
            dL_dcov_2d_inv, dL_drgb, dL_dopacity_act, dL_dpositions, dL_drotations, dL_dscales,
            dL_dopacities, dL_dsh_coeffs, grad_accum, grad_count, max_radii,
            reinterpret_cast<const float4*>(grad_acc), dL_dmeans_2d_out, accumulate, touch_mask,
            sparse_rows && touch_mask != nullptr, list, list_count);
    else if (v->num_coeffs == 16)
        k_preprocess_bwd<true, false><<<grid, kPreBlock, 0, s>>>(
            n, vp, positions, rotations, scales, opacities, sh_coeffs, radii, rgb, dL_dmeans_2d,
            dL_dcov_2d_inv, dL_drgb, dL_dopacity_act, dL_dpositions, dL_drotations, dL_dscales,
            dL_dopacities, dL_dsh_coeffs, grad_accum, grad_count, max_radii,
            reinterpret_cast<const float4*>(grad_acc), dL_dmeans_2d_out, accumulate, touch_mask,
            sparse_rows && touch_mask != nullptr, list, list_count);
    else
        k_preprocess_bwd<false, false><<<grid, kPreBlock, 0, s>>>(
            n, vp, positions, rotations, scales, opacities, sh_coeffs, radii, rgb, dL_dmeans_2d,
            dL_dcov_2d_inv, dL_drgb, dL_dopacity_act, dL_dpositions, dL_drotations, dL_dscales,
            dL_dopacities, dL_dsh_coeffs, grad_accum, grad_count, max_radii,
            reinterpret_cast<const float4*>(grad_acc), dL_dmeans_2d_out, accumulate, touch_mask,
            sparse_rows && touch_mask != nullptr, list, list_count);
    CUGS_LAUNCH_CHECK(h, "k_preprocess_bwd");
    return CUGS_OK;
}

extern "C" int cugs_b200_preprocess_bwd(cugs_handle_t* h, void* stream, int64_t n,
                                        const cugs_view_t* v, const float* positions,
                                        const float* rotations, const float* scales,
                                        const float* opacities, const float* sh_coeffs,
                                        const int32_t* radii, const float* rgb,
                                        const float* dL_dmeans_2d, const float* dL_dcov_2d_inv,
                                        const float* dL_drgb, const float* dL_dopacity_act,
                                        float* dL_dpositions, float* dL_drotations,
                                        float* dL_dscales, float* dL_dopacities,
                                        float* dL_dsh_coeffs, float* grad_accum, float* grad_count,
                                        float* max_radii) {
    CUGS_REQUIRE(h, h != nullptr, "handle is null");
    CUGS_REQUIRE(h, n >= 0, "n must be >= 0");
    if (int e = check_view(h, v)) return e;
    if (n == 0) return CUGS_OK;
    CUGS_REQUIRE(h, positions && rotations && scales && opacities && sh_coeffs && radii && rgb,
                 "null input");
    CUGS_REQUIRE(h, dL_dmeans_2d && dL_dcov_2d_inv && dL_drgb && dL_dopacity_act, "null gradient input");
    CUGS_REQUIRE(h, dL_dpositions && dL_drotations && dL_dscales && dL_dopacities && dL_dsh_coeffs,
                 "null output");
    const bool any_stats = grad_accum || grad_count || max_radii;
    CUGS_REQUIRE(h, !any_stats || (grad_accum && grad_count && max_radii),
                 "stats pointers must be all set or all null");
    return cugs_preprocess_bwd_launch(h, (cudaStream_t)stream, n, v, positions, rotations, scales,
                                      opacities, sh_coeffs, radii, rgb, dL_dmeans_2d, dL_dcov_2d_inv,
                                      dL_drgb, dL_dopacity_act, dL_dpositions, dL_drotations, dL_dscales,
                                      dL_dopacities, dL_dsh_coeffs, grad_accum, grad_count, max_radii,
                                      nullptr, nullptr, false, nullptr, false, nullptr, nullptr, 0);
}

extern "C" int cugs_b200_sh_forward(cugs_handle_t* h, void* stream, int64_t n, int degree,
                                    int num_coeffs, const float* sh_coeffs, const float* directions,
                                    float* rgb) {
    CUGS_REQUIRE(h, h != nullptr, "handle is null");
    CUGS_REQUIRE(h, degree >= 0 && degree <= 3, "SH degree must be 0..3");
    CUGS_REQUIRE(h, num_coeffs >= (degree + 1) * (degree + 1), "num_coeffs < (degree+1)^2");
    CUGS_REQUIRE(h, n >= 0, "n must be >= 0");
    if (n == 0) return CUGS_OK;
    CUGS_REQUIRE(h, sh_coeffs && directions && rgb, "null pointer");
    k_sh_forward<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        n, degree, num_coeffs, sh_coeffs, directions, rgb);
    CUGS_LAUNCH_CHECK(h, "k_sh_forward");
    return CUGS_OK;
}

extern "C" int cugs_b200_sh_backward(cugs_handle_t* h, void* stream, int64_t n, int degree,
                                     int num_coeffs, const float* sh_coeffs, const float* directions,
                                     const float* dL_drgb, float* dL_dsh) {
    CUGS_REQUIRE(h, h != nullptr, "handle is null");
    CUGS_REQUIRE(h, degree >= 0 && degree <= 3, "SH degree must be 0..3");
    CUGS_REQUIRE(h, num_coeffs >= (degree + 1) * (degree + 1), "num_coeffs < (degree+1)^2");
    CUGS_REQUIRE(h, n >= 0, "n must be >= 0");
    if (n == 0) return CUGS_OK;
    CUGS_REQUIRE(h, sh_coeffs && directions && dL_drgb && dL_dsh, "null pointer");
    k_sh_backward<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        n, degree, num_coeffs, sh_coeffs, directions, dL_drgb, dL_dsh);
    CUGS_LAUNCH_CHECK(h, "k_sh_backward");
    return CUGS_OK;
}
