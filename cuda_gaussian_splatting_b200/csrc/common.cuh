// common.cuh — shared device/host helpers for the sm_100a rasterizer kernels.
// No torch headers anywhere under csrc/ (keeps each TU at seconds of nvcc time).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/cugs_b200.h"

struct cugs_handle {
    int device;
    int sm_count;
    int64_t* pinned;      // mapped pinned host words: [0,16) fixed slots (4..6 densify / relocate counts),
                          // [16, 16 + kPinnedRing) a ring of one-shot slots for the blocking count reads
    unsigned pinned_seq;  // next ring slot
    char err[512];
    // optional per-stage timing of the fused entry points (cugs_b200_set_stage_timing)
    bool timing;
    cudaEvent_t ev[12];
    bool ev_recorded[12];
    int last_sort_passes, last_sort_key_bits;  // plan of the most recent render_finish
    unsigned long long launches;  // kernels launched through this handle (cugs_b200_launch_count)
};

namespace cugs {

constexpr int kPinnedWords = 80;
constexpr int kPinnedRing = 64;
// one-shot pinned word for a blocking device->host count (render_plan, scan): the ring is long enough
// for every frame that can be in flight through one handle
inline int64_t* cugs_pinned_slot(cugs_handle* h) { return h->pinned + 16 + (h->pinned_seq++ % kPinnedRing); }

constexpr int kTile = CUGS_TILE;
constexpr unsigned kFull = 0xffffffffu;

inline int set_error(cugs_handle* h, int code, const char* fmt, ...) {
    if (h) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(h->err, sizeof(h->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

#define CUGS_CUDA_TRY(h, expr)                                                                  \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return cugs::set_error((h), (int)_e, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,   \
                                   cudaGetErrorString(_e));                                     \
    } while (0)

#define CUGS_LAUNCH_CHECK(h, name)                                                              \
    do {                                                                                        \
        cudaError_t _e = cudaGetLastError();                                                    \
        if (_e != cudaSuccess)                                                                  \
            return cugs::set_error((h), (int)_e, "launch %s failed: %s", (name),                \
                                   cudaGetErrorString(_e));                                     \
        ++(h)->launches;                                                                        \
    } while (0)

#define CUGS_REQUIRE(h, cond, msg)                                                              \
    do {                                                                                        \
        if (!(cond)) return cugs::set_error((h), CUGS_ERR_INVALID_ARG, "%s (%s)", (msg), #cond); \
    } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- explicit-rounding helpers: spell out the FMA contraction nvcc applies to the reference's
// expressions so integer outputs derived from float math stay bit-identical. ----
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fma_rn(float a, float b, float c) { return __fmaf_rn(a, b, c); }
// a*b + c*d  ->  fma(a, b, rn(c*d))
__device__ __forceinline__ float dot2c(float a, float b, float c, float d) {
    return __fmaf_rn(a, b, __fmul_rn(c, d));
}
// a*b + c*d + e*f  ->  fma(e, f, fma(a, b, rn(c*d)))
__device__ __forceinline__ float dot3c(float a, float b, float c, float d, float e, float f) {
    return __fmaf_rn(e, f, __fmaf_rn(a, b, __fmul_rn(c, d)));
}

// Lanes of the warp holding the same `bits`-bit digit. A ballot per bit: match.any.sync is several
// times slower when the warp holds many distinct digits (ncu: 45 % of the onesweep kernel's stall
// samples sat on MATCH.ANY results).
__device__ __forceinline__ unsigned match_digit(unsigned d, int bits) {
    unsigned peers = kFull;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        if (b < bits) {
            const bool set = (d >> b) & 1u;
            const unsigned m = __ballot_sync(kFull, set);
            peers &= set ? m : ~m;
        }
    }
    return peers;
}
// same with the digit width known at compile time. The digit's bits go to predicates (ptxas: one R2P),
// each bit is balloted, and the ballot is folded under that predicate into one of two accumulators:
// `same` = AND of the ballots of the bits this lane has set, `diff` = OR of the ballots of the bits it has
// clear; peers = same & ~diff. 3 instructions per bit (vote + two predicated lop3) instead of the 6
// (shift, mask, compare, decrement, vote, lop3) nvcc emits for the C++ form.
template <int kBits>
__device__ __forceinline__ unsigned match_digit_fixed(unsigned d) {
    unsigned same = kFull, diff = 0u;
#pragma unroll
    for (int b = 0; b < kBits; ++b) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            ".reg .b32 t, m;\n\t"
            "and.b32 t, %2, %3;\n\t"
            "setp.ne.u32 p, t, 0;\n\t"
            "vote.sync.ballot.b32 m, p, 0xffffffff;\n\t"
            "@p and.b32 %0, %0, m;\n\t"
            "@!p or.b32 %1, %1, m;\n\t"
            "}"
            : "+r"(same), "+r"(diff)
            : "r"(d), "r"(1u << b));
    }
    return same & ~diff;
}

// Tile rectangle of a projected Gaussian, shared by preprocess (projection.cu:172-188) and key
// emission (sorting.cu:52-57). (int) casts are cvt.rzi (truncate, saturate, NaN -> 0).
struct TileRect {
    int tx0, ty0, tx1, ty1;
};
__device__ __forceinline__ TileRect tile_rect(float x, float y, int radius, int w, int h, int ntx,
                                              int nty) {
    const float r = (float)radius;
    const int rminx = max(0, (int)(x - r));
    const int rminy = max(0, (int)(y - r));
    const int rmaxx = min(w, (int)(x + r + 1.0f));
    const int rmaxy = min(h, (int)(y + r + 1.0f));
    TileRect t;
    t.tx0 = rminx / kTile;
    t.ty0 = rminy / kTile;
    t.tx1 = min(ntx, (rmaxx + kTile - 1) / kTile);
    t.ty1 = min(nty, (rmaxy + kTile - 1) / kTile);
    return t;
}

// SH basis Y_k(dir), constants and signs as core/sh.cu:44-74. Entries >= (deg+1)^2 are 0.
__device__ __forceinline__ void sh_basis(int deg, float x, float y, float z, float Y[16]) {
#pragma unroll
    for (int k = 0; k < 16; ++k) Y[k] = 0.0f;
    Y[0] = 0.28209479177387814f;
    if (deg >= 1) {
        Y[1] = -0.4886025119029199f * y;
        Y[2] = 0.4886025119029199f * z;
        Y[3] = -0.4886025119029199f * x;
    }
    if (deg >= 2) {
        const float xx = x * x, yy = y * y, zz = z * z;
        Y[4] = 1.0925484305920792f * (x * y);
        Y[5] = 1.0925484305920792f * (y * z);
        Y[6] = 0.31539156525252005f * (2.0f * zz - xx - yy);
        Y[7] = 1.0925484305920792f * (x * z);
        Y[8] = 0.5462742152960396f * (xx - yy);
        if (deg >= 3) {
            Y[9] = 0.5900435899266435f * y * (3.0f * xx - yy);
            Y[10] = 2.890611442640554f * x * y * z;
            Y[11] = 0.4570457994644658f * y * (4.0f * zz - xx - yy);
            Y[12] = 0.3731763325901154f * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
            Y[13] = 0.4570457994644658f * x * (4.0f * zz - xx - yy);
            Y[14] = 1.4453057213202769f * z * (xx - yy);
            Y[15] = 0.5900435899266435f * x * (xx - 3.0f * yy);
        }
    }
}

// Normalised view direction (projection.cu:273-280: (p - c) / clamp_min(||p - c||, 1e-8)).
__device__ __forceinline__ void view_dir(float px, float py, float pz, const float c[3], float& dx,
                                         float& dy, float& dz) {
    dx = px - c[0];
    dy = py - c[1];
    dz = pz - c[2];
    const float n = fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-8f);
    dx = dx / n;
    dy = dy / n;
    dz = dz / n;
}

// Conservative early-reject threshold for the blend kernels: alpha = op * exp(power) can reach
// 1/255 only if power >= -log(255 * op). Evaluations with power below (that - 1e-4) are rejected
// without evaluating exp; everything else takes the exact path (accurate expf, the reference's
// own comparisons), so n_contrib / final_T decisions stay bit-identical. The 1e-4 guard band is
// ~300x the combined rounding error of __logf, expf and the product. NaN opacity -> -inf
// (always exact path); op <= 0 -> +1 (never contributes, as in the reference).
__device__ __forceinline__ float blend_reject_threshold(float op) {
    if (op != op) return -INFINITY;
    if (!(op > 0.0f)) return 1.0f;
    return -__logf(255.0f * op) - 1e-4f;
}

// Conservative axis-aligned half-extents of the region where a Gaussian can pass the cheap reject
// (power >= thr): the ellipse d^T Q d <= t, t = -2 thr, Q = Sigma'^-1, spans |dx| <= sqrt(t Sigma'_xx),
// |dy| <= sqrt(t Sigma'_yy). Inflated by 1e-4 relative + 1e-3 px against rounding. The blend kernels
// drop a Gaussian from a warp's list when this box misses the warp's pixel patch, which cannot
// change any result. thr > 0 (opacity below 1/255): nothing can pass -> -inf (always dropped);
// NaN inputs give NaN extents, whose comparisons are false -> never dropped (exact path decides).
__device__ __forceinline__ void blend_extents(float thr, float sxx, float syy, float& hx, float& hy) {
    const float t = -2.0f * thr;
    if (t < 0.0f) { hx = -INFINITY; hy = -INFINITY; return; }
    hx = sqrtf(t * sxx) * 1.0001f + 1e-3f;
    hy = sqrtf(t * syy) * 1.0001f + 1e-3f;
}

// Per-launch constants shared by the per-Gaussian kernels (passed by value: lives in the
// constant bank, so the 16-entry view matrix is not re-read from global by every thread).
struct ViewParams {
    float W[9];       // rotation rows of the row-major view matrix
    float t[3];       // view[3], view[7], view[11]
    float cam[3];
    float fx, fy, cx, cy;
    float scale_mod;  // kernels add logf(scale_mod + 1e-8f) on the device (projection.cu:126-130)
    int width, height, ntx, nty;
    int deg, C;
};

inline ViewParams make_view_params(const cugs_view_t* v) {
    ViewParams p;
    p.W[0] = v->view[0]; p.W[1] = v->view[1]; p.W[2] = v->view[2];
    p.W[3] = v->view[4]; p.W[4] = v->view[5]; p.W[5] = v->view[6];
    p.W[6] = v->view[8]; p.W[7] = v->view[9]; p.W[8] = v->view[10];
    p.t[0] = v->view[3]; p.t[1] = v->view[7]; p.t[2] = v->view[11];
    p.cam[0] = v->cam_center[0]; p.cam[1] = v->cam_center[1]; p.cam[2] = v->cam_center[2];
    p.fx = v->fx; p.fy = v->fy; p.cx = v->cx; p.cy = v->cy;
    p.scale_mod = v->scale_modifier;
    p.width = v->width; p.height = v->height;
    p.ntx = (v->width + kTile - 1) / kTile;
    p.nty = (v->height + kTile - 1) / kTile;
    p.deg = v->active_sh_degree;
    p.C = v->num_coeffs;
    return p;
}

// Per-step scalars of the training step that change every iteration (learning rates, Adam bias
// corrections, noise scale, step number), kept in DEVICE memory so that a captured CUDA graph of the
// step can be replayed unchanged: the host refreshes this struct with one small H2D copy before each
// launch. ok = 0 (a frame of the step overflowed its pair capacity) turns the update kernels into no-ops:
// the step is transactional, the host re-runs it with larger buffers.
struct StepDyn {
    float lr[5];
    float bc1, bc2;
    float noise_lr;
    unsigned step;
    int ok;
    int pad[2];
};

// ---- counter-based random numbers (MCMC noise, split / relocation jitter) ----------------------
// Philox-4x32-10; the caller chooses key = seed and a counter that names the draw (Gaussian index,
// step, purpose), so results do not depend on the launch shape and every rank of a view-parallel
// run draws the same numbers.
__device__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0,
                                              unsigned k1, unsigned (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const unsigned n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// three standard normals (Box-Muller on 24-bit uniforms in (0,1)) and one spare uniform
__device__ __forceinline__ void philox_normal3(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0,
                                               unsigned k1, float& z0, float& z1, float& z2,
                                               float* spare_uniform = nullptr) {
    unsigned r[4];
    philox4x32_10(c0, c1, c2, c3, k0, k1, r);
    const float u0 = ((float)(r[0] >> 8) + 0.5f) * 5.9604644775390625e-08f;
    const float u1 = ((float)(r[1] >> 8) + 0.5f) * 5.9604644775390625e-08f;
    const float u2 = ((float)(r[2] >> 8) + 0.5f) * 5.9604644775390625e-08f;
    const float u3 = ((float)(r[3] >> 8) + 0.5f) * 5.9604644775390625e-08f;
    const float ra = sqrtf(-2.0f * logf(u0)), rb = sqrtf(-2.0f * logf(u2));
    float s0, cs0, s1, cs1;
    sincospif(2.0f * u1, &s0, &cs0);
    sincospif(2.0f * u3, &s1, &cs1);
    z0 = ra * cs0; z1 = ra * s0; z2 = rb * cs1;
    if (spare_uniform) *spare_uniform = u3;
}

}  // namespace cugs
