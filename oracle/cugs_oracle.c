/* oracle/cugs_oracle.c — CPU restatement of the reference rasterizer hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE. Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library. The product
 * (cuda_gaussian_splatting_b200/) never links or calls it and has no CPU fallback.
 *
 * Parity status: PINNED. The reference ships no golden files; this oracle is
 * checked, in the CPU-only suite, against
 *  (a) the known-answer tests of the reference's own test-suite, ported one by one in
 *      tests/test_oracle_known_answers.py;
 *  (b) tests/golden/ref_gpu_golden.npz: stage outputs, images, gradients, loss and three
 *      FusedAdam steps of the UNMODIFIED reference CUDA kernels (oracle/_ref, built from
 *      /root/reference by oracle/Makefile.ref) recorded on a B200 by
 *      tests/golden/make_golden_gpu.py -- tests/test_oracle_vs_reference_golden.py: radii, tile
 *      counts, sorted keys, sort order, tile ranges, n_contrib and Adam bit-exact, floats <= 2e-6;
 *  (c) tests/golden/sh_cpu_golden.npz (the reference's evaluate_sh_cpu, make_golden_cpu.py) and
 *      the PLY files written by the reference (make_golden_ply.py);
 * and, on the GPU box, against oracle/_ref live (tests/test_gpu_parity.py).
 *
 * Every function cites the reference file:line (relative to /root/reference/src) it restates.
 * Build: gcc -O2 -ffp-contract=off -fopenmp (see oracle/Makefile). -ffp-contract=off matters:
 * every fused multiply-add below is written out with fmaf() in the order nvcc 12.9 contracts
 * the reference's expressions, so that integer outputs derived from float
 * math (radii, tile counts, depth key bits) agree with the GPU except where libm and
 * libdevice differ in the last ulp (expf, rsqrtf, logf).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define TILE 16

/* ---- nvcc contraction model (A.10): a*b + c*d -> fma(a,b,c*d);
 *      a*b + c*d + e*f -> fma(e,f,fma(a,b,c*d)) ---- */
static inline float dot2c(float a, float b, float c, float d) { return fmaf(a, b, c * d); }
static inline float dot3c(float a, float b, float c, float d, float e, float f) {
    return fmaf(e, f, fmaf(a, b, c * d));
}

/* CUDA static_cast<int>(float): cvt.rzi.s32.f32 — truncates, saturates, NaN -> 0. */
static inline int f2i_rz(float x) {
    if (x != x) return 0;
    if (x >= 2147483648.0f) return 2147483647;
    if (x <= -2147483648.0f) return (-2147483647 - 1);
    return (int)x;
}
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * Shared forward math: rasterizer/projection.cuh:29-226
 * ---------------------------------------------------------------------------------------- */

/* projection.cuh:29-49 quat_to_rotation (w,x,y,z), row-major R. */
static void quat_to_rotation(float w, float x, float y, float z, float R[9], float* inv_norm_out) {
    float n2 = fmaf(z, z, fmaf(y, y, fmaf(w, w, x * x))) + 1e-12f;
    float inv_norm = 1.0f / sqrtf(n2); /* GPU: rsqrtf (approx, <=1ulp off) */
    w *= inv_norm; x *= inv_norm; y *= inv_norm; z *= inv_norm;
    if (inv_norm_out) *inv_norm_out = inv_norm;
    /* nvcc shares y*y and z*z with R[4], R[8]: here both products are rounded (sm_100 SASS) */
    R[0] = fmaf(-2.0f, y * y + z * z, 1.0f);
    R[1] = 2.0f * fmaf(x, y, -(w * z));
    R[2] = 2.0f * dot2c(x, z, w, y);
    R[3] = 2.0f * dot2c(x, y, w, z);
    R[4] = fmaf(-2.0f, dot2c(x, x, z, z), 1.0f);
    R[5] = 2.0f * fmaf(y, z, -(w * x));
    R[6] = 2.0f * fmaf(x, z, -(w * y));
    R[7] = 2.0f * dot2c(y, z, w, x);
    R[8] = fmaf(-2.0f, dot2c(x, x, y, y), 1.0f);
}

/* projection.cuh:66-90 compute_cov_3d: M = R diag(exp(log_scale)), Sigma = M M^T (upper tri). */
static void compute_cov3d(const float ls[3], const float q[4], float cov[6], float M[9], float R[9],
                          float s[3], float* inv_norm) {
    s[0] = expf(ls[0]); s[1] = expf(ls[1]); s[2] = expf(ls[2]);
    quat_to_rotation(q[0], q[1], q[2], q[3], R, inv_norm);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[i * 3 + j] = R[i * 3 + j] * s[j];
    cov[0] = dot3c(M[1], M[1], M[0], M[0], M[2], M[2]); /* M0*M0 is the rounded product (SASS) */
    cov[1] = dot3c(M[0], M[3], M[1], M[4], M[2], M[5]);
    cov[2] = dot3c(M[1], M[7], M[0], M[6], M[2], M[8]); /* M0*M6 is the rounded product (SASS) */
    cov[3] = dot3c(M[3], M[3], M[4], M[4], M[5], M[5]);
    cov[4] = dot3c(M[3], M[6], M[4], M[7], M[5], M[8]);
    cov[5] = dot3c(M[6], M[6], M[7], M[7], M[8], M[8]);
}

/* projection.cuh:114-165 compute_cov_2d: Sigma' = (J W) Sigma (J W)^T + 0.3 I.
 * Tm (2x3) is returned for the backward pass. J[1] = J[3] = 0 terms are kept as in the
 * reference: a*b + 0*c + e*f contracts to fma(e, f, fma(a, b, 0*c)) = fma(e, f, rn(a*b)). */
static void compute_cov2d(const float S[6], const float W[9], const float t[3], float fx, float fy,
                          float cov2d[3], float Tm[6]) {
    float tx = t[0], ty = t[1], tz = t[2];
    float tz_inv = 1.0f / (tz + 1e-6f);
    float tz_inv2 = tz_inv * tz_inv;
    float J0 = fx * tz_inv, J2 = -fx * tx * tz_inv2;
    float J4 = fy * tz_inv, J5 = -fy * ty * tz_inv2;
    Tm[0] = fmaf(J2, W[6], fmaf(J0, W[0], 0.0f * W[3]));
    Tm[1] = fmaf(J2, W[7], fmaf(J0, W[1], 0.0f * W[4]));
    Tm[2] = fmaf(J2, W[8], fmaf(J0, W[2], 0.0f * W[5]));
    Tm[3] = fmaf(J5, W[6], fmaf(0.0f, W[0], J4 * W[3]));
    Tm[4] = fmaf(J5, W[7], fmaf(0.0f, W[1], J4 * W[4]));
    Tm[5] = fmaf(J5, W[8], fmaf(0.0f, W[2], J4 * W[5]));
    float TS[6];
    TS[0] = dot3c(Tm[0], S[0], Tm[1], S[1], Tm[2], S[2]);
    TS[1] = dot3c(Tm[0], S[1], Tm[1], S[3], Tm[2], S[4]);
    TS[2] = dot3c(Tm[0], S[2], Tm[1], S[4], Tm[2], S[5]);
    TS[3] = dot3c(Tm[3], S[0], Tm[4], S[1], Tm[5], S[2]);
    TS[4] = dot3c(Tm[3], S[1], Tm[4], S[3], Tm[5], S[4]);
    TS[5] = dot3c(Tm[3], S[2], Tm[4], S[4], Tm[5], S[5]);
    cov2d[0] = dot3c(TS[0], Tm[0], TS[1], Tm[1], TS[2], Tm[2]) + 0.3f;
    cov2d[1] = dot3c(TS[0], Tm[3], TS[1], Tm[4], TS[2], Tm[5]);
    cov2d[2] = dot3c(TS[3], Tm[3], TS[4], Tm[4], TS[5], Tm[5]) + 0.3f;
}

/* projection.cuh:179-195 compute_radius. */
static int compute_radius(const float c2[3]) {
    float a = c2[0], b = c2[1], c = c2[2];
    float det = fmaf(a, c, -(b * b));
    float trace = a + c;
    float disc = fmaxf(fmaf(trace, trace, -(4.0f * det)), 0.0f);
    float lambda_max = 0.5f * (trace + sqrtf(disc));
    if (lambda_max <= 0.0f) return 0;
    return f2i_rz(ceilf(3.0f * sqrtf(lambda_max)));
}

/* Tile rectangle shared by projection.cu:172-188 and sorting.cu:52-57. */
static void tile_rect(float x, float y, int radius, int W, int H, int ntx, int nty, int* tx0,
                      int* ty0, int* tx1, int* ty1) {
    float r = (float)radius;
    int rminx = imax(0, f2i_rz(x - r));
    int rminy = imax(0, f2i_rz(y - r));
    int rmaxx = imin(W, f2i_rz(x + r + 1.0f));
    int rmaxy = imin(H, f2i_rz(y + r + 1.0f));
    *tx0 = rminx / TILE;
    *ty0 = rminy / TILE;
    *tx1 = imin(ntx, (rmaxx + TILE - 1) / TILE);
    *ty1 = imin(nty, (rmaxy + TILE - 1) / TILE);
}

/* SH basis, signs and constants exactly as core/sh.cu:44-74 / core/sh_backward.cu:50-82. */
static void sh_basis(int deg, float x, float y, float z, float Y[16]) {
    for (int k = 0; k < 16; ++k) Y[k] = 0.0f;
    Y[0] = 0.28209479177387814f;
    if (deg >= 1) {
        Y[1] = -0.4886025119029199f * y;
        Y[2] = 0.4886025119029199f * z;
        Y[3] = -0.4886025119029199f * x;
    }
    if (deg >= 2) {
        float xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z;
        Y[4] = 1.0925484305920792f * xy;
        Y[5] = 1.0925484305920792f * yz;
        Y[6] = 0.31539156525252005f * (2.0f * zz - xx - yy);
        Y[7] = 1.0925484305920792f * xz;
        Y[8] = 0.5462742152960396f * (xx - yy);
    }
    if (deg >= 3) {
        float xx = x * x, yy = y * y, zz = z * z;
        Y[9] = 0.5900435899266435f * y * (3.0f * xx - yy);
        Y[10] = 2.890611442640554f * x * y * z;
        Y[11] = 0.4570457994644658f * y * (4.0f * zz - xx - yy);
        Y[12] = 0.3731763325901154f * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
        Y[13] = 0.4570457994644658f * x * (4.0f * zz - xx - yy);
        Y[14] = 1.4453057213202769f * z * (xx - yy);
        Y[15] = 0.5900435899266435f * x * (xx - 3.0f * yy);
    }
}

/* View direction: projection.cu:273-280 (libtorch: (p - c) / max(||p - c||, 1e-8)). */
static void view_dir(const float p[3], const float c[3], float d[3]) {
    float dx = p[0] - c[0], dy = p[1] - c[1], dz = p[2] - c[2];
    float n = sqrtf(dx * dx + dy * dy + dz * dz);
    if (n < 1e-8f) n = 1e-8f;
    d[0] = dx / n; d[1] = dy / n; d[2] = dz / n;
}

/* core/sh.cu:19-79 k_evaluate_sh (+0.5) followed by clamp_min(0) (projection.cu:284). */
void oracle_sh_forward(int64_t n, int deg, int C, const float* sh, const float* dirs, float* rgb,
                       int clamp) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        float Y[16];
        sh_basis(deg, dirs[i * 3 + 0], dirs[i * 3 + 1], dirs[i * 3 + 2], Y);
        int na = (deg + 1) * (deg + 1);
        for (int ch = 0; ch < 3; ++ch) {
            const float* c = sh + (i * 3 + ch) * C;
            float col = 0.0f;
            for (int k = 0; k < na; ++k) col += c[k] * Y[k];
            col += 0.5f;
            if (clamp && col < 0.0f) col = 0.0f;
            rgb[i * 3 + ch] = col;
        }
    }
}

/* core/sh_backward.cu:29-112 k_evaluate_sh_backward. */
void oracle_sh_backward(int64_t n, int deg, int C, const float* sh, const float* dirs,
                        const float* dL_drgb, float* dL_dsh) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        float Y[16];
        sh_basis(deg, dirs[i * 3 + 0], dirs[i * 3 + 1], dirs[i * 3 + 2], Y);
        int na = (deg + 1) * (deg + 1);
        for (int ch = 0; ch < 3; ++ch) {
            const float* c = sh + (i * 3 + ch) * C;
            float* d = dL_dsh + (i * 3 + ch) * C;
            float raw = 0.0f;
            for (int k = 0; k < na; ++k) raw += c[k] * Y[k];
            raw += 0.5f;
            float g = dL_drgb[i * 3 + ch] * ((raw > 0.0f) ? 1.0f : 0.0f);
            for (int k = 0; k < na; ++k) d[k] = g * Y[k];
            for (int k = na; k < C; ++k) d[k] = 0.0f;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * A.1 preprocess forward: rasterizer/projection.cu:55-189 + :273-284 (+ core/sh.cu)
 * view: row-major 4x4 world->camera. All outputs fully defined (zeros where the reference
 * leaves its torch::zeros untouched).
 * ---------------------------------------------------------------------------------------- */
void oracle_preprocess_fwd(int64_t n, const float* view, float fx, float fy, float cx, float cy,
                           int W, int H, float scale_mod, const float* cam_center, int deg, int C,
                           const float* pos, const float* rot, const float* scl, const float* opa,
                           const float* sh, float* means2d, float* depths, float* cov2d_inv,
                           int* radii, int* tiles, float* rgb, float* opa_act) {
    const float Wm[9] = {view[0], view[1], view[2], view[4], view[5], view[6], view[8], view[9], view[10]};
    const int ntx = (W + TILE - 1) / TILE, nty = (H + TILE - 1) / TILE;
    const float lsm = logf(scale_mod + 1e-8f);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        radii[i] = 0; tiles[i] = 0;
        means2d[i * 2] = means2d[i * 2 + 1] = 0.0f;
        depths[i] = 0.0f; opa_act[i] = 0.0f;
        cov2d_inv[i * 3] = cov2d_inv[i * 3 + 1] = cov2d_inv[i * 3 + 2] = 0.0f;

        /* SH colour for every Gaussian, culled or not (sh.cu:27-28). */
        float d[3], Y[16];
        view_dir(pos + i * 3, cam_center, d);
        sh_basis(deg, d[0], d[1], d[2], Y);
        {
            int na = (deg + 1) * (deg + 1);
            for (int ch = 0; ch < 3; ++ch) {
                const float* c = sh + (i * 3 + ch) * C;
                float col = 0.0f;
                for (int k = 0; k < na; ++k) col += c[k] * Y[k];
                col += 0.5f;
                rgb[i * 3 + ch] = (col < 0.0f) ? 0.0f : col;
            }
        }

        float px = pos[i * 3], py = pos[i * 3 + 1], pz = pos[i * 3 + 2];
        /* A.10: t = fadd(fma(pz, W2, fma(px, W0, py*W1)), view[3]) */
        float t[3];
        t[0] = fmaf(pz, Wm[2], fmaf(px, Wm[0], py * Wm[1])) + view[3];
        t[1] = fmaf(pz, Wm[5], fmaf(px, Wm[3], py * Wm[4])) + view[7];
        t[2] = fmaf(pz, Wm[8], fmaf(px, Wm[6], py * Wm[7])) + view[11];
        if (t[2] <= 0.2f) continue; /* projection.cu:104 */

        float xs = (fx * t[0]) / t[2] + cx; /* projection.cu:109-110 */
        float ys = (fy * t[1]) / t[2] + cy;
        depths[i] = t[2];
        means2d[i * 2] = xs; means2d[i * 2 + 1] = ys;
        opa_act[i] = 1.0f / (1.0f + expf(-opa[i])); /* :119-121 */

        float ls[3] = {scl[i * 3] + lsm, scl[i * 3 + 1] + lsm, scl[i * 3 + 2] + lsm};
        float cov3[6], M[9], R[9], s[3], c2[3], Tm[6];
        compute_cov3d(ls, rot + i * 4, cov3, M, R, s, NULL);
        compute_cov2d(cov3, Wm, t, fx, fy, c2, Tm);

        float det = fmaf(c2[0], c2[2], -(c2[1] * c2[1])); /* projection.cuh:209-226 */
        if (det <= 0.0f) continue; /* projection.cu:151-152 */
        float inv_det = 1.0f / det;
        cov2d_inv[i * 3 + 0] = c2[2] * inv_det;
        cov2d_inv[i * 3 + 1] = -c2[1] * inv_det;
        cov2d_inv[i * 3 + 2] = c2[0] * inv_det;

        int radius = compute_radius(c2);
        if (radius <= 0) continue;
        radius = imin(radius, imax(W, H)); /* projection.cu:165-166 */
        radii[i] = radius;
        int tx0, ty0, tx1, ty1;
        tile_rect(xs, ys, radius, W, H, ntx, nty, &tx0, &ty0, &tx1, &ty1);
        int nt = (tx1 - tx0) * (ty1 - ty0); /* A.1-11 quirk: (-a)*(-b) > 0 is kept */
        tiles[i] = imax(nt, 0);
    }
}

/* ------------------------------------------------------------------------------------------
 * A.2 scan + key emission: rasterizer/sorting.cu:145-152, :30-72
 * ---------------------------------------------------------------------------------------- */
int64_t oracle_scan(int64_t n, const int* tiles, int* offsets) {
    int32_t acc = 0; /* int32 cumsum like sorting.cu:145 */
    for (int64_t i = 0; i < n; ++i) { offsets[i] = acc; acc += tiles[i]; }
    return (int64_t)acc;
}

void oracle_fill_keys(int64_t n, const float* means2d, const float* depths, const int* radii,
                      const int* offsets, int W, int H, int64_t P, uint64_t* keys, int* values) {
    const int ntx = (W + TILE - 1) / TILE, nty = (H + TILE - 1) / TILE;
    memset(keys, 0, (size_t)P * 8); /* sorting.cu:166-167 torch::zeros */
    memset(values, 0, (size_t)P * 4);
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t i = 0; i < n; ++i) {
        int radius = radii[i];
        if (radius <= 0) continue;
        int tx0, ty0, tx1, ty1;
        tile_rect(means2d[i * 2], means2d[i * 2 + 1], radius, W, H, ntx, nty, &tx0, &ty0, &tx1, &ty1);
        uint64_t db = (uint64_t)f2u(depths[i]);
        int64_t wp = offsets[i];
        for (int ty = ty0; ty < ty1; ++ty)
            for (int tx = tx0; tx < tx1; ++tx) {
                uint64_t tile_id = (uint64_t)(int64_t)(ty * ntx + tx);
                keys[wp] = (tile_id << 32) | db;
                values[wp] = (int)i;
                ++wp;
            }
    }
}

/* ------------------------------------------------------------------------------------------
 * A.3 sort + ranges: sorting.cu:190-211 (cub::DeviceRadixSort::SortPairs, bits [0,64), stable
 * ascending — CUB 2.8.2 from the CUDA 12.9 toolkit, not vendored in the reference; restated
 * here as a stable LSD radix sort, which yields the identical permutation) and :82-109.
 * ---------------------------------------------------------------------------------------- */
void oracle_sort_pairs(int64_t P, const uint64_t* kin, const int* vin, uint64_t* kout, int* vout) {
    if (P <= 0) return;
    uint64_t* ka = (uint64_t*)malloc((size_t)P * 8);
    uint64_t* kb = (uint64_t*)malloc((size_t)P * 8);
    int* va = (int*)malloc((size_t)P * 4);
    int* vb = (int*)malloc((size_t)P * 4);
    memcpy(ka, kin, (size_t)P * 8);
    memcpy(va, vin, (size_t)P * 4);
    for (int pass = 0; pass < 4; ++pass) { /* 4 x 16-bit digits = 64 bits */
        int shift = pass * 16;
        int64_t* cnt = (int64_t*)calloc(65537, sizeof(int64_t));
        for (int64_t i = 0; i < P; ++i) cnt[((ka[i] >> shift) & 0xFFFF) + 1]++;
        for (int d = 0; d < 65536; ++d) cnt[d + 1] += cnt[d];
        for (int64_t i = 0; i < P; ++i) {
            int64_t dst = cnt[(ka[i] >> shift) & 0xFFFF]++;
            kb[dst] = ka[i]; vb[dst] = va[i];
        }
        free(cnt);
        uint64_t* tk = ka; ka = kb; kb = tk;
        int* tv = va; va = vb; vb = tv;
    }
    memcpy(kout, ka, (size_t)P * 8);
    memcpy(vout, va, (size_t)P * 4);
    free(ka); free(kb); free(va); free(vb);
}

void oracle_tile_ranges(int64_t P, const uint64_t* keys, int num_tiles, int* ranges) {
    memset(ranges, 0, (size_t)num_tiles * 8); /* sorting.cu:216 */
    for (int64_t i = 0; i < P; ++i) {
        uint32_t cur = (uint32_t)(keys[i] >> 32);
        if (i == 0) ranges[cur * 2] = 0;
        else {
            uint32_t prev = (uint32_t)(keys[i - 1] >> 32);
            if (cur != prev) { ranges[prev * 2 + 1] = (int)i; ranges[cur * 2] = (int)i; }
        }
        if (i == P - 1) ranges[cur * 2 + 1] = (int)P;
    }
}

/* ------------------------------------------------------------------------------------------
 * A.4 blend forward: rasterizer/forward.cu:48-174. stats (optional, 3 x int64): evaluations,
 * alpha-rejected evaluations, contributing evaluations (the E_fwd work units of SURVEY §8d).
 * ---------------------------------------------------------------------------------------- */
static inline float blend_power(float dx, float dy, float a, float b, float c) {
    /* A.10: s1 = fma(dx,a,dy*b); s2 = fma(dx,b,dy*c); power = fmul(fma(dx,s1,dy*s2), -0.5) */
    float s1 = fmaf(dx, a, dy * b);
    float s2 = fmaf(dx, b, dy * c);
    return fmaf(dx, s1, dy * s2) * -0.5f;
}

void oracle_blend_fwd(int W, int H, const float* bg, const int* ranges, const int* gidx,
                      const float* means2d, const float* conic, const float* rgb, const float* opa,
                      float* color, float* final_T, int* n_contrib, int64_t* stats) {
    const int ntx = (W + TILE - 1) / TILE, nty = (H + TILE - 1) / TILE;
    int64_t ev = 0, rej = 0, con = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : ev, rej, con)
    for (int tile = 0; tile < ntx * nty; ++tile) {
        int ty = tile / ntx, tx = tile % ntx;
        int r0 = ranges[tile * 2], r1 = ranges[tile * 2 + 1];
        for (int ly = 0; ly < TILE; ++ly)
            for (int lx = 0; lx < TILE; ++lx) {
                int px = tx * TILE + lx, py = ty * TILE + ly;
                if (px >= W || py >= H) continue;
                float pxf = (float)px + 0.5f, pyf = (float)py + 0.5f;
                float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
                int cnt = 0;
                for (int k = r0; k < r1; ++k) {
                    int g = gidx[k];
                    float dx = pxf - means2d[g * 2], dy = pyf - means2d[g * 2 + 1];
                    float power = blend_power(dx, dy, conic[g * 3], conic[g * 3 + 1], conic[g * 3 + 2]);
                    ++ev;
                    if (power > 0.0f) { ++rej; continue; }
                    float alpha = fminf(opa[g] * expf(power), 0.99f);
                    if (alpha < 1.0f / 255.0f) { ++rej; continue; }
                    ++con;
                    float w = alpha * T;
                    C0 = fmaf(w, rgb[g * 3], C0);
                    C1 = fmaf(w, rgb[g * 3 + 1], C1);
                    C2 = fmaf(w, rgb[g * 3 + 2], C2);
                    T *= (1.0f - alpha);
                    ++cnt;
                    if (T < 1.0f / 255.0f) break; /* the crossing Gaussian IS composited */
                }
                int pi = py * W + px;
                color[pi * 3] = fmaf(T, bg[0], C0);
                color[pi * 3 + 1] = fmaf(T, bg[1], C1);
                color[pi * 3 + 2] = fmaf(T, bg[2], C2);
                final_T[pi] = T;
                n_contrib[pi] = cnt;
            }
    }
    if (stats) { stats[0] = ev; stats[1] = rej; stats[2] = con; }
}

/* ------------------------------------------------------------------------------------------
 * A.5 blend backward: rasterizer/backward.cu:31-233. Accumulates in double (the reference's
 * float atomicAdd order is nondeterministic, so this side is the tolerance oracle).
 * Walks the WHOLE tile range from the end and stops after n_contrib alpha-passing Gaussians
 * (backward.cu:141-145) — i.e. the LAST n_contrib of the range for saturated pixels.
 * ---------------------------------------------------------------------------------------- */
void oracle_blend_bwd(int W, int H, const float* bg, const int* ranges, const int* gidx,
                      const float* means2d, const float* conic, const float* rgb, const float* opa,
                      const float* dL_dcolor, const float* final_T, const int* n_contrib, int64_t n,
                      float* dL_drgb, float* dL_dopa, float* dL_dmean, float* dL_dconic,
                      int64_t* stats) {
    const int ntx = (W + TILE - 1) / TILE, nty = (H + TILE - 1) / TILE;
    double* acc = (double*)calloc((size_t)n * 9, sizeof(double));
    int64_t ev = 0, con = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : ev, con)
    for (int tile = 0; tile < ntx * nty; ++tile) {
        int ty = tile / ntx, tx = tile % ntx;
        int r0 = ranges[tile * 2], r1 = ranges[tile * 2 + 1];
        for (int ly = 0; ly < TILE; ++ly)
            for (int lx = 0; lx < TILE; ++lx) {
                int px = tx * TILE + lx, py = ty * TILE + ly;
                if (px >= W || py >= H) continue;
                int pi = py * W + px;
                float pxf = (float)px + 0.5f, pyf = (float)py + 0.5f;
                float T = final_T[pi];
                int maxc = n_contrib[pi];
                float g0 = dL_dcolor[pi * 3], g1 = dL_dcolor[pi * 3 + 1], g2 = dL_dcolor[pi * 3 + 2];
                float S0 = T * bg[0], S1 = T * bg[1], S2 = T * bg[2];
                int found = 0;
                for (int k = r1 - 1; k >= r0; --k) {
                    int g = gidx[k];
                    float a = conic[g * 3], b = conic[g * 3 + 1], c = conic[g * 3 + 2];
                    float dx = pxf - means2d[g * 2], dy = pyf - means2d[g * 2 + 1];
                    float power = blend_power(dx, dy, a, b, c);
                    ++ev;
                    if (power > 0.0f) continue;
                    float ex = expf(power);
                    float alpha = fminf(opa[g] * ex, 0.99f);
                    if (alpha < 1.0f / 255.0f) continue;
                    if (++found > maxc) break;
                    ++con;
                    float oma = fmaxf(1.0f - alpha, 1e-5f);
                    T /= oma;
                    float w = alpha * T;
                    float r = rgb[g * 3], gg = rgb[g * 3 + 1], bb = rgb[g * 3 + 2];
                    float dLa = 0.0f;
                    dLa += g0 * (T * r - S0 / oma);
                    dLa += g1 * (T * gg - S1 / oma);
                    dLa += g2 * (T * bb - S2 / oma);
                    S0 += w * r; S1 += w * gg; S2 += w * bb;
                    int clamped = (opa[g] * ex >= 0.99f);
                    float dLo = clamped ? 0.0f : dLa * ex;
                    float dLp = clamped ? 0.0f : dLa * alpha;
                    float v[9];
                    v[0] = g0 * w; v[1] = g1 * w; v[2] = g2 * w;
                    v[3] = dLo;
                    v[4] = dLp * (a * dx + b * dy);
                    v[5] = dLp * (b * dx + c * dy);
                    v[6] = dLp * (-0.5f * dx * dx);
                    v[7] = dLp * (-dx * dy);
                    v[8] = dLp * (-0.5f * dy * dy);
                    for (int q = 0; q < 9; ++q) {
#pragma omp atomic
                        acc[(size_t)g * 9 + q] += (double)v[q];
                    }
                }
            }
    }
    for (int64_t i = 0; i < n; ++i) {
        dL_drgb[i * 3] = (float)acc[i * 9]; dL_drgb[i * 3 + 1] = (float)acc[i * 9 + 1];
        dL_drgb[i * 3 + 2] = (float)acc[i * 9 + 2];
        dL_dopa[i] = (float)acc[i * 9 + 3];
        dL_dmean[i * 2] = (float)acc[i * 9 + 4]; dL_dmean[i * 2 + 1] = (float)acc[i * 9 + 5];
        dL_dconic[i * 3] = (float)acc[i * 9 + 6]; dL_dconic[i * 3 + 1] = (float)acc[i * 9 + 7];
        dL_dconic[i * 3 + 2] = (float)acc[i * 9 + 8];
    }
    free(acc);
    if (stats) { stats[0] = ev; stats[1] = con; }
}

/* ------------------------------------------------------------------------------------------
 * A.6 preprocess backward: rasterizer/projection_backward.cu:26-247 + backward.cuh:37-346,
 * followed by A.7 SH backward (sh_backward.cu). All outputs fully defined.
 * ---------------------------------------------------------------------------------------- */
void oracle_preprocess_bwd(int64_t n, const float* view, float fx, float fy, float scale_mod,
                           const float* cam_center, int deg, int C, const float* pos,
                           const float* rot, const float* scl, const float* opa, const float* sh,
                           const int* radii, const float* dL_dmean2d, const float* dL_dconic,
                           const float* dL_drgb, const float* dL_dopa_act, float* dL_dpos,
                           float* dL_drot, float* dL_dscl, float* dL_dopa, float* dL_dsh) {
    const float Wm[9] = {view[0], view[1], view[2], view[4], view[5], view[6], view[8], view[9], view[10]};
    const float lsm = logf(scale_mod + 1e-8f);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        for (int k = 0; k < 3; ++k) { dL_dpos[i * 3 + k] = 0.f; dL_dscl[i * 3 + k] = 0.f; }
        for (int k = 0; k < 4; ++k) dL_drot[i * 4 + k] = 0.f;
        dL_dopa[i] = 0.f;

        /* SH backward runs for all N (sh_backward.cu:38; culled have dL_drgb = 0). */
        {
            float d[3], Y[16];
            view_dir(pos + i * 3, cam_center, d);
            sh_basis(deg, d[0], d[1], d[2], Y);
            int na = (deg + 1) * (deg + 1);
            for (int ch = 0; ch < 3; ++ch) {
                const float* c = sh + (i * 3 + ch) * C;
                float* o = dL_dsh + (i * 3 + ch) * C;
                float raw = 0.0f;
                for (int k = 0; k < na; ++k) raw += c[k] * Y[k];
                raw += 0.5f;
                float g = dL_drgb[i * 3 + ch] * ((raw > 0.0f) ? 1.0f : 0.0f);
                for (int k = 0; k < na; ++k) o[k] = g * Y[k];
                for (int k = na; k < C; ++k) o[k] = 0.0f;
            }
        }
        if (radii[i] <= 0) continue; /* projection_backward.cu:48 */

        float px = pos[i * 3], py = pos[i * 3 + 1], pz = pos[i * 3 + 2];
        float t[3];
        t[0] = fmaf(pz, Wm[2], fmaf(px, Wm[0], py * Wm[1])) + view[3];
        t[1] = fmaf(pz, Wm[5], fmaf(px, Wm[3], py * Wm[4])) + view[7];
        t[2] = fmaf(pz, Wm[8], fmaf(px, Wm[6], py * Wm[7])) + view[11];
        float ls[3] = {scl[i * 3] + lsm, scl[i * 3 + 1] + lsm, scl[i * 3 + 2] + lsm};
        float cov3[6], M[9], R[9], s[3], c2[3], Tfw[6], inv_norm;
        compute_cov3d(ls, rot + i * 4, cov3, M, R, s, &inv_norm);
        compute_cov2d(cov3, Wm, t, fx, fy, c2, Tfw);
        float det = fmaf(c2[0], c2[2], -(c2[1] * c2[1]));
        if (det <= 0.0f) continue; /* :95 */
        float inv_det = 1.0f / det;
        float ia = c2[2] * inv_det, ib = -c2[1] * inv_det, ic = c2[0] * inv_det;

        float tz_inv = 1.0f / (t[2] + 1e-6f), tz_inv2 = tz_inv * tz_inv;
        float J0 = fx * tz_inv, J2 = -fx * t[0] * tz_inv2, J4 = fy * tz_inv, J5 = -fy * t[1] * tz_inv2;
        float Tm[6] = {J0 * Wm[0] + J2 * Wm[6], J0 * Wm[1] + J2 * Wm[7], J0 * Wm[2] + J2 * Wm[8],
                       J4 * Wm[3] + J5 * Wm[6], J4 * Wm[4] + J5 * Wm[7], J4 * Wm[5] + J5 * Wm[8]};

        /* backward.cuh:37-64: dSigma' = -Sinv dSinv Sinv, off-diagonal halved first. */
        float da = dL_dconic[i * 3], db = dL_dconic[i * 3 + 1] * 0.5f, dc = dL_dconic[i * 3 + 2];
        float t00 = ia * da + ib * db, t01 = ia * db + ib * dc;
        float t10 = ib * da + ic * db, t11 = ib * db + ic * dc;
        float d2a = -(t00 * ia + t01 * ib), d2b = -(t00 * ib + t01 * ic), d2c = -(t10 * ib + t11 * ic);

        /* backward.cuh:82-107: dSigma = T^T dSigma' T (upper triangle). */
        float TtD[6] = {Tm[0] * d2a + Tm[3] * d2b, Tm[0] * d2b + Tm[3] * d2c,
                        Tm[1] * d2a + Tm[4] * d2b, Tm[1] * d2b + Tm[4] * d2c,
                        Tm[2] * d2a + Tm[5] * d2b, Tm[2] * d2b + Tm[5] * d2c};
        float d3[6] = {TtD[0] * Tm[0] + TtD[1] * Tm[3], TtD[0] * Tm[1] + TtD[1] * Tm[4],
                       TtD[0] * Tm[2] + TtD[1] * Tm[5], TtD[2] * Tm[1] + TtD[3] * Tm[4],
                       TtD[2] * Tm[2] + TtD[3] * Tm[5], TtD[4] * Tm[2] + TtD[5] * Tm[5]};

        /* backward.cuh:123-153: dM = 2 dSigma_full M. */
        float F[9] = {d3[0], d3[1], d3[2], d3[1], d3[3], d3[4], d3[2], d3[4], d3[5]};
        float dM[9];
        for (int r = 0; r < 3; ++r)
            for (int j = 0; j < 3; ++j)
                dM[r * 3 + j] = 2.0f * (F[r * 3] * M[j] + F[r * 3 + 1] * M[3 + j] + F[r * 3 + 2] * M[6 + j]);

        /* projection_backward.cu:170-184 */
        float dR[9];
        for (int r = 0; r < 3; ++r)
            for (int j = 0; j < 3; ++j) dR[r * 3 + j] = dM[r * 3 + j] * s[j];
        for (int j = 0; j < 3; ++j) {
            float ds = dM[j] * R[j] + dM[3 + j] * R[3 + j] + dM[6 + j] * R[6 + j];
            dL_dscl[i * 3 + j] = ds * s[j];
        }

        /* backward.cuh:168-227 compute_dL_dquat (normalised q, then normalisation Jacobian). */
        {
            float w = rot[i * 4] * inv_norm, x = rot[i * 4 + 1] * inv_norm;
            float y = rot[i * 4 + 2] * inv_norm, z = rot[i * 4 + 3] * inv_norm;
            float dw = 2.0f * (-z * dR[1] + y * dR[2] + z * dR[3] - x * dR[5] + -y * dR[6] + x * dR[7]);
            float dxq = 2.0f * (y * dR[1] + z * dR[2] + y * dR[3] - 2.0f * x * dR[4] - w * dR[5] +
                                z * dR[6] + w * dR[7] - 2.0f * x * dR[8]);
            float dyq = 2.0f * (-2.0f * y * dR[0] + x * dR[1] + w * dR[2] + x * dR[3] + z * dR[5] +
                                -w * dR[6] + z * dR[7] - 2.0f * y * dR[8]);
            float dzq = 2.0f * (-2.0f * z * dR[0] - w * dR[1] + x * dR[2] + w * dR[3] -
                                2.0f * z * dR[4] + y * dR[5] + x * dR[6] + y * dR[7]);
            float dot = dw * w + dxq * x + dyq * y + dzq * z;
            dL_drot[i * 4 + 0] = inv_norm * (dw - w * dot);
            dL_drot[i * 4 + 1] = inv_norm * (dxq - x * dot);
            dL_drot[i * 4 + 2] = inv_norm * (dyq - y * dot);
            dL_drot[i * 4 + 3] = inv_norm * (dzq - z * dot);
        }

        /* projection_backward.cu:196-205: through means_2d. */
        float m0 = dL_dmean2d[i * 2], m1 = dL_dmean2d[i * 2 + 1];
        float dt[3] = {0.f, 0.f, 0.f};
        dt[0] += m0 * fx * tz_inv;
        dt[1] += m1 * fy * tz_inv;
        dt[2] += m0 * (-fx * t[0] * tz_inv2) + m1 * (-fy * t[1] * tz_inv2);

        /* backward.cuh:248-346: through J's dependence on t_cam. */
        {
            float TS[6] = {Tm[0] * cov3[0] + Tm[1] * cov3[1] + Tm[2] * cov3[2],
                           Tm[0] * cov3[1] + Tm[1] * cov3[3] + Tm[2] * cov3[4],
                           Tm[0] * cov3[2] + Tm[1] * cov3[4] + Tm[2] * cov3[5],
                           Tm[3] * cov3[0] + Tm[4] * cov3[1] + Tm[5] * cov3[2],
                           Tm[3] * cov3[1] + Tm[4] * cov3[3] + Tm[5] * cov3[4],
                           Tm[3] * cov3[2] + Tm[4] * cov3[4] + Tm[5] * cov3[5]};
            float dT[6] = {2.0f * (d2a * TS[0] + d2b * TS[3]), 2.0f * (d2a * TS[1] + d2b * TS[4]),
                           2.0f * (d2a * TS[2] + d2b * TS[5]), 2.0f * (d2b * TS[0] + d2c * TS[3]),
                           2.0f * (d2b * TS[1] + d2c * TS[4]), 2.0f * (d2b * TS[2] + d2c * TS[5])};
            float dJ0 = dT[0] * Wm[0] + dT[1] * Wm[1] + dT[2] * Wm[2];
            float dJ2 = dT[0] * Wm[6] + dT[1] * Wm[7] + dT[2] * Wm[8];
            float dJ4 = dT[3] * Wm[3] + dT[4] * Wm[4] + dT[5] * Wm[5];
            float dJ5 = dT[3] * Wm[6] + dT[4] * Wm[7] + dT[5] * Wm[8];
            float tz_inv3 = tz_inv2 * tz_inv;
            dt[0] += dJ2 * (-fx * tz_inv2);
            dt[1] += dJ5 * (-fy * tz_inv2);
            dt[2] += dJ0 * (-fx * tz_inv2) + dJ2 * (2.0f * fx * t[0] * tz_inv3) +
                     dJ4 * (-fy * tz_inv2) + dJ5 * (2.0f * fy * t[1] * tz_inv3);
        }
        /* :217-219 dpos = W^T dt */
        dL_dpos[i * 3 + 0] = Wm[0] * dt[0] + Wm[3] * dt[1] + Wm[6] * dt[2];
        dL_dpos[i * 3 + 1] = Wm[1] * dt[0] + Wm[4] * dt[1] + Wm[7] * dt[2];
        dL_dpos[i * 3 + 2] = Wm[2] * dt[0] + Wm[5] * dt[1] + Wm[8] * dt[2];
        /* :226-228 */
        float sg = 1.0f / (1.0f + expf(-opa[i]));
        dL_dopa[i] = dL_dopa_act[i] * sg * (1.0f - sg);
    }
}

/* ------------------------------------------------------------------------------------------
 * A.8 loss: training/loss.cpp:83-135 (L1, SSIM 11x11 sigma 1.5 zero-padded, combined) and its
 * gradient (the reference obtains it by libtorch autograd, trainer.cpp:214-217; restated
 * analytically). Computed in double: this is the tolerance oracle.
 * scalars: [loss, l1, ssim_mean]. dL_dcolor may be NULL.
 * ---------------------------------------------------------------------------------------- */
static void ssim_window(double w[11]) {
    /* loss.cpp:57-70: float 1-D gaussian, normalised; 2-D = outer product, re-normalised. */
    float k[11], sum = 0.f;
    for (int i = 0; i < 11; ++i) { float x = (float)(i - 5); k[i] = expf(-x * x / (2.0f * 1.5f * 1.5f)); sum += k[i]; }
    for (int i = 0; i < 11; ++i) k[i] = k[i] / sum;
    double s2 = 0.0;
    for (int i = 0; i < 11; ++i) for (int j = 0; j < 11; ++j) s2 += (double)(k[i] * k[j]);
    for (int i = 0; i < 11; ++i) w[i] = (double)k[i] / sqrt(s2);
}

static void conv_sep(int W, int H, const double* in, const double* w, double* tmp, double* out) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            double a = 0.0;
            for (int k = -5; k <= 5; ++k) { int xx = x + k; if (xx >= 0 && xx < W) a += w[k + 5] * in[y * W + xx]; }
            tmp[y * W + x] = a;
        }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            double a = 0.0;
            for (int k = -5; k <= 5; ++k) { int yy = y + k; if (yy >= 0 && yy < H) a += w[k + 5] * tmp[yy * W + x]; }
            out[y * W + x] = a;
        }
}

void oracle_loss(int W, int H, float lambda, const float* rendered, const float* target,
                 float* dL_dcolor, float* scalars) {
    const size_t np = (size_t)W * H;
    const double C1 = (double)(0.01f * 0.01f), C2 = (double)(0.03f * 0.03f);
    double w[11];
    ssim_window(w);
    double* buf = (double*)malloc(np * 12 * sizeof(double));
    double *X = buf, *Y = buf + np, *in = buf + 2 * np, *tmp = buf + 3 * np, *mux = buf + 4 * np,
           *muy = buf + 5 * np, *exx = buf + 6 * np, *eyy = buf + 7 * np, *exy = buf + 8 * np,
           *g1 = buf + 9 * np, *g2 = buf + 10 * np, *g3 = buf + 11 * np;
    double l1 = 0.0, ssum = 0.0;
    const double inv_n = 1.0 / (3.0 * (double)np);
    for (int ch = 0; ch < 3; ++ch) {
        for (size_t p = 0; p < np; ++p) { X[p] = rendered[p * 3 + ch]; Y[p] = target[p * 3 + ch]; l1 += fabs(X[p] - Y[p]); }
        conv_sep(W, H, X, w, tmp, mux);
        conv_sep(W, H, Y, w, tmp, muy);
        for (size_t p = 0; p < np; ++p) in[p] = X[p] * X[p];
        conv_sep(W, H, in, w, tmp, exx);
        for (size_t p = 0; p < np; ++p) in[p] = Y[p] * Y[p];
        conv_sep(W, H, in, w, tmp, eyy);
        for (size_t p = 0; p < np; ++p) in[p] = X[p] * Y[p];
        conv_sep(W, H, in, w, tmp, exy);
        for (size_t p = 0; p < np; ++p) {
            double mx = mux[p], my = muy[p];
            double sxx = exx[p] - mx * mx, syy = eyy[p] - my * my, sxy = exy[p] - mx * my;
            double A1 = 2.0 * mx * my + C1, A2 = 2.0 * sxy + C2;
            double B1 = mx * mx + my * my + C1, B2 = sxx + syy + C2;
            double S = (A1 * A2) / (B1 * B2);
            ssum += S;
            /* partials wrt (mu_x, E[x^2], E[xy]) holding the raw moments fixed */
            double dmu = (2.0 * my * A2 + A1 * (-2.0 * my)) / (B1 * B2) - S * (2.0 * mx / B1 + (-2.0 * mx) / B2);
            g1[p] = dmu;
            g2[p] = -S / B2;
            g3[p] = 2.0 * A1 / (B1 * B2);
        }
        if (dL_dcolor) {
            conv_sep(W, H, g1, w, tmp, mux);
            conv_sep(W, H, g2, w, tmp, exx);
            conv_sep(W, H, g3, w, tmp, exy);
            for (size_t p = 0; p < np; ++p) {
                double d = X[p] - Y[p];
                double sgn = (d > 0.0) ? 1.0 : ((d < 0.0) ? -1.0 : 0.0);
                double dssim = mux[p] + 2.0 * X[p] * exx[p] + Y[p] * exy[p];
                dL_dcolor[p * 3 + ch] = (float)((1.0 - (double)lambda) * sgn * inv_n - (double)lambda * inv_n * dssim);
            }
        }
    }
    l1 *= inv_n; ssum *= inv_n;
    scalars[0] = (float)((1.0 - (double)lambda) * l1 + (double)lambda * (1.0 - ssum));
    scalars[1] = (float)l1;
    scalars[2] = (float)ssum;
    free(buf);
}

/* ------------------------------------------------------------------------------------------
 * A.9 Adam: optimizer/fused_adam.cu:44-76 (one group). bc1/bc2 computed by the caller in
 * double as fused_adam.cu:140-149.
 * ---------------------------------------------------------------------------------------- */
void oracle_adam(int64_t n, float* p, const float* g, float* m, float* v, float lr, float b1,
                 float b2, float eps, float bc1, float bc2) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        float gi = g[i];
        /* contraction nvcc 12.9 applies to fused_adam.cu:62-67 (read off the sm_100 SASS) */
        float mi = fmaf(gi, 1.0f - b1, b1 * m[i]);
        m[i] = mi;
        float vi = fmaf(gi, (1.0f - b2) * gi, b2 * v[i]);
        v[i] = vi;
        float mh = mi * bc1, vh = vi * bc2;
        p[i] -= lr * mh / (sqrtf(vh) + eps);
    }
}

/* optimizer/densification.cpp:59-88 accumulate_gradients. */
void oracle_accumulate_stats(int64_t n, const float* dL_dmean2d, const int* radii, float* grad_accum,
                             float* grad_count, float* max_radii) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        if (radii[i] > 0) {
            float gx = dL_dmean2d[i * 2], gy = dL_dmean2d[i * 2 + 1];
            grad_accum[i] += sqrtf(gx * gx + gy * gy);
            grad_count[i] += 1.0f;
        }
        float r = (float)radii[i];
        if (r > max_radii[i]) max_radii[i] = r;
    }
}
