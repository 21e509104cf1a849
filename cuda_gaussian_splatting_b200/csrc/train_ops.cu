// train_ops.cu — the per-step kernels either side of the rasterizer:
//   * fused L1 + SSIM loss, value AND gradient w.r.t. the rendered image, replacing the libtorch
//     op graph of reference training/loss.cpp:83-135 (5 grouped 11x11 conv2d + ~20 elementwise
//     kernels forward, the same again in autograd backward, training/trainer.cpp:214-217);
//   * multi-tensor fused Adam: all five parameter groups in ONE launch, float4 accesses,
//     replacing 5 x k_fused_adam (reference optimizer/fused_adam.cu:44-76, :140-164);
//   * densification statistics (reference optimizer/densification.cpp:59-88, ~10 libtorch kernels
//     with boolean-mask gathers) as one elementwise kernel.
// All three are HBM-bound: loss 36 B/pixel algorithmic (+72 B/pixel for the three derivative maps
// kept between the two passes), Adam 28 B/element, stats 36 B/Gaussian.
#include "common.cuh"

#include <cmath>

namespace cugs {

// ================================================================================================
// L1 + SSIM
// ================================================================================================
constexpr int kLossTile = 16;
constexpr int kHalo = 5;
constexpr int kLossIn = kLossTile + 2 * kHalo;  // 26
constexpr int kInStride = 28;                   // padded planar row: every 4-column run starts 16-byte aligned
constexpr int kRun = 4;                         // outputs per work item along the filter direction
constexpr int kRunIn = kRun + 10;               // inputs a run needs (11 taps)
constexpr int kHItems = kLossIn * 3 * (kLossTile / kRun);      // horizontal work items: (row, run, channel), row fastest
constexpr int kVItems = kLossTile * 3 * (kLossTile / kRun);    // vertical work items: (column*3+channel, run)
constexpr int kCC = kLossTile * 3;                             // interleaved output columns of a tile (48)
// Shared-memory layouts are chosen for conflict-free VECTOR accesses in both passes:
//   inputs   [channel][row][28]: horizontal items of one quarter-warp are 8 consecutive rows; a row stride
//            of 28 floats (= -4 mod 32 banks) makes their LDS.128 tile all 32 banks;
//   filtered [map][column*3+channel][28] (row fastest): the horizontal pass stores with consecutive rows in
//            consecutive lanes, the vertical pass reads its 14 rows with LDS.128, 8 consecutive columns
//            again tiling the banks.

struct SsimWindow {
    float w[11];  // separable factor of the reference's 2-D window (loss.cpp:57-70)
};

// Both passes are separable 11-tap filters over a 16x16 tile + halo staged in shared memory. Shared
// memory bandwidth, not FP32, bounds a tap-by-tap version (11 LDS per output and map), so every work
// item filters a RUN of 4 outputs from 14 inputs held in registers (3.5 LDS per output, vector loads
// in the horizontal direction: the input tiles are stored planar per channel with padded rows).
__device__ __forceinline__ void load_run(const float* __restrict__ row, float (&v)[kRunIn]) {
    const float4 q0 = *reinterpret_cast<const float4*>(row);
    const float4 q1 = *reinterpret_cast<const float4*>(row + 4);
    const float4 q2 = *reinterpret_cast<const float4*>(row + 8);
    const float2 q3 = *reinterpret_cast<const float2*>(row + 12);
    v[0] = q0.x; v[1] = q0.y; v[2] = q0.z; v[3] = q0.w;
    v[4] = q1.x; v[5] = q1.y; v[6] = q1.z; v[7] = q1.w;
    v[8] = q2.x; v[9] = q2.y; v[10] = q2.z; v[11] = q2.w;
    v[12] = q3.x; v[13] = q3.y;
}

__device__ __forceinline__ void h_item(int item, int& r, int& run, int& ch) {
    r = item % kLossIn;
    const int t = item / kLossIn;
    run = t & 3;
    ch = t >> 2;
}

// out[o] = sum_k w[k] * v[o + k], k ascending (the summation order of the tap-by-tap loop)
__device__ __forceinline__ void filter_run(const SsimWindow& win, const float (&v)[kRunIn], float (&out)[kRun]) {
#pragma unroll
    for (int o = 0; o < kRun; ++o) out[o] = 0.f;
#pragma unroll
    for (int j = 0; j < kRunIn; ++j) {
#pragma unroll
        for (int o = 0; o < kRun; ++o) {
            const int k = j - o;
            if (k >= 0 && k < 11) out[o] = fmaf(win.w[k], v[j], out[o]);
        }
    }
}

// tile + halo of kImgs interleaved [H,W,3] images -> planar shared memory, zero outside the image
// (conv2d padding = 5, loss.cpp:102-103). A row of the tile is 78 contiguous floats in global memory;
// thread t owns one of them (t % 78) in every third row (t / 78), so all index arithmetic is hoisted
// out of the loop and the 9 x kImgs loads of a thread are independent (in flight together).
template <int kImgs>
__device__ __forceinline__ void load_tiles_planar(const float* const (&img)[kImgs], int width, int height, int tx0,
                                                  int ty0, float (*const (&dst)[kImgs])[kLossIn][kInStride]) {
    constexpr int kRowF = kLossIn * 3;  // 78
    const int t = threadIdx.x;
    const int rsub = t / kRowF, cc = t - rsub * kRowF;
    if (rsub >= 3) return;  // 234 of the 256 threads load
    const int col = cc / 3, ch = cc - col * 3;
    const int gx = tx0 - kHalo + col;
    const bool x_ok = gx >= 0 && gx < width;
    const int64_t g0 = ((int64_t)(ty0 - kHalo + rsub) * width + (tx0 - kHalo)) * 3 + cc;
    const int64_t gstep = (int64_t)3 * width * 3;
    float v[kImgs][9];
#pragma unroll
    for (int it = 0; it < 9; ++it) {
        const int r = rsub + 3 * it;
        const int gy = ty0 - kHalo + r;
        const bool ok = x_ok && r < kLossIn && gy >= 0 && gy < height;
#pragma unroll
        for (int m = 0; m < kImgs; ++m) v[m][it] = ok ? __ldg(img[m] + g0 + it * gstep) : 0.f;
    }
#pragma unroll
    for (int it = 0; it < 9; ++it) {
        const int r = rsub + 3 * it;
        if (r < kLossIn) {
#pragma unroll
            for (int m = 0; m < kImgs; ++m) dst[m][ch][r][col] = v[m][it];
        }
    }
}

// pass 1: local moments -> SSIM map value (summed) and the three partial-derivative maps
//   g1 = dS/dmu_x, g2 = dS/dE[x^2], g3 = dS/dE[xy]   (raw moments held fixed)
__global__ void __launch_bounds__(256, 4)
k_ssim_moments(int width, int height, SsimWindow win, const float* __restrict__ x_img,
               const float* __restrict__ y_img, float* __restrict__ g1, float* __restrict__ g2,
               float* __restrict__ g3, double* __restrict__ sums /* [2]: sum|x-y|, sum S */,
               float* __restrict__ ssim_map /* optional [H,W]: channel mean of S (loss.cpp:123) */) {
    __shared__ __align__(16) float sx[3][kLossIn][kInStride];
    __shared__ __align__(16) float sy[3][kLossIn][kInStride];
    // four moment maps: mu_x, mu_y, E[x^2 + y^2], E[xy] (only the SUM of the two variances is needed),
    // horizontally filtered, channel-interleaved so that the vertical pass reads them conflict-free
    __shared__ __align__(16) float sh[4][kCC][kInStride];
    __shared__ float s_red[2][8];

    const int tx0 = blockIdx.x * kLossTile, ty0 = blockIdx.y * kLossTile;
    {
        const float* const imgs[2] = {x_img, y_img};
        float (*const dsts[2])[kLossIn][kInStride] = {sx, sy};
        load_tiles_planar<2>(imgs, width, height, tx0, ty0, dsts);
    }
    __syncthreads();
    // horizontal pass: (26 rows x 4 runs x 3 channels) items, four moments each
    for (int item = threadIdx.x; item < kHItems; item += 256) {
        int r, run, ch;
        h_item(item, r, run, ch);
        float a[kRunIn], b[kRunIn], p[kRunIn], q[kRunIn], o4[kRun];
        load_run(&sx[ch][r][run * kRun], a);
        load_run(&sy[ch][r][run * kRun], b);
#pragma unroll
        for (int j = 0; j < kRunIn; ++j) {
            p[j] = fmaf(a[j], a[j], b[j] * b[j]);
            q[j] = a[j] * b[j];
        }
        float* o0 = &sh[0][run * kRun * 3 + ch][r];
        filter_run(win, a, o4);
#pragma unroll
        for (int o = 0; o < kRun; ++o) o0[o * 3 * kInStride] = o4[o];
        filter_run(win, b, o4);
#pragma unroll
        for (int o = 0; o < kRun; ++o) o0[(kCC + o * 3) * kInStride] = o4[o];
        filter_run(win, p, o4);
#pragma unroll
        for (int o = 0; o < kRun; ++o) o0[(2 * kCC + o * 3) * kInStride] = o4[o];
        filter_run(win, q, o4);
#pragma unroll
        for (int o = 0; o < kRun; ++o) o0[(3 * kCC + o * 3) * kInStride] = o4[o];
    }
    __syncthreads();
    // vertical pass + SSIM: (48 interleaved columns x 4 runs of 4 rows) items
    float l1_local = 0.f, s_local = 0.f;
    float* s_S = &sy[0][0][0];  // per (pixel, channel) S for the optional map; sy is dead after the l1 term below
    float S_keep[kRun];
    bool have_item = false;
    int it_cc = 0, it_run = 0;
    if (threadIdx.x < kVItems) {
        have_item = true;
        const int cc = threadIdx.x % (kLossTile * 3), run = threadIdx.x / (kLossTile * 3);
        it_cc = cc; it_run = run;
        const int col = cc / 3, ch = cc - col * 3;
        float m[4][kRun];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float v[kRunIn];
            load_run(&sh[k][cc][run * kRun], v);
            filter_run(win, v, m[k]);
        }
        const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
        const int gx = tx0 + col;
#pragma unroll
        for (int o = 0; o < kRun; ++o) {
            const int ly = run * kRun + o, gy = ty0 + ly;
            S_keep[o] = 0.f;
            if (gx < width && gy < height) {
                const float mx = m[0][o], my = m[1][o], qq = m[2][o], xy = m[3][o];
                const float sxy = xy - mx * my;
                const float A1 = 2.0f * mx * my + C1, A2 = 2.0f * sxy + C2;
                const float B1 = mx * mx + my * my + C1, B2 = (qq - mx * mx - my * my) + C2;  // sigma_x^2 + sigma_y^2 + C2
                const float inv = __frcp_rn(B1 * B2);  // one reciprocal: 1/B1 = B2 inv, 1/B2 = B1 inv
                const float S = A1 * A2 * inv;
                const float rB1 = B2 * inv, rB2 = B1 * inv;
                const int64_t gi = ((int64_t)gy * width + gx) * 3 + ch;
                // dS/dmu_x = (A1' A2 + A1 A2')/(B1 B2) - S (B1'/B1 + B2'/B2)
                g1[gi] = 2.0f * my * (A2 - A1) * inv - 2.0f * mx * S * (rB1 - rB2);
                g2[gi] = -S * rB2;
                g3[gi] = 2.0f * A1 * inv;
                s_local += S;
                S_keep[o] = S;
                l1_local += fabsf(sx[ch][ly + kHalo][col + kHalo] - sy[ch][ly + kHalo][col + kHalo]);
            }
        }
    }
    if (ssim_map != nullptr) {  // uniform branch
        __syncthreads();
        if (have_item) {
#pragma unroll
            for (int o = 0; o < kRun; ++o) s_S[(it_run * kRun + o) * (kLossTile * 3) + it_cc] = S_keep[o];
        }
        __syncthreads();
        const int lx = threadIdx.x % kLossTile, ly = threadIdx.x / kLossTile;
        if (tx0 + lx < width && ty0 + ly < height) {
            const float* t = &s_S[ly * (kLossTile * 3) + lx * 3];
            ssim_map[(int64_t)(ty0 + ly) * width + tx0 + lx] = ((t[0] + t[1]) + t[2]) / 3.0f;
        }
    }
    // block reduction -> one double atomic per block and quantity
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        l1_local += __shfl_xor_sync(kFull, l1_local, d);
        s_local += __shfl_xor_sync(kFull, s_local, d);
    }
    if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = l1_local; s_red[1][threadIdx.x >> 5] = s_local; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < 8; ++w) { a += (double)s_red[0][w]; b += (double)s_red[1][w]; }
        atomicAdd(&sums[0], a);
        atomicAdd(&sums[1], b);
    }
}

// pass 2: dL/dx = (1-l) sign(x-y)/n - l/n * (W*g1 + 2 x (W*g2) + y (W*g3))
__global__ void __launch_bounds__(256, 4)
k_ssim_gradient(int width, int height, SsimWindow win, float lambda, const float* __restrict__ x_img,
                const float* __restrict__ y_img, const float* __restrict__ g1,
                const float* __restrict__ g2, const float* __restrict__ g3, float* __restrict__ dL_dx) {
    __shared__ __align__(16) float sg[3][3][kLossIn][kInStride];  // [map][channel] planar
    __shared__ __align__(16) float sh[3][kCC][kInStride];
    const int tx0 = blockIdx.x * kLossTile, ty0 = blockIdx.y * kLossTile;
    {
        const float* const imgs[3] = {g1, g2, g3};
        float (*const dsts[3])[kLossIn][kInStride] = {sg[0], sg[1], sg[2]};
        load_tiles_planar<3>(imgs, width, height, tx0, ty0, dsts);
    }
    // the centre pixels of this thread's vertical run, fetched now so that their latency hides behind
    // the two filter passes
    const int cc = threadIdx.x % (kLossTile * 3), run = threadIdx.x / (kLossTile * 3);
    const int col = cc / 3, ch = cc - col * 3;
    const int gx = tx0 + col;
    float xc[kRun], yc[kRun];
#pragma unroll
    for (int o = 0; o < kRun; ++o) {
        const int gy = ty0 + run * kRun + o;
        const bool ok = threadIdx.x < kVItems && gx < width && gy < height;
        const int64_t gi = ((int64_t)gy * width + gx) * 3 + ch;
        xc[o] = ok ? __ldg(x_img + gi) : 0.f;
        yc[o] = ok ? __ldg(y_img + gi) : 0.f;
    }
    __syncthreads();
    for (int item = threadIdx.x; item < kHItems; item += 256) {
        int r, hrun, hch;
        h_item(item, r, hrun, hch);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float a[kRunIn], o4[kRun];
            load_run(&sg[k][hch][r][hrun * kRun], a);
            filter_run(win, a, o4);
            float* o0 = &sh[k][hrun * kRun * 3 + hch][r];
#pragma unroll
            for (int o = 0; o < kRun; ++o) o0[o * 3 * kInStride] = o4[o];
        }
    }
    __syncthreads();
    if (threadIdx.x >= kVItems) return;
    float m[3][kRun];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float v[kRunIn];
        load_run(&sh[k][cc][run * kRun], v);
        filter_run(win, v, m[k]);
    }
    if (gx >= width) return;
    const float inv_n = 1.0f / (3.0f * (float)width * (float)height);
#pragma unroll
    for (int o = 0; o < kRun; ++o) {
        const int gy = ty0 + run * kRun + o;
        if (gy >= height) break;
        const int64_t gi = ((int64_t)gy * width + gx) * 3 + ch;
        const float xv = xc[o], yv = yc[o];
        const float d = xv - yv;
        const float sgn = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
        const float dssim = m[0][o] + 2.0f * xv * m[1][o] + yv * m[2][o];
        dL_dx[gi] = (1.0f - lambda) * sgn * inv_n - lambda * inv_n * dssim;
    }
}

__global__ void k_loss_finalize(int width, int height, float lambda, const double* __restrict__ sums,
                                float* __restrict__ scalars3) {
    const double n = 3.0 * (double)width * (double)height;
    const double l1 = sums[0] / n, ss = sums[1] / n;
    scalars3[0] = (float)((1.0 - (double)lambda) * l1 + (double)lambda * (1.0 - ss));
    scalars3[1] = (float)l1;
    scalars3[2] = (float)ss;
}

// ------------------------------------------------------------------------------------------------
// any odd window (loss.hpp:33-44 lets the caller choose window_size; loss.cpp:88-124): the tuned kernels
// above are specialised for the default 11 taps, these two cover every other odd size 3..kMaxWindow with
// the same two-pass formulation (moments -> S and its three partial-derivative maps; gradient = the
// window applied to those maps). One thread per (pixel, channel), separable factor in constant-bank
// kernel arguments, direct w x w accumulation with the row factor hoisted: a fall-back for evaluation
// callers (metrics with a different window), not a tuned path.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxWindow = 33;
struct SsimWindowAny {
    int size;
    float w[kMaxWindow];
};

__global__ void __launch_bounds__(256)
k_ssim_moments_any(int width, int height, SsimWindowAny win, const float* __restrict__ x_img,
                   const float* __restrict__ y_img, float* __restrict__ g1, float* __restrict__ g2,
                   float* __restrict__ g3, double* __restrict__ sums, float* __restrict__ ssim_map) {
    __shared__ float s_red[2][8];
    const int64_t total = (int64_t)width * height * 3;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float l1_local = 0.f, s_local = 0.f;
    if (e < total) {
        const int ch = (int)(e % 3);
        const int64_t pix = e / 3;
        const int gx = (int)(pix % width), gy = (int)(pix / width);
        const int half = win.size / 2;
        float mx = 0.f, my = 0.f, qq = 0.f, xy = 0.f;
        for (int dy = -half; dy <= half; ++dy) {
            const int yy = gy + dy;
            if (yy < 0 || yy >= height) continue;  // zero padding (conv2d padding = window / 2)
            float rx = 0.f, ry = 0.f, rq = 0.f, rxy = 0.f;
            for (int dx = -half; dx <= half; ++dx) {
                const int xx = gx + dx;
                if (xx < 0 || xx >= width) continue;
                const int64_t gi = ((int64_t)yy * width + xx) * 3 + ch;
                const float a = __ldg(x_img + gi), b = __ldg(y_img + gi), w = win.w[dx + half];
                rx = fmaf(w, a, rx);
                ry = fmaf(w, b, ry);
                rq = fmaf(w, fmaf(a, a, b * b), rq);
                rxy = fmaf(w, a * b, rxy);
            }
            const float wy = win.w[dy + half];
            mx = fmaf(wy, rx, mx); my = fmaf(wy, ry, my); qq = fmaf(wy, rq, qq); xy = fmaf(wy, rxy, xy);
        }
        const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
        const float sxy = xy - mx * my;
        const float A1 = 2.0f * mx * my + C1, A2 = 2.0f * sxy + C2;
        const float B1 = mx * mx + my * my + C1, B2 = (qq - mx * mx - my * my) + C2;
        const float inv = __frcp_rn(B1 * B2);
        const float S = A1 * A2 * inv;
        const float rB1 = B2 * inv, rB2 = B1 * inv;
        g1[e] = 2.0f * my * (A2 - A1) * inv - 2.0f * mx * S * (rB1 - rB2);
        g2[e] = -S * rB2;
        g3[e] = 2.0f * A1 * inv;
        s_local = S;
        l1_local = fabsf(__ldg(x_img + e) - __ldg(y_img + e));
        if (ssim_map != nullptr) {  // channel mean: the three channels of a pixel sit in adjacent lanes / threads
            // (written by the channel-0 thread after a neighbour exchange through global memory would race:
            //  use an atomic on a map the host zeroed)
            atomicAdd(&ssim_map[pix], S / 3.0f);
        }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        l1_local += __shfl_xor_sync(kFull, l1_local, d);
        s_local += __shfl_xor_sync(kFull, s_local, d);
    }
    if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = l1_local; s_red[1][threadIdx.x >> 5] = s_local; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < 8; ++w) { a += (double)s_red[0][w]; b += (double)s_red[1][w]; }
        atomicAdd(&sums[0], a);
        atomicAdd(&sums[1], b);
    }
}

__global__ void __launch_bounds__(256)
k_ssim_gradient_any(int width, int height, SsimWindowAny win, float lambda, const float* __restrict__ x_img,
                    const float* __restrict__ y_img, const float* __restrict__ g1, const float* __restrict__ g2,
                    const float* __restrict__ g3, float* __restrict__ dL_dx) {
    const int64_t total = (int64_t)width * height * 3;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int ch = (int)(e % 3);
    const int64_t pix = e / 3;
    const int gx = (int)(pix % width), gy = (int)(pix / width);
    const int half = win.size / 2;
    float m1 = 0.f, m2 = 0.f, m3 = 0.f;
    for (int dy = -half; dy <= half; ++dy) {
        const int yy = gy + dy;
        if (yy < 0 || yy >= height) continue;
        float r1 = 0.f, r2 = 0.f, r3 = 0.f;
        for (int dx = -half; dx <= half; ++dx) {
            const int xx = gx + dx;
            if (xx < 0 || xx >= width) continue;
            const int64_t gi = ((int64_t)yy * width + xx) * 3 + ch;
            const float w = win.w[dx + half];
            r1 = fmaf(w, __ldg(g1 + gi), r1);
            r2 = fmaf(w, __ldg(g2 + gi), r2);
            r3 = fmaf(w, __ldg(g3 + gi), r3);
        }
        const float wy = win.w[dy + half];
        m1 = fmaf(wy, r1, m1); m2 = fmaf(wy, r2, m2); m3 = fmaf(wy, r3, m3);
    }
    const float inv_n = 1.0f / (3.0f * (float)width * (float)height);
    const float xv = __ldg(x_img + e), yv = __ldg(y_img + e);
    const float d = xv - yv;
    const float sgn = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
    dL_dx[e] = (1.0f - lambda) * sgn * inv_n - lambda * inv_n * (m1 + 2.0f * xv * m2 + yv * m3);
}

// loss.cpp:57-70 for any size: float 1-D gaussian (sigma 1.5) / sum; 2-D = outer product / its sum
static void window_factor(int size, float* out) {
    float k[kMaxWindow], sum = 0.f;
    const int half = size / 2;
    for (int i = 0; i < size; ++i) {
        const float x = (float)(i - half);
        k[i] = std::exp(-x * x / (2.0f * 1.5f * 1.5f));
        sum += k[i];
    }
    for (int i = 0; i < size; ++i) k[i] = k[i] / sum;
    double s2 = 0.0;
    for (int i = 0; i < size; ++i)
        for (int j = 0; j < size; ++j) s2 += (double)(k[i] * k[j]);
    for (int i = 0; i < size; ++i) out[i] = (float)((double)k[i] / std::sqrt(s2));
}

static SsimWindow make_window() {
    SsimWindow w;
    window_factor(11, w.w);
    return w;
}

// ================================================================================================
// multi-tensor Adam
// ================================================================================================
struct AdamGroups {
    float* p[5];
    const float* g[5];
    float* m[5];
    float* v[5];
    int64_t count[5];
    int64_t chunk_end[5];  // cumulative number of float4 chunks
    float lr[5];
};

__device__ __forceinline__ void adam_element(float& p, float g, float& m, float& v, float lr, float b1,
                                             float b2, float eps, float bc1, float bc2) {
    // fused_adam.cu:57-75 with the contraction nvcc 12.9 applies to the reference spelled out
    // (SASS of k_fused_adam for sm_100: FMUL b1*m, FFMA g*(1-b1)+.; FMUL (1-b2)*g, FMUL b2*v, FFMA)
    const float mi = fma_rn(g, add_rn(1.0f, -b1), mul_rn(b1, m));
    const float vi = fma_rn(g, mul_rn(add_rn(1.0f, -b2), g), mul_rn(b2, v));
    m = mi;
    v = vi;
    const float m_hat = mul_rn(mi, bc1), v_hat = mul_rn(vi, bc2);
    p = add_rn(p, -(mul_rn(lr, m_hat) / add_rn(sqrtf(v_hat), eps)));
}

// MCMC regulariser (optimizer/mcmc_densification.cpp:167-186): loss = lambda_o * mean(sigmoid(opacity)) +
// lambda_s * mean(exp(scale)); its gradient is closed-form per element and is added to the incoming
// gradient of the opacity / scale groups inside the Adam launch (the reference builds two autograd
// graphs and two extra tensors per step, training/trainer.cpp:232-237).
__device__ __forceinline__ float mcmc_reg_grad(int grp, float p, float reg_opa, float reg_scl) {
    if (grp == 2) {
        const float sg = 1.0f / (1.0f + expf(-p));
        return reg_opa * sg * (1.0f - sg);
    }
    if (grp == 3) return reg_scl * expf(p);
    return 0.0f;
}

__global__ void __launch_bounds__(256)
k_adam_multi(AdamGroups G, int64_t total_chunks, float b1, float b2, float eps, float bc1, float bc2,
             float grad_scale, float reg_opa /* lambda_o / N */, float reg_scl /* lambda_s / (3N) */,
             const StepDyn* __restrict__ dyn /* optional: per-step scalars in device memory (graph replay) */) {
    if (dyn != nullptr) {
        if (!dyn->ok) return;  // a frame of this step overflowed: the host re-runs the step
        bc1 = dyn->bc1;
        bc2 = dyn->bc2;
    }
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < total_chunks;
         c += (int64_t)gridDim.x * blockDim.x) {
        int grp = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (c >= G.chunk_end[k]) grp = k + 1;
        const int64_t local = c - (grp ? G.chunk_end[grp - 1] : 0);
        const int64_t e0 = local * 4;
        float* p = G.p[grp];
        const float* g = G.g[grp];
        float* m = G.m[grp];
        float* v = G.v[grp];
        const float lr = dyn ? dyn->lr[grp] : G.lr[grp];
        const bool reg = (reg_opa != 0.0f && grp == 2) || (reg_scl != 0.0f && grp == 3);
        if (e0 + 4 <= G.count[grp]) {
            float4 pv = *reinterpret_cast<float4*>(p + e0);
            float4 gv = __ldcs(reinterpret_cast<const float4*>(g + e0));
            if (reg) {  // grad_scale == 1 in this mode is not required: the regulariser is added after scaling
                gv.x = gv.x * grad_scale + mcmc_reg_grad(grp, pv.x, reg_opa, reg_scl);
                gv.y = gv.y * grad_scale + mcmc_reg_grad(grp, pv.y, reg_opa, reg_scl);
                gv.z = gv.z * grad_scale + mcmc_reg_grad(grp, pv.z, reg_opa, reg_scl);
                gv.w = gv.w * grad_scale + mcmc_reg_grad(grp, pv.w, reg_opa, reg_scl);
                float4 mv = *reinterpret_cast<float4*>(m + e0);
                float4 vv = *reinterpret_cast<float4*>(v + e0);
                adam_element(pv.x, gv.x, mv.x, vv.x, lr, b1, b2, eps, bc1, bc2);
                adam_element(pv.y, gv.y, mv.y, vv.y, lr, b1, b2, eps, bc1, bc2);
                adam_element(pv.z, gv.z, mv.z, vv.z, lr, b1, b2, eps, bc1, bc2);
                adam_element(pv.w, gv.w, mv.w, vv.w, lr, b1, b2, eps, bc1, bc2);
                *reinterpret_cast<float4*>(p + e0) = pv;
                *reinterpret_cast<float4*>(m + e0) = mv;
                *reinterpret_cast<float4*>(v + e0) = vv;
                continue;
            }
            float4 mv = *reinterpret_cast<float4*>(m + e0);
            float4 vv = *reinterpret_cast<float4*>(v + e0);
            adam_element(pv.x, gv.x * grad_scale, mv.x, vv.x, lr, b1, b2, eps, bc1, bc2);
            adam_element(pv.y, gv.y * grad_scale, mv.y, vv.y, lr, b1, b2, eps, bc1, bc2);
            adam_element(pv.z, gv.z * grad_scale, mv.z, vv.z, lr, b1, b2, eps, bc1, bc2);
            adam_element(pv.w, gv.w * grad_scale, mv.w, vv.w, lr, b1, b2, eps, bc1, bc2);
            *reinterpret_cast<float4*>(p + e0) = pv;
            *reinterpret_cast<float4*>(m + e0) = mv;
            *reinterpret_cast<float4*>(v + e0) = vv;
        } else {
            for (int64_t e = e0; e < G.count[grp]; ++e) {
                float pe = p[e], me = m[e], ve = v[e];
                const float ge = g[e] * grad_scale + (reg ? mcmc_reg_grad(grp, pe, reg_opa, reg_scl) : 0.0f);
                adam_element(pe, ge, me, ve, lr, b1, b2, eps, bc1, bc2);
                p[e] = pe; m[e] = me; v[e] = ve;
            }
        }
    }
}

// ================================================================================================
// MCMC position noise (optimizer/mcmc_densification.cpp:144-161): after the optimizer step
//   positions += noise_lr * exp(scales) * sigmoid(-k (sigmoid(opacity) - t)) * N(0, 1)
// One elementwise kernel (the reference: 6 libtorch kernels + a randn tensor); the normals come from
// Philox-4x32-10 keyed by (seed) with counter (Gaussian index, step), so every rank of a
// view-parallel run draws the same noise and the replicas stay bit-identical.
// ================================================================================================
__global__ void __launch_bounds__(256)
k_mcmc_noise(int64_t n, float* __restrict__ positions, const float* __restrict__ scales,
             const float* __restrict__ opacities, float noise_lr, float gate_k, float gate_t, unsigned seed_lo,
             unsigned seed_hi, unsigned step, float* __restrict__ normals_out /* optional [N,3] */,
             const StepDyn* __restrict__ dyn) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (dyn != nullptr) {
        if (!dyn->ok) return;
        noise_lr = dyn->noise_lr;
        step = dyn->step;
    }
    float z0, z1, z2;
    philox_normal3((unsigned)i, (unsigned)((uint64_t)i >> 32), step, 0x3c6ef372u, seed_lo, seed_hi, z0, z1, z2);
    const float sg = 1.0f / (1.0f + expf(-opacities[i]));
    const float gate = 1.0f / (1.0f + expf(gate_k * (sg - gate_t)));  // sigmoid(-k (sg - t))
    const float f = noise_lr * gate;
    positions[i * 3 + 0] += f * expf(scales[i * 3 + 0]) * z0;
    positions[i * 3 + 1] += f * expf(scales[i * 3 + 1]) * z1;
    positions[i * 3 + 2] += f * expf(scales[i * 3 + 2]) * z2;
    if (normals_out != nullptr) {
        normals_out[i * 3 + 0] = z0; normals_out[i * 3 + 1] = z1; normals_out[i * 3 + 2] = z2;
    }
}

// ================================================================================================
// densification statistics
// ================================================================================================
__global__ void __launch_bounds__(256)
k_accumulate_stats(int64_t n, const float* __restrict__ dL_dmeans_2d, const int* __restrict__ radii,
                   float* __restrict__ grad_accum, float* __restrict__ grad_count,
                   float* __restrict__ max_radii) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = radii[i];
    if (r > 0) {  // visible = radii > 0 (densification.cpp:71)
        const float2 g = reinterpret_cast<const float2*>(dL_dmeans_2d)[i];
        grad_accum[i] += sqrtf(g.x * g.x + g.y * g.y);
        grad_count[i] += 1.0f;
    }
    max_radii[i] = fmaxf(max_radii[i], (float)r);
}

}  // namespace cugs

using namespace cugs;

int cugs_adam_launch(cugs_handle_t* h, void* stream, float* const params[5], const float* const grads[5],
                     float* const m[5], float* const v[5], const int64_t counts[5], const float lr[5], float beta1,
                     float beta2, float eps, float bc1, float bc2, float grad_scale, float lambda_opacity,
                     float lambda_scale, const StepDyn* dyn);
int cugs_noise_launch(cugs_handle_t* h, void* stream, int64_t n, float* positions, const float* scales,
                      const float* opacities, float noise_lr, float gate_k, float gate_t, uint64_t seed,
                      uint32_t step, float* normals_out, const StepDyn* dyn);

extern "C" size_t cugs_b200_loss_workspace_bytes(int width, int height) {
    if (width <= 0 || height <= 0) return 64;
    return 64 + (size_t)3 * 3 * (size_t)width * (size_t)height * sizeof(float);
}

extern "C" int cugs_b200_loss_l1_ssim(cugs_handle_t* h, void* stream, int width, int height, float lambda,
                                      int window_size, const float* rendered, const float* target,
                                      float* dL_dcolor, float* scalars3, void* workspace, size_t workspace_bytes,
                                      float* ssim_map) {
    CUGS_REQUIRE(h, h != nullptr, "handle is null");
    CUGS_REQUIRE(h, width > 0 && height > 0, "image size must be positive");
    CUGS_REQUIRE(h, window_size % 2 == 1, "window_size must be odd");                 // loss.cpp:91
    CUGS_REQUIRE(h, window_size >= 3, "window_size must be >= 3");                    // loss.cpp:92
    if (window_size > kMaxWindow)
        return set_error(h, CUGS_ERR_UNSUPPORTED, "window_size %d > %d is not supported", window_size, kMaxWindow);
    CUGS_REQUIRE(h, rendered && target && scalars3 && workspace, "null pointer");
    if (workspace_bytes < cugs_b200_loss_workspace_bytes(width, height))
        return set_error(h, CUGS_ERR_WORKSPACE, "loss workspace too small: %zu < %zu", workspace_bytes,
                         cugs_b200_loss_workspace_bytes(width, height));
    static const SsimWindow win = make_window();
    cudaStream_t s = (cudaStream_t)stream;
    double* sums = reinterpret_cast<double*>(workspace);
    const size_t plane = (size_t)3 * width * height;
    float* g1 = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 64);
    float* g2 = g1 + plane;
    float* g3 = g2 + plane;
    CUGS_CUDA_TRY(h, cudaMemsetAsync(sums, 0, 64, s));
    if (window_size != 11) {  // generic window: the fall-back kernels
        SsimWindowAny wa;
        wa.size = window_size;
        for (int i = 0; i < kMaxWindow; ++i) wa.w[i] = 0.f;
        window_factor(window_size, wa.w);
        const unsigned blocks = (unsigned)((plane + 255) / 256);
        if (ssim_map) CUGS_CUDA_TRY(h, cudaMemsetAsync(ssim_map, 0, (size_t)width * height * sizeof(float), s));
        k_ssim_moments_any<<<blocks, 256, 0, s>>>(width, height, wa, rendered, target, g1, g2, g3, sums, ssim_map);
        CUGS_LAUNCH_CHECK(h, "k_ssim_moments_any");
        if (dL_dcolor) {
            k_ssim_gradient_any<<<blocks, 256, 0, s>>>(width, height, wa, lambda, rendered, target, g1, g2, g3,
                                                       dL_dcolor);
            CUGS_LAUNCH_CHECK(h, "k_ssim_gradient_any");
        }
        k_loss_finalize<<<1, 1, 0, s>>>(width, height, lambda, sums, scalars3);
        CUGS_LAUNCH_CHECK(h, "k_loss_finalize");
        return CUGS_OK;
    }
    const dim3 grid((width + kLossTile - 1) / kLossTile, (height + kLossTile - 1) / kLossTile);
    k_ssim_moments<<<grid, 256, 0, s>>>(width, height, win, rendered, target, g1, g2, g3, sums, ssim_map);
    CUGS_LAUNCH_CHECK(h, "k_ssim_moments");
    if (dL_dcolor) {
        k_ssim_gradient<<<grid, 256, 0, s>>>(width, height, win, lambda, rendered, target, g1, g2, g3, dL_dcolor);
        CUGS_LAUNCH_CHECK(h, "k_ssim_gradient");
    }
    k_loss_finalize<<<1, 1, 0, s>>>(width, height, lambda, sums, scalars3);
    CUGS_LAUNCH_CHECK(h, "k_loss_finalize");
    return CUGS_OK;
}

extern "C" int cugs_b200_adam_step(cugs_handle_t* h, void* stream, float* const params[5],
                                   const float* const grads[5], float* const m[5], float* const v[5],
                                   const int64_t counts[5], const float lr[5], float beta1, float beta2,
                                   float eps, float bc1, float bc2, float grad_scale) {
    return cugs_b200_adam_step_mcmc(h, stream, params, grads, m, v, counts, lr, beta1, beta2, eps, bc1, bc2,
                                    grad_scale, 0.0f, 0.0f);
}

extern "C" int cugs_b200_adam_step_mcmc(cugs_handle_t* h, void* stream, float* const params[5],
                                        const float* const grads[5], float* const m[5], float* const v[5],
                                        const int64_t counts[5], const float lr[5], float beta1, float beta2,
                                        float eps, float bc1, float bc2, float grad_scale, float lambda_opacity,
                                        float lambda_scale) {
    return cugs_adam_launch(h, stream, params, grads, m, v, counts, lr, beta1, beta2, eps, bc1, bc2, grad_scale,
                            lambda_opacity, lambda_scale, nullptr);
}

// dyn != NULL: learning rates and bias corrections are read from device memory (trainer.cu)
int cugs_adam_launch(cugs_handle_t* h, void* stream, float* const params[5], const float* const grads[5],
                     float* const m[5], float* const v[5], const int64_t counts[5], const float lr[5], float beta1,
                     float beta2, float eps, float bc1, float bc2, float grad_scale, float lambda_opacity,
                     float lambda_scale, const StepDyn* dyn) {
    CUGS_REQUIRE(h, h != nullptr, "handle is null");
    CUGS_REQUIRE(h, params && grads && m && v && counts && lr, "null pointer");
    AdamGroups G;
    int64_t chunks = 0;
    for (int k = 0; k < 5; ++k) {
        CUGS_REQUIRE(h, counts[k] >= 0, "negative count");
        CUGS_REQUIRE(h, counts[k] == 0 || (params[k] && grads[k] && m[k] && v[k]), "null group pointer");
        G.p[k] = params[k]; G.g[k] = grads[k]; G.m[k] = m[k]; G.v[k] = v[k];
        G.count[k] = counts[k];
        G.lr[k] = lr[k];
        chunks += (counts[k] + 3) / 4;
        G.chunk_end[k] = chunks;
    }
    if (chunks == 0) return CUGS_OK;
    int64_t blocks = (chunks + 255) / 256;
    const int64_t cap = (int64_t)h->sm_count * 8 * 4;  // persistent-ish grid, multiple of the SM count
    if (blocks > cap) blocks = cap;
    // mean over N opacities / 3N scale components (mcmc_densification.cpp:177-178)
    const float reg_opa = (lambda_opacity != 0.0f && counts[2] > 0) ? lambda_opacity / (float)counts[2] : 0.0f;
    const float reg_scl = (lambda_scale != 0.0f && counts[3] > 0) ? lambda_scale / (float)counts[3] : 0.0f;
    k_adam_multi<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(G, chunks, beta1, beta2, eps, bc1, bc2,
                                                                    grad_scale, reg_opa, reg_scl, dyn);
    CUGS_LAUNCH_CHECK(h, "k_adam_multi");
    return CUGS_OK;
}

extern "C" int cugs_b200_mcmc_inject_noise(cugs_handle_t* h, void* stream, int64_t n, float* positions,
                                           const float* scales, const float* opacities, float noise_lr,
                                           float gate_k, float gate_t, uint64_t seed, uint32_t step,
                                           float* normals_out) {
    return cugs_noise_launch(h, stream, n, positions, scales, opacities, noise_lr, gate_k, gate_t, seed, step,
                             normals_out, nullptr);
}

int cugs_noise_launch(cugs_handle_t* h, void* stream, int64_t n, float* positions, const float* scales,
                      const float* opacities, float noise_lr, float gate_k, float gate_t, uint64_t seed,
                      uint32_t step, float* normals_out, const StepDyn* dyn) {
    CUGS_REQUIRE(h, h != nullptr, "handle is null");
    CUGS_REQUIRE(h, n >= 0, "n must be >= 0");
    if (n == 0) return CUGS_OK;
    CUGS_REQUIRE(h, positions && scales && opacities, "null pointer");
    k_mcmc_noise<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        n, positions, scales, opacities, noise_lr, gate_k, gate_t, (unsigned)(seed & 0xffffffffu),
        (unsigned)(seed >> 32), step, normals_out, dyn);
    CUGS_LAUNCH_CHECK(h, "k_mcmc_noise");
    return CUGS_OK;
}

extern "C" int cugs_b200_accumulate_stats(cugs_handle_t* h, void* stream, int64_t n,
                                          const float* dL_dmeans_2d, const int32_t* radii,
                                          float* grad_accum, float* grad_count, float* max_radii) {
    CUGS_REQUIRE(h, h != nullptr, "handle is null");
    CUGS_REQUIRE(h, n >= 0, "n must be >= 0");
    if (n == 0) return CUGS_OK;
    CUGS_REQUIRE(h, dL_dmeans_2d && radii && grad_accum && grad_count && max_radii, "null pointer");
    k_accumulate_stats<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        n, dL_dmeans_2d, radii, grad_accum, grad_count, max_radii);
    CUGS_LAUNCH_CHECK(h, "k_accumulate_stats");
    return CUGS_OK;
}
