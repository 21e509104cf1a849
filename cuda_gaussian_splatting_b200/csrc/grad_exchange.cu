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
// | rotations [M,4].
#include "common.cuh"

namespace cugs {

struct GradGroups {
    float* g[5];  // positions, sh_coeffs, opacities, scales, rotations
};

template <bool kGather>
__global__ void __launch_bounds__(256)
k_move_grad_rows(int64_t n, int C, const int* __restrict__ touch, const int* __restrict__ offsets, int64_t m,
                 GradGroups dense, float* __restrict__ compact) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (i >= n || touch[i] == 0) return;
    const int64_t j = offsets[i];
    const int shw = 3 * C;
    const int row = shw + 11;
    const int64_t b_sh = 3 * m, b_op = b_sh + (int64_t)shw * m, b_sc = b_op + m, b_ro = b_sc + 3 * m;
    for (int e = lane; e < row; e += 32) {
        float* d;
        float* c;
        if (e < 3) { d = dense.g[0] + i * 3 + e; c = compact + j * 3 + e; }
        else if (e < 3 + shw) { d = dense.g[1] + i * shw + (e - 3); c = compact + b_sh + j * shw + (e - 3); }
        else if (e < 4 + shw) { d = dense.g[2] + i; c = compact + b_op + j; }
        else if (e < 7 + shw) { d = dense.g[3] + i * 3 + (e - 4 - shw); c = compact + b_sc + j * 3 + (e - 4 - shw); }
        else { d = dense.g[4] + i * 4 + (e - 7 - shw); c = compact + b_ro + j * 4 + (e - 7 - shw); }
        if (kGather) *c = *d;
        else *d = *c;
    }
}

}  // namespace cugs

using namespace cugs;

static int move_rows(cugs_handle_t* h, void* stream, int64_t n, int num_coeffs, const int32_t* touch,
                     const int32_t* offsets, int64_t m, float* const grads[5], float* compact, bool gather) {
    CUGS_REQUIRE(h, h != nullptr, "handle is null");
    CUGS_REQUIRE(h, n >= 0 && m >= 0 && m <= n, "bad n / m");
    CUGS_REQUIRE(h, num_coeffs >= 1 && num_coeffs <= 64, "bad num_coeffs");
    if (n == 0 || m == 0) return CUGS_OK;
    CUGS_REQUIRE(h, touch && offsets && grads && compact, "null pointer");
    GradGroups G;
    for (int k = 0; k < 5; ++k) {
        CUGS_REQUIRE(h, grads[k] != nullptr, "null gradient group");
        G.g[k] = grads[k];
    }
    const unsigned grid = (unsigned)((n + 7) / 8);
    if (gather)
        k_move_grad_rows<true><<<grid, 256, 0, (cudaStream_t)stream>>>(n, num_coeffs, touch, offsets, m, G, compact);
    else
        k_move_grad_rows<false><<<grid, 256, 0, (cudaStream_t)stream>>>(n, num_coeffs, touch, offsets, m, G, compact);
    CUGS_LAUNCH_CHECK(h, "k_move_grad_rows");
    return CUGS_OK;
}

extern "C" int cugs_b200_gather_grad_rows(cugs_handle_t* h, void* stream, int64_t n, int num_coeffs,
                                          const int32_t* touch, const int32_t* offsets, int64_t m,
                                          const float* const grads[5], float* compact) {
    return move_rows(h, stream, n, num_coeffs, touch, offsets, m, const_cast<float* const*>(grads), compact, true);
}

extern "C" int cugs_b200_scatter_grad_rows(cugs_handle_t* h, void* stream, int64_t n, int num_coeffs,
                                           const int32_t* touch, const int32_t* offsets, int64_t m,
                                           const float* compact, float* const grads[5]) {
    return move_rows(h, stream, n, num_coeffs, touch, offsets, m, grads, const_cast<float*>(compact), false);
}
