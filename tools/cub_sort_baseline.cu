// cub_sort_baseline.cu — BENCHMARK BASELINE ONLY (never linked into libcugs_b200.so).
//
// The sort exactly as the reference calls it (rasterizer/sorting.cu:190-211:
// cub::DeviceRadixSort::SortPairs on (uint64 key, int32 value), all 64 bits) plus the same call with
// begin_bit/end_bit trimmed, behind a C ABI so tools/sort_bench.py can time CUB, trimmed CUB and this
// repository's hand-written onesweep on identical device buffers. SURVEY §2.2 sets the bar: beat CUB
// as called, and report against end_bit = 32 + ceil(log2(tiles)).
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC \
//        tools/cub_sort_baseline.cu -o tools/_build/libcub_sort_baseline.so
#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>
#include <stdint.h>

extern "C" size_t cub_sort_pairs_temp_bytes(int64_t p, int begin_bit, int end_bit) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                    (const int32_t*)nullptr, (int32_t*)nullptr, p, begin_bit, end_bit);
    return bytes;
}

extern "C" int cub_sort_pairs(void* temp, size_t temp_bytes, const uint64_t* keys_in, uint64_t* keys_out,
                              const int32_t* vals_in, int32_t* vals_out, int64_t p, int begin_bit, int end_bit,
                              void* stream) {
    return (int)cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, vals_in, vals_out, p, begin_bit,
                                                end_bit, (cudaStream_t)stream);
}
