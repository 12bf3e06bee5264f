// TEST ONLY: the one piece of pbk_kernels.cu that is library code on the GPU (CUB radix sort) in host form, for the emulated
// C ABI (see cuda_rt_shim.h).  Same contract as sort_export in pbk_kernels.cu: ascending in the reference's key order.
#include "../../platanus_b_b200/csrc/pbk_kernels.cuh"

#include <algorithm>
#include <numeric>
#include <vector>

namespace pbk {

size_t sort_export_scratch_bytes(u64, int) { return 0; }

cudaError_t sort_export(u64 *keys, uint16_t *counts, u64 n, int words, int, void *, size_t, cudaStream_t)
{
    std::vector<u64> perm(n);
    std::iota(perm.begin(), perm.end(), 0ull);
    std::sort(perm.begin(), perm.end(), [&](u64 a, u64 b) {
        for (int j = words - 1; j >= 0; --j)
            if (keys[a * words + j] != keys[b * words + j]) return keys[a * words + j] < keys[b * words + j];
        return false;
    });
    std::vector<u64> k2((size_t)n * words);
    std::vector<uint16_t> c2(n);
    for (u64 i = 0; i < n; ++i) {
        for (int j = 0; j < words; ++j) k2[i * words + j] = keys[perm[i] * words + j];
        c2[i] = counts[perm[i]];
    }
    std::copy(k2.begin(), k2.end(), keys);
    std::copy(c2.begin(), c2.end(), counts);
    return cudaSuccess;
}

void launch_microbench(void *, int, u64, int, u64, int, cudaStream_t) {}

}  // namespace pbk
