// Microbenchmark: cycles per warp-instruction of the candidates for Pass A's per-key bucket counter on sm_100a:
// shared-memory atomicAdd on random counters, __match_any_sync, an 8-ballot emulation of match_any, and a
// non-atomic LDS+STS pair.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o warp_ops warp_ops.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t rnd(uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int MODE>
__global__ void k(uint32_t *out, long long *cyc, int iters, int nbuckets)
{
    extern __shared__ uint32_t cnt[];
    for (int i = threadIdx.x; i < nbuckets; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    uint32_t s = threadIdx.x * 2654435761u + blockIdx.x, acc = 0;
    const uint32_t lt = (1u << (threadIdx.x & 31)) - 1u;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        const uint32_t b = rnd(s) % nbuckets;
        if (MODE == 0) acc += atomicAdd(&cnt[b], 1u);
        else if (MODE == 1) { const unsigned m = __match_any_sync(0xffffffffu, b); acc += __popc(m & lt); }
        else if (MODE == 2) {
            unsigned m = 0xffffffffu;
#pragma unroll
            for (int bit = 0; bit < 8; ++bit) { const unsigned v = __ballot_sync(0xffffffffu, (b >> bit) & 1u); m &= ((b >> bit) & 1u) ? v : ~v; }
            acc += __popc(m & lt);
        } else if (MODE == 3) { const uint32_t c = cnt[b]; cnt[b] = c + 1; acc += c; }
        else if (MODE == 4) {            // ballot match + leader LDS/STS + shuffle: the full non-atomic rank
            unsigned m = 0xffffffffu;
#pragma unroll
            for (int bit = 0; bit < 8; ++bit) { const unsigned v = __ballot_sync(0xffffffffu, (b >> bit) & 1u); m &= ((b >> bit) & 1u) ? v : ~v; }
            const int leader = __ffs(m) - 1;
            uint32_t c = 0;
            if ((int)(threadIdx.x & 31) == leader) { c = cnt[b]; cnt[b] = c + __popc(m); }
            c = __shfl_sync(0xffffffffu, c, leader);
            acc += c + __popc(m & lt);
            __syncwarp();
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main()
{
    const int blocks = 148 * 2, threads = 256, iters = 4096, nb = 256;
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, blocks * threads * 4); cudaMallocManaged(&cyc, blocks * 8);
    const char *names[] = {"ATOMS.ADD random counter", "match_any", "8-ballot match", "LDS+STS (non-atomic)", "ballot match + leader LDS/STS + shfl"};
    for (int mode = 0; mode < 5; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            switch (mode) {
            case 0: k<0><<<blocks, threads, nb * 4>>>(out, cyc, iters, nb); break;
            case 1: k<1><<<blocks, threads, nb * 4>>>(out, cyc, iters, nb); break;
            case 2: k<2><<<blocks, threads, nb * 4>>>(out, cyc, iters, nb); break;
            case 3: k<3><<<blocks, threads, nb * 4>>>(out, cyc, iters, nb); break;
            case 4: k<4><<<blocks, threads, nb * 4>>>(out, cyc, iters, nb); break;
            }
            cudaDeviceSynchronize();
        }
        double avg = 0; for (int i = 0; i < blocks; ++i) avg += cyc[i]; avg /= blocks;
        // 2 CTAs x 8 warps per SM issue concurrently: SM-cycles per warp-instruction = cycles / (iters * 16 warps)
        printf("{\"op\": \"%s\", \"cta_cycles\": %.0f, \"sm_cycles_per_warp_op\": %.2f}\n", names[mode], avg, avg / (iters * 16.0));
    }
    return 0;
}
