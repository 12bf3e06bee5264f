// cuda_shim.h -- TEST ONLY.  Just enough of the CUDA device vocabulary to compile
// platanus_b_b200/csrc/pbk_kernels_impl.cuh with g++ and run each kernel as ONE sequential thread
// (grid = block = 1, a "warp" with a single live lane).  It checks the kernels' indexing and bit
// logic in the GPU-less build container; it is not a fallback and is never shipped or timed.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>

#define PBK_CPU_EMUL 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __grid_constant__
#define __launch_bounds__(...)
#define __shared__ static

struct dim3_ { unsigned x, y, z; };
static dim3_ blockIdx{0, 0, 0}, blockDim{1, 1, 1}, gridDim{1, 1, 1}, threadIdx{0, 0, 0};
struct uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
template <typename T> static inline T __ldg(const T *p) { return *p; }

static inline unsigned long long __brevll(unsigned long long x)
{
    unsigned long long r = 0;
    for (int i = 0; i < 64; ++i) r |= ((x >> i) & 1ull) << (63 - i);
    return r;
}
static inline int __clz(unsigned x) { return x ? __builtin_clz(x) : 32; }
static inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b)
{
    return (unsigned long long)(((unsigned __int128)a * b) >> 64);
}
template <typename T, typename U> static inline T atomicAdd(T *p, U v) { T o = *p; *p = (T)(o + v); return o; }
template <typename T, typename U> static inline T atomicOr(T *p, U v) { T o = *p; *p = (T)(o | v); return o; }
template <typename T, typename U, typename V> static inline T atomicCAS(T *p, U cmp, V val)
{
    T o = *p;
    if (o == (T)cmp) *p = (T)val;
    return o;
}
template <typename T, typename U> static inline T atomicMax(T *p, U v) { T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <typename T, typename U> static inline T atomicExch(T *p, U v) { T o = *p; *p = (T)v; return o; }
static inline void __threadfence() {}
static inline void __syncthreads() {}
static inline unsigned __ballot_sync(unsigned, bool p) { return p ? 1u : 0u; }
static inline bool __any_sync(unsigned, bool p) { return p; }
template <typename T> static inline unsigned __match_any_sync(unsigned, T) { return 1u; }
template <typename T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
template <typename T> static inline T __shfl_xor_sync(unsigned, T, int) { return (T)0; }   // other lanes hold 0
template <typename T> static inline T __shfl_up_sync(unsigned, T v, int) { return v; }    // lane 0 never uses the result
using std::max;
using std::min;
