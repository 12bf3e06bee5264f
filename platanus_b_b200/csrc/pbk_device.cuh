// pbk_device.cuh -- device-side building blocks of the B200 k-mer counter (sm_100a).
//
// Data layout in HBM:
//   * read stream : 2-bit bases packed MSB-first into u64 words (32 bases per word; stream position
//                   p lives in word p/32 at bit 2*(31 - p%32)), reads concatenated without separators.
//                   Because a window that ends on a word boundary IS a key word of the reference
//                   layout (kmer.h:129-137, binstr.h:321-324: leftmost base most significant, word 0 =
//                   last 32 bases), the rolling forward k-mer needs no realignment.
//   * nflag/rflag : one bit per stream position (LSB-first in u32 words): "is N" and "first base
//                   of a read".  They replace SEQ::positionUnknown (common.h:468-476) and the per-read
//                   loop structure of countKmerPerThreadFirst (counter.h:405-432).
//   * table       : open addressing, linear probing, home = mulhi(hash, capacity), slot = key words +
//                   32-bit count/state.  Replaces the 1024 lock-striped DoubleHash sub-tables
//                   (counter.h:280-314, doubleHash.h:202-218).
#pragma once

#include <cstdint>
#ifndef PBK_CPU_EMUL
#include <cuda_runtime.h>
#endif

namespace pbk {

typedef unsigned long long u64;
typedef uint32_t u32;

constexpr u64 KEY_EMPTY = ~0ull;          // W == 1: never a canonical k-mer (T^32 > A^32 = its revcomp)
constexpr u32 CS_LOCKED = 0xFFFFFFFFu;    // W >= 2: slot claimed, key words being written
constexpr u32 COUNT_SAT = 65534u;         // counter.h:468
constexpr int MAX_PROBE = 192;            // longer probe runs go to the overflow list
constexpr int STREAM_PAD_WORDS = 16;      // zero words in front of stream / flag arrays
constexpr int PART_MAX_BUCKETS = 512;     // hash-range buckets of the partitioned path

template <int W>
struct alignas(8) Slot {
    u64 key[W];
    u32 cs;      // W == 1: count.  W >= 2: 0 = empty, CS_LOCKED = being written, else count
    u32 pad;
};

struct Counters {
    u64 instances;       // windows inserted (locally owned or staged)
    u64 new_keys;        // slots claimed in the main table
    u64 new_keys_remote; // slots claimed in the remote-staging table
    u64 overflow_n;      // keys appended to the overflow list
    u32 error_flags;     // ERR_*
    u32 pad;
};

enum : u32 { ERR_BAD_BASE = 1u, ERR_READ_TOO_LONG = 2u, ERR_OVERFLOW_LOST = 4u };

__host__ __device__ __forceinline__ u64 fmix64(u64 x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

template <int W>
__host__ __device__ __forceinline__ u64 hash_key(const u64 *key)
{
    u64 h = fmix64(key[0]);
#pragma unroll
    for (int j = 1; j < W; ++j) h = fmix64(h ^ key[j]);
    return h;
}

// owner shard: range partition of the LOW 32 hash bits (the table index uses the high bits)
__host__ __device__ __forceinline__ u32 shard_of_hash(u64 h, u32 n_shards)
{
    return (u32)(((h & 0xFFFFFFFFull) * (u64)n_shards) >> 32);
}

#ifndef PBK_CPU_EMUL
__device__ __forceinline__ u64 ld_cg_u64(const u64 *p)
{
    u64 v;
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ u32 ld_cg_u32(const u32 *p)
{
    u32 v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_add_u32(u32 *p, u32 v)
{
    asm volatile("red.global.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_cg_u64(u64 *p, u64 v)
{
    asm volatile("st.global.cg.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_release_u32(u32 *p, u32 v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

#else   // tests/cpu_emul: same source run sequentially on the host, see tests/cpu_emul/cuda_shim.h
inline u64 ld_cg_u64(const u64 *p) { return *p; }
inline u32 ld_cg_u32(const u32 *p) { return *p; }
inline void red_add_u32(u32 *p, u32 v) { *p += v; }
inline void st_cg_u64(u64 *p, u64 v) { *p = v; }
inline void st_release_u32(u32 *p, u32 v) { *p = v; }
#endif

// reverse the order of the 32 2-bit groups of a word
__device__ __forceinline__ u64 pair_reverse64(u64 x)
{
    x = __brevll(x);
    return ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
}

// numeric order of the reference: top word first (binstr.h:460-466); plain < for one word
template <int W>
__device__ __forceinline__ bool key_less(const u64 *a, const u64 *b)
{
#pragma unroll
    for (int j = W - 1; j > 0; --j)
        if (a[j] != b[j]) return a[j] < b[j];
    return a[0] < b[0];
}

// Insert `key` with weight `add` (>= 1).  Returns 1 if a new slot was claimed, 0 if an existing key was
// incremented, -1 if MAX_PROBE slots were tried without success (caller spills the key).
// Replaces countKmerOrWriteTemporary (counter.h:459-476): lock + find_times_any + insert/++/spill.
template <int W>
__device__ __forceinline__ int table_insert(Slot<W> *table, u64 cap, const u64 *key, u64 h, u32 add)
{
    u64 idx = __umul64hi(h, cap);          // capacity need not be a power of two
    if constexpr (W == 1) {
        const u64 k0 = key[0];
#pragma unroll 1
        for (int probe = 0; probe < MAX_PROBE; ++probe, idx = (idx + 1 == cap) ? 0 : idx + 1) {
            Slot<1> *s = table + idx;
            u64 cur = ld_cg_u64(&s->key[0]);
            if (cur == k0) { red_add_u32(&s->cs, add); return 0; }
            if (cur == KEY_EMPTY) {
                u64 old = atomicCAS(&s->key[0], KEY_EMPTY, k0);
                if (old == KEY_EMPTY) { red_add_u32(&s->cs, add); return 1; }
                if (old == k0) { red_add_u32(&s->cs, add); return 0; }
            }
        }
        return -1;
    } else {
        int probe = 0;
#pragma unroll 1
        while (probe < MAX_PROBE) {
            Slot<W> *s = table + idx;
            u32 cs = ld_cg_u32(&s->cs);
            if (cs == 0) {
                const u32 old = atomicCAS(&s->cs, 0u, CS_LOCKED);
                if (old == 0) {
#pragma unroll
                    for (int j = 0; j < W; ++j) st_cg_u64(&s->key[j], key[j]);
                    st_release_u32(&s->cs, add);          // key words become visible before the count
                    return 1;
                }
                cs = old;
            }
            // Claimed but not yet published: look at the SAME slot again.  (Written as a re-probe and not
            // as an inner `while (cs == LOCKED)` spin: nvcc deleted that loop because nothing after it
            // used `cs`, and threads then compared half-written keys.)
            if (cs == CS_LOCKED) continue;
            bool eq = true;
#pragma unroll
            for (int j = 0; j < W; ++j) eq &= (ld_cg_u64(&s->key[j]) == key[j]);
            if (eq) { red_add_u32(&s->cs, add); return 0; }
            ++probe;
            idx = (idx + 1 == cap) ? 0 : idx + 1;
        }
        return -1;
    }
}

template <int W>
__device__ __forceinline__ bool slot_occupied(const Slot<W> &s)
{
    if constexpr (W == 1) return s.key[0] != KEY_EMPTY;
    else return s.cs != 0;
}

__device__ __forceinline__ u64 warp_sum_u64(u64 v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ u32 warp_sum_u32(u32 v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace pbk
