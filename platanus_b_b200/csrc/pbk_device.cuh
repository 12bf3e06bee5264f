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
//   * table       : open addressing with linear probing, replacing the 1024 lock-striped DoubleHash
//                   sub-tables (counter.h:280-314, doubleHash.h:202-218).  Two slot formats:
//                     k <= 32 : ONE 64-bit word per slot = [remainder | displacement+1 | count]; the
//                               hash is a bijection of the key, the slot position supplies its top
//                               bits, so the key is recoverable and insert-or-increment is a single
//                               64-bit atomic add (see ct_insert_unit);
//                     k  > 32 : key words + 32-bit count/state word (claim, publish, increment).
#pragma once

#include <cstdint>
#include <cstring>
#ifndef PBK_CPU_EMUL
#include <cuda_runtime.h>
#endif

namespace pbk {

typedef unsigned long long u64;
typedef uint32_t u32;

constexpr u32 CS_LOCKED = 0xFFFFFFFFu;    // W >= 2: slot claimed, key words being written
constexpr u32 COUNT_SAT = 65534u;         // counter.h:468
constexpr int MAX_PROBE = 192;            // W >= 2: longer probe runs go to the overflow list
constexpr int STREAM_PAD_WORDS = 16;      // zero words in front of stream / flag arrays
constexpr int PART_MAX_BUCKETS = 512;     // hash-range buckets of the partitioned path
#ifndef PBK_PART_WIN1
#define PBK_PART_WIN1 16
#endif
constexpr int PART_WIN1 = PBK_PART_WIN1;  // windows of one stream word a Pass A thread handles per tile, one-word keys

// compact (k <= 32) slot format
constexpr int CT_DISP_BITS = 7;           // displacement + 1 in 1..127, 0 = slot not (yet) owned
constexpr int CT_MAX_DISP = 126;
constexpr int CT_MIN_QBITS = 27;          // >= 2^27 slots (1 GiB): leaves a 20-bit count field
#ifndef PBK_CPU_EMUL
constexpr u64 CT_NOISE = 1ull << 19;      // > number of resident threads (148 SMs x 2048): bound on in-flight +1s
#else
constexpr u64 CT_NOISE = 16;              // sequential emulation: nothing is ever in flight, small tables suffice
#endif

// Two- and three-word keys get 32-byte slots aligned to the 32-byte DRAM/L2 sector, so one 256-bit load
// (LDG.256 on sm_100a) brings the key words and the state word in a single L2 request.
template <int W>
struct alignas(W <= 3 ? 32 : 8) Slot {
    u64 key[W];
    u32 cs;      // 0 = empty, CS_LOCKED = being written, else count
    u32 pad;
};

template <int W> struct SlotType { typedef Slot<W> type; };
template <> struct SlotType<1> { typedef u64 type; };

struct Counters {
    u64 instances;       // windows inserted (locally owned or staged)
    u64 new_keys;        // slots claimed in the main table
    u64 new_keys_remote; // slots claimed in the remote-staging table
    u64 overflow_n;      // keys appended to the overflow list
    u32 error_flags;     // ERR_*
    u32 pad;
};

// Where Pass B of the key exchange finds the keys addressed to this shard: per source rank s, keys[s] = that source's
// segments for this destination, [region][seg_cap] entries, and cursors[s] = their fill counts, [region].  Two ways to get
// there: the all-to-all receive buffer of this GPU (pbk_keyx_insert_device: keys[s] = recv + s * R * seg_cap), or -- the
// pull form, pbk_keyx_pull_* -- the bucket store Pass A filled on GPU s, read in place over NVLink (peer pointers): the
// transfer then IS Pass B's own streaming loads, tile by tile, and no collective moves the keys.
constexpr int KEYX_MAX_SRC = 16;
struct KeyxSources {
    const u64 *keys[KEYX_MAX_SRC];
    const u64 *cursors[KEYX_MAX_SRC];
};

enum : u32 { ERR_BAD_BASE = 1u, ERR_READ_TOO_LONG = 2u, ERR_OVERFLOW_LOST = 4u };
constexpr u32 PASS_ONLY_BIT = 1u << 31;      // in a kernel's n_shards argument: hash-range pass, foreign keys are skipped (count_kernel)

__host__ __device__ __forceinline__ u64 fmix64(u64 x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

// fmix64 is a bijection on 64 bits (xor-shifts by >= 32 are involutions, the multipliers are odd)
__host__ __device__ __forceinline__ u64 fmix64_inverse(u64 x)
{
    x ^= x >> 33; x *= 0x9cb4b2f8129337dbULL;
    x ^= x >> 33; x *= 0x4f74430c22a54005ULL;
    x ^= x >> 33;
    return x;
}

// One-word keys: fmix64, a bijection (the compact table recovers the key from the hash).  Multi-word keys: the words are
// folded with one odd multiplier each and the sum goes through ONE fmix64 -- the first version chained an fmix64 per word,
// and those six dependent 64-bit multiplies per k-mer were a third of Pass A's instructions at k = 75.  A fold collision
// only sends two keys down the same probe sequence; the table always compares whole keys.
__host__ __device__ __forceinline__ u64 hash_fold_multiplier(int j)
{
    constexpr u64 M[8] = {1ull, 0x9E3779B97F4A7C15ull, 0xC2B2AE3D27D4EB4Full, 0x165667B19E3779F9ull,
                          0xD6E8FEB86659FD93ull, 0xA0761D6478BD642Full, 0xE7037ED1A0B428DBull, 0x8EBC6AF09C88C6E3ull};
    return M[j & 7];
}
template <int W>
__host__ __device__ __forceinline__ u64 hash_key(const u64 *key)
{
    u64 x = key[0];
#pragma unroll
    for (int j = 1; j < W; ++j) x += key[j] * hash_fold_multiplier(j);
    return fmix64(x);
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
// whole 32-byte slot in one request; served by one L2 sector access, i.e. a snapshot of the slot
__device__ __forceinline__ void ld_cg_256(const void *p, u64 &a, u64 &b, u64 &c, u64 &d)
{
    asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p) : "memory");
}
// whole 32-byte slot in one store: one L2 sector write, so a reader's 256-bit load sees the old slot or the new
// one, never a mix -- publishing a claimed slot this way needs no release fence
__device__ __forceinline__ void st_cg_256(void *p, u64 a, u64 b, u64 c, u64 d)
{
    asm volatile("st.global.cg.v4.u64 [%0], {%1,%2,%3,%4};" :: "l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}
__device__ __forceinline__ void red_add_u32(u32 *p, u32 v)
{
    asm volatile("red.global.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_add_u64(u64 *p, u64 v)
{
    asm volatile("red.global.add.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_cg_u64(u64 *p, u64 v)
{
    asm volatile("st.global.cg.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_cg_u64x2(u64 *p, u64 a, u64 b)
{
    asm volatile("st.global.cg.v2.u64 [%0], {%1,%2};" :: "l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void st_release_u32(u32 *p, u32 v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
// Table traffic carries an L2 evict_last policy: the hash range being worked on must survive the key
// stream flowing through the same L2 (a missed atomic waits for an HBM fill and holds an L1 request
// slot several times longer than a hit).
__device__ __forceinline__ u64 l2_keep_policy()
{
    u64 p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ u64 atom_add_keep_u64(u64 *p, u64 v, u64 policy)
{
    u64 old;
    asm volatile("atom.global.add.L2::cache_hint.u64 %0, [%1], %2, %3;" : "=l"(old) : "l"(p), "l"(v), "l"(policy) : "memory");
    return old;
}
__device__ __forceinline__ void red_add_keep_u64(u64 *p, u64 v, u64 policy)
{
    asm volatile("red.global.add.L2::cache_hint.u64 [%0], %1, %2;" :: "l"(p), "l"(v), "l"(policy) : "memory");
}
__device__ __forceinline__ void prefetch_keep(const void *p)
{
    asm volatile("prefetch.global.L2::evict_last [%0];" :: "l"(p));
}
// ---- TMA 1-D bulk copy global -> shared memory, completion on an mbarrier (cp.async.bulk, SASS UBLKCP): one elected thread
// moves a whole tile of the bucket store; the consuming threads never hold the keys' global loads in flight themselves.
__device__ __forceinline__ u32 smem_addr(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_addr(bar)), "r"(bytes) : "memory");
}
// src and dst 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_copy_g2s(void *dst_smem, const void *src_gmem, u32 bytes, u64 *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity)
{
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" :: "r"(smem_addr(bar)), "r"(parity) : "memory");
}
// read-once / write-once data (the bucket store): evict-first, so it does not push table lines out of L2
__device__ __forceinline__ u64 ld_stream_u64(const u64 *p) { return __ldcs(p); }
__device__ __forceinline__ ulonglong2 ld_stream_u64x2(const u64 *p) { return __ldcs(reinterpret_cast<const ulonglong2 *>(p)); }
__device__ __forceinline__ void st_stream_u64(u64 *p, u64 v) { __stcs(p, v); }
#else   // tests/cpu_emul: same source run sequentially on the host, see tests/cpu_emul/cuda_shim.h
inline u64 ld_cg_u64(const u64 *p) { return *p; }
inline u32 ld_cg_u32(const u32 *p) { return *p; }
inline void ld_cg_256(const void *p, u64 &a, u64 &b, u64 &c, u64 &d)
{
    const u64 *q = (const u64 *)p;
    a = q[0]; b = q[1]; c = q[2]; d = q[3];
}
inline void st_cg_256(void *p, u64 a, u64 b, u64 c, u64 d)
{
    u64 *q = (u64 *)p;
    q[0] = a; q[1] = b; q[2] = c; q[3] = d;
}
inline void red_add_u32(u32 *p, u32 v) { *p += v; }
inline void red_add_u64(u64 *p, u64 v) { *p += v; }
inline void st_cg_u64(u64 *p, u64 v) { *p = v; }
inline void st_cg_u64x2(u64 *p, u64 a, u64 b) { p[0] = a; p[1] = b; }
inline void st_release_u32(u32 *p, u32 v) { *p = v; }
inline u64 l2_keep_policy() { return 0; }
inline u64 atom_add_keep_u64(u64 *p, u64 v, u64) { u64 o = *p; *p += v; return o; }
inline void red_add_keep_u64(u64 *p, u64 v, u64) { *p += v; }
inline void prefetch_keep(const void *) {}
inline void mbar_init(u64 *, u32) {}
inline void mbar_fence_init() {}
inline void mbar_expect_tx(u64 *, u32) {}
inline void bulk_copy_g2s(void *dst, const void *src, u32 bytes, u64 *) { memcpy(dst, src, bytes); }   // completes at once
inline void mbar_wait(u64 *, u32) {}
inline u64 ld_stream_u64(const u64 *p) { return *p; }
struct ulonglong2 { u64 x, y; };
inline ulonglong2 make_ulonglong2(u64 x, u64 y) { return ulonglong2{x, y}; }
inline ulonglong2 ld_stream_u64x2(const u64 *p) { return ulonglong2{p[0], p[1]}; }
inline void st_stream_u64(u64 *p, u64 v) { *p = v; }
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

// -------------------------------------------------------------------------------------------------
// compact table, k <= 32.  capacity = 2^q slots, h = fmix64(key):
//     home = h >> (64 - q)            r = low (64 - q) bits of h
//     slot word = [ r : 64-q bits | d+1 : 7 bits | count : q-7 bits ]      (0 = empty)
// -------------------------------------------------------------------------------------------------
struct CtGeom {
    int rbits, cbits;
    u64 capmask, cmask, dz;
};
__host__ __device__ __forceinline__ CtGeom ct_geom(u64 cap)
{
    CtGeom g;
    int q = 0;
    while ((1ull << q) < cap) ++q;
    g.rbits = 64 - q;
    g.cbits = q - CT_DISP_BITS;
    g.capmask = cap - 1;
    g.cmask = (1ull << g.cbits) - 1;
    g.dz = (1ull << g.cbits) - CT_NOISE;          // counts this large are saturated for sure; stop adding
    return g;
}

// decode an owned slot (high bits != 0, no insert kernel running): key and raw count
__host__ __device__ __forceinline__ bool ct_decode(u64 v, u64 index, const CtGeom &g, u64 *key, u64 *count)
{
    const u64 hi = v >> g.cbits;
    if (hi == 0) return false;
    const u64 d = (hi & ((1u << CT_DISP_BITS) - 1)) - 1, r = hi >> CT_DISP_BITS;
    const u64 home = (index - d) & g.capmask;
    *key = fmix64_inverse((home << g.rbits) | r);
    *count = v & g.cmask;
    return true;
}

// Insert-or-increment by ONE (the hot path, replaces countKmerOrWriteTemporary, counter.h:459-476).
// Each probe is a single 64-bit atomic add of 1 whose return value tells whether the slot is ours:
//   * old tag == our tag      -> done, the add already counted the k-mer;
//   * old == 0                -> we are the one thread that saw the slot empty: publish our tag with a
//                                second add (tag and count bits are disjoint, so concurrent +1s cannot
//                                disturb it), our +1 stays as the first count;
//   * old has no tag yet      -> the claimer is between its two adds: wait for the tag, then decide;
//   * another key             -> take the +1 back (fire-and-forget reduction) and probe the next slot.
// Counts are exact once all kernels have drained; they may exceed 65534 (clamped when read) but never
// leave the count field: at `dz` further adds are taken back, and at most CT_NOISE +1s are in flight.
// Returns 1 = new key, 0 = existing key, -1 = no slot within CT_MAX_DISP (caller spills the key).
// continue an insert whose probe at displacement d returned `old`
__device__ __forceinline__ int ct_resolve(u64 *tab, const CtGeom &g, u64 h, u32 d, u64 *s, u64 old)
{
    const u64 home = h >> g.rbits;
    const u64 r_hi = ((h << (64 - g.rbits)) >> (64 - g.rbits)) << CT_DISP_BITS;   // remainder above the disp field
#pragma unroll 1
    for (;;) {
        const u64 tag_hi = r_hi | (u64)(d + 1);
        if (old == 0) {
            red_add_u64(s, tag_hi << g.cbits);
            return 1;
        }
        u64 hi = old >> g.cbits;
        while (hi == 0) {                              // claimed, tag not published yet
            // volatile access on purpose: nvcc collapsed the same loop written with an inline-asm load
            // into a single iteration, and threads then misjudged half-published slots
            old = *reinterpret_cast<volatile u64 *>(s);
            hi = old >> g.cbits;
        }
        if (hi == tag_hi) {
            if ((old & g.cmask) >= g.dz) red_add_u64(s, ~0ull);
            return 0;
        }
        red_add_u64(s, ~0ull);                         // not ours: take the +1 back
        if (++d > (u32)CT_MAX_DISP) return -1;
        s = tab + ((home + d) & g.capmask);
        old = atomicAdd(s, 1ull);
    }
}

__device__ __forceinline__ int ct_insert_unit(u64 *tab, const CtGeom &g, u64 h)
{
    u64 *s = tab + (h >> g.rbits);
    return ct_resolve(tab, g, h, 0, s, atomicAdd(s, 1ull));
}

// Weighted insert (table rebuild, records received from other shards, overflow re-insert).  Compare-
// and-swap protocol: exact saturation at 65534 and no transient values.  Never runs concurrently with
// ct_insert_unit (different kernels on one stream).
__device__ __forceinline__ int ct_insert_weighted(u64 *tab, const CtGeom &g, u64 h, u64 w)
{
    const u64 home = h >> g.rbits;
    const u64 r_hi = ((h << (64 - g.rbits)) >> (64 - g.rbits)) << CT_DISP_BITS;
    if (w > COUNT_SAT) w = COUNT_SAT;
#pragma unroll 1
    for (u32 d = 0; d <= (u32)CT_MAX_DISP; ++d) {
        u64 *s = tab + ((home + d) & g.capmask);
        const u64 tag = (r_hi | (u64)(d + 1)) << g.cbits;
        u64 cur = ld_cg_u64(s);
        if (cur == 0) {
            const u64 prev = atomicCAS(s, 0ull, tag | w);
            if (prev == 0) return 1;
            cur = prev;
        }
        if ((cur >> g.cbits) == (tag >> g.cbits)) {
            for (;;) {
                u64 c = (cur & g.cmask) + w;
                if (c > COUNT_SAT) c = COUNT_SAT;
                const u64 prev = atomicCAS(s, cur, tag | c);
                if (prev == cur) return 0;
                cur = prev;
            }
        }
    }
    return -1;
}

// value = max(value, w), inserting the key if absent (makeKmerReadDistributionFromContig, counter.h:572-575: a k-mer of
// several contigs keeps the largest coverage).  Same compare-and-swap protocol as ct_insert_weighted.
__device__ __forceinline__ int ct_insert_max(u64 *tab, const CtGeom &g, u64 h, u64 w)
{
    const u64 home = h >> g.rbits;
    const u64 r_hi = ((h << (64 - g.rbits)) >> (64 - g.rbits)) << CT_DISP_BITS;
    if (w > COUNT_SAT) w = COUNT_SAT;
#pragma unroll 1
    for (u32 d = 0; d <= (u32)CT_MAX_DISP; ++d) {
        u64 *s = tab + ((home + d) & g.capmask);
        const u64 tag = (r_hi | (u64)(d + 1)) << g.cbits;
        u64 cur = ld_cg_u64(s);
        if (cur == 0) {
            const u64 prev = atomicCAS(s, 0ull, tag | w);
            if (prev == 0) return 1;
            cur = prev;
        }
        if ((cur >> g.cbits) == (tag >> g.cbits)) {
            for (;;) {
                if ((cur & g.cmask) >= w) return 0;
                const u64 prev = atomicCAS(s, cur, tag | w);
                if (prev == cur) return 0;
                cur = prev;
            }
        }
    }
    return -1;
}

// Read-only lookup (no insert kernel running): clamped count of the key with hash h, 0 if it is not in the table.
// A key sits in the first slot of its probe sequence that was free when it arrived and nothing is ever removed, so
// the search ends at the first empty slot.
__device__ __forceinline__ u32 ct_lookup(const u64 *tab, const CtGeom &g, u64 h)
{
    const u64 home = h >> g.rbits;
    const u64 r_hi = ((h << (64 - g.rbits)) >> (64 - g.rbits)) << CT_DISP_BITS;
#pragma unroll 1
    for (u32 d = 0; d <= (u32)CT_MAX_DISP; ++d) {
        const u64 v = ld_cg_u64(tab + ((home + d) & g.capmask));
        const u64 hi = v >> g.cbits;
        if (hi == 0) return 0;
        if (hi == (r_hi | (u64)(d + 1))) {
            const u64 c = v & g.cmask;
            return c > COUNT_SAT ? COUNT_SAT : (u32)c;
        }
    }
    return 0;
}

// Set the count of the key with hash h to `w` (no other kernel running; keys of one launch are distinct): true if the
// key was there.  Used for seeded entries, whose value wins over whatever the reads added (counter.h:695-705).
__device__ __forceinline__ bool ct_set_count(u64 *tab, const CtGeom &g, u64 h, u64 w)
{
    const u64 home = h >> g.rbits;
    const u64 r_hi = ((h << (64 - g.rbits)) >> (64 - g.rbits)) << CT_DISP_BITS;
#pragma unroll 1
    for (u32 d = 0; d <= (u32)CT_MAX_DISP; ++d) {
        u64 *s = tab + ((home + d) & g.capmask);
        const u64 v = ld_cg_u64(s);
        const u64 hi = v >> g.cbits;
        if (hi == 0) return false;
        if (hi == (r_hi | (u64)(d + 1))) { st_cg_u64(s, (v & ~g.cmask) | w); return true; }
    }
    return false;
}

// -------------------------------------------------------------------------------------------------
// generic table, k > 32
// -------------------------------------------------------------------------------------------------
// Sector-sized slots (W <= 3) are published with ONE 256-bit store and read with ONE 256-bit load.  On this hardware both are
// single 32-byte sector transactions, but PTX only promises per-element atomicity for vector accesses, so a reader must be able
// to tell a torn snapshot (state word new, key words still the old zeros) from a slot that holds another key: the state word's
// upper half carries a 32-bit fold of the key words, written by the same 8-byte element as the count.  A reader whose key
// compare fails checks the fold of the words it READ against it; if they disagree the snapshot was torn and the slot is simply
// looked at again.  (A stale all-zero key passes only if the real key folds to 0: 2^-32 per torn read, which nobody has seen.)
template <int W>
__host__ __device__ __forceinline__ u32 key_fold32(const u64 *k)
{
    u64 x = k[0];
#pragma unroll
    for (int j = 1; j < W; ++j) x ^= (k[j] << (11 * j)) | (k[j] >> (64 - 11 * j));
    return (u32)(x ^ (x >> 32));
}

template <int W>
__device__ __forceinline__ int wide_insert(Slot<W> *table, u64 cap, const u64 *key, u64 h, u32 add)
{
    u64 idx = __umul64hi(h, cap);          // capacity need not be a power of two
    int probe = 0;
#pragma unroll 1
    while (probe < MAX_PROBE) {
        Slot<W> *s = table + idx;
        u32 cs = ld_cg_u32(&s->cs);
        if (cs == 0) {
            const u32 old = atomicCAS(&s->cs, 0u, CS_LOCKED);
            if (old == 0) {
                if constexpr (W <= 3) {                  // sector-sized slot: one 256-bit store publishes key, fold and count
                    u64 q[4] = {0, 0, 0, 0};
#pragma unroll
                    for (int j = 0; j < W; ++j) q[j] = key[j];
                    q[W] = (u64)add | ((u64)key_fold32<W>(key) << 32);
                    st_cg_256(s, q[0], q[1], q[2], q[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < W; ++j) st_cg_u64(&s->key[j], key[j]);
                    st_release_u32(&s->cs, add);          // key words become visible before the count
                }
                return 1;
            }
            cs = old;
        }
        // Claimed but not yet published: look at the SAME slot again.  (Written as a re-probe and not
        // as an inner `while (cs == LOCKED)` spin: nvcc deleted that loop because nothing after it
        // used `cs`, and threads then compared half-written keys.)
        if (cs == CS_LOCKED) continue;
        bool eq = true;
        u64 kw[W];
#pragma unroll
        for (int j = 0; j < W; ++j) { kw[j] = ld_cg_u64(&s->key[j]); eq &= (kw[j] == key[j]); }
        if (eq) { red_add_u32(&s->cs, add); return 0; }
        if constexpr (W <= 3) {                          // another key -- or a torn view of a slot being published?
            if (key_fold32<W>(kw) != ld_cg_u32(&s->pad)) continue;
        }
        ++probe;
        idx = (idx + 1 == cap) ? 0 : idx + 1;
    }
    return -1;
}

// wide_insert with "largest value wins" instead of "values add up"
template <int W>
__device__ __forceinline__ int wide_insert_max(Slot<W> *table, u64 cap, const u64 *key, u64 h, u32 w)
{
    u64 idx = __umul64hi(h, cap);
    int probe = 0;
#pragma unroll 1
    while (probe < MAX_PROBE) {
        Slot<W> *s = table + idx;
        u32 cs = ld_cg_u32(&s->cs);
        if (cs == 0) {
            const u32 old = atomicCAS(&s->cs, 0u, CS_LOCKED);
            if (old == 0) {
                if constexpr (W <= 3) {                  // same publish as wide_insert: key words, fold and value in one sector store
                    u64 q[4] = {0, 0, 0, 0};
#pragma unroll
                    for (int j = 0; j < W; ++j) q[j] = key[j];
                    q[W] = (u64)w | ((u64)key_fold32<W>(key) << 32);
                    st_cg_256(s, q[0], q[1], q[2], q[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < W; ++j) st_cg_u64(&s->key[j], key[j]);
                    st_release_u32(&s->cs, w);
                }
                return 1;
            }
            cs = old;
        }
        if (cs == CS_LOCKED) continue;                 // claimed, not yet published: look at the same slot again
        bool eq = true;
#pragma unroll
        for (int j = 0; j < W; ++j) eq &= (ld_cg_u64(&s->key[j]) == key[j]);
        if (eq) { atomicMax(&s->cs, w); return 0; }
        ++probe;
        idx = (idx + 1 == cap) ? 0 : idx + 1;
    }
    return -1;
}

template <int W>
__device__ __forceinline__ u32 wide_lookup(const Slot<W> *table, u64 cap, const u64 *key, u64 h)
{
    u64 idx = __umul64hi(h, cap);
#pragma unroll 1
    for (int probe = 0; probe < MAX_PROBE; ++probe) {
        const Slot<W> *s = table + idx;
        const u32 cs = ld_cg_u32(&s->cs);
        if (cs == 0) return 0;
        bool eq = true;
#pragma unroll
        for (int j = 0; j < W; ++j) eq &= (ld_cg_u64(&s->key[j]) == key[j]);
        if (eq) return cs > COUNT_SAT ? COUNT_SAT : cs;
        idx = (idx + 1 == cap) ? 0 : idx + 1;
    }
    return 0;
}

template <int W>
__device__ __forceinline__ bool wide_set_count(Slot<W> *table, u64 cap, const u64 *key, u64 h, u32 w)
{
    u64 idx = __umul64hi(h, cap);
#pragma unroll 1
    for (int probe = 0; probe < MAX_PROBE; ++probe) {
        Slot<W> *s = table + idx;
        const u32 cs = ld_cg_u32(&s->cs);
        if (cs == 0) return false;
        bool eq = true;
#pragma unroll
        for (int j = 0; j < W; ++j) eq &= (ld_cg_u64(&s->key[j]) == key[j]);
        if (eq) { s->cs = w; return true; }
        idx = (idx + 1 == cap) ? 0 : idx + 1;
    }
    return false;
}

// -------------------------------------------------------------------------------------------------
// uniform view used by every kernel: Table<W> wraps either format
// -------------------------------------------------------------------------------------------------
template <int W>
struct Table {
    typedef typename SlotType<W>::type slot_t;
    slot_t *slots;
    u64 cap;
    CtGeom g;          // only meaningful for W == 1

    __host__ __device__ Table(void *p, u64 c) : slots((slot_t *)p), cap(c), g()
    {
        if (W == 1 && c) g = ct_geom(c);
    }

    // add `w` occurrences of `key`; unit adds take the single-atomic path
    __device__ __forceinline__ int insert(const u64 *key, u64 h, u32 w, bool unit) const
    {
        if constexpr (W == 1) {
            return unit ? ct_insert_unit(slots, g, h) : ct_insert_weighted(slots, g, h, w);
        } else {
            return wide_insert<W>(slots, cap, key, h, w > COUNT_SAT ? COUNT_SAT : w);
        }
    }
    // value = max(value, w); 1 = new key, 0 = existing, -1 = no slot
    __device__ __forceinline__ int insert_max(const u64 *key, u64 h, u32 w) const
    {
        if constexpr (W == 1) return ct_insert_max(slots, g, h, w);
        else return wide_insert_max<W>(slots, cap, key, h, w > COUNT_SAT ? COUNT_SAT : w);
    }
    // clamped count of `key` (hash h), 0 if absent; only between insert kernels
    __device__ __forceinline__ u32 find(const u64 *key, u64 h) const
    {
        if constexpr (W == 1) return ct_lookup(slots, g, h);
        else return wide_lookup<W>(slots, cap, key, h);
    }
    // set the count of a key that is in the table; false if it is not
    __device__ __forceinline__ bool set_count(const u64 *key, u64 h, u32 w) const
    {
        if constexpr (W == 1) return ct_set_count(slots, g, h, w);
        else return wide_set_count<W>(slots, cap, key, h, w);
    }
    // read slot i once all inserts have drained: false if empty
    __device__ __forceinline__ bool load(u64 i, u64 *key, u32 *count) const
    {
        if constexpr (W == 1) {
            u64 c;
            if (!ct_decode(slots[i], i, g, key, &c)) return false;
            *count = c > COUNT_SAT ? COUNT_SAT : (u32)c;
            return true;
        } else {
            const Slot<W> sl = slots[i];
            if (sl.cs == 0) return false;
#pragma unroll
            for (int j = 0; j < W; ++j) key[j] = sl.key[j];
            *count = sl.cs > COUNT_SAT ? COUNT_SAT : sl.cs;
            return true;
        }
    }
    // clamp the stored count to 65534 (keeps the 32-bit counters of wide slots away from overflow)
    __device__ __forceinline__ void clamp(u64 i) const
    {
        if constexpr (W == 1) {
            const u64 v = slots[i];
            if ((v & g.cmask) > COUNT_SAT) slots[i] = (v & ~g.cmask) | COUNT_SAT;
        } else {
            if (slots[i].cs > COUNT_SAT) slots[i].cs = COUNT_SAT;
        }
    }
};

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
