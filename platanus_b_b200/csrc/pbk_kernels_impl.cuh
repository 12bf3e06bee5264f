// pbk_kernels_impl.cuh -- the kernels themselves (no launch syntax), so that tests/cpu_emul can
// run the very same source sequentially on the host as a logic check.  Launch wrappers: pbk_kernels.cu.
#pragma once
#include "pbk_device.cuh"

namespace pbk {

// =================================================================================================
// ingest: ASCII -> 2-bit stream (SEQ::convertFromString + Char2Bin, common.h:256, 460-477)
// =================================================================================================

// four characters per 32-bit lane.  Char2Bin only looks at the low nibble: 1->A 3->C 7->G 4->T
// 14->N 15->0(A); every other nibble has no defined code in the reference.
// `platanus` carries the encoding (bit 0: PBK_ENC_PLATANUS code bytes) and bit 8: characters without a Char2Bin code are
// treated as N instead of being reported (PBK_F_UNKNOWN_AS_N: IUPAC ambiguity codes occur in real FASTQ files)
__device__ __forceinline__ void pack4(u32 v, int mode, u32 &byte_out, u32 &nflag4, u32 &bad)
{
    const bool platanus = (mode & 1) != 0;
    if (platanus) {
        u32 codes = v & 0x03030303u;
        byte_out = (codes * 0x40100401u) >> 24;
        nflag4 = 0;
        return;
    }
    u32 codes = ((v >> 1) ^ (v >> 2)) & 0x03030303u;
    byte_out = (codes * 0x40100401u) >> 24;                  // c0<<6 | c1<<4 | c2<<2 | c3
    u32 eq = (v & 0x0F0F0F0Fu) ^ 0x0E0E0E0Eu;                // zero byte <=> nibble 14 ('N','n')
    u32 z = ~(eq + 0x7F7F7F7Fu) & 0x80808080u;
    nflag4 = (((z >> 7) * 0x01020408u) >> 24) & 0xFu;        // bit i <=> byte i
    u32 b0 = v, b1 = v >> 1, b2 = v >> 2, b3 = v >> 3;
    u32 ok = (~b3 & ~b2 & b0) | (~b3 & b2 & ~(b1 ^ b0)) | (b3 & b2 & b1);
    const u32 notok = (~ok) & 0x01010101u;
    if (mode & 0x100) nflag4 |= ((notok * 0x01020408u) >> 24) & 0xFu;
    else bad |= notok;
}

template <bool ALIGNED>
__global__ void __launch_bounds__(256)
pack_kernel(const uint8_t *__restrict__ bases, u64 n_valid, u64 n_words, int platanus,
            u64 *__restrict__ stream, u32 *__restrict__ nflag, u64 word0, Counters *ctr)
{
    u32 bad = 0;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += stride) {
        const u64 b0 = w * 32;
        u32 x[8];
        u32 tail_flags = 0;
        if (ALIGNED && b0 + 32 <= n_valid) {
            const uint4 *p = reinterpret_cast<const uint4 *>(bases + b0);
            uint4 lo = __ldg(p), hi = __ldg(p + 1);
            x[0] = lo.x; x[1] = lo.y; x[2] = lo.z; x[3] = lo.w;
            x[4] = hi.x; x[5] = hi.y; x[6] = hi.z; x[7] = hi.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                u32 v = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    u64 pos = b0 + 4 * j + b;
                    u32 c;
                    if (pos < n_valid) c = bases[pos];
                    else { c = (platanus & 1) ? 0u : (u32)'A'; tail_flags |= 1u << (4 * j + b); }
                    v |= c << (8 * b);
                }
                x[j] = v;
            }
        }
        u64 word = 0;
        u32 nf = tail_flags;        // positions past the end can never be inside a window
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            u32 byte, f4;
            pack4(x[j], platanus, byte, f4, bad);
            word |= (u64)byte << (56 - 8 * j);
            nf |= f4 << (4 * j);
        }
        stream[word0 + w] = word;
        nflag[word0 + w] = nf;
    }
    if (__any_sync(0xffffffffu, bad != 0) && (threadIdx.x & 31) == 0)
        atomicOr(&ctr->error_flags, ERR_BAD_BASE);
}


// per read: ++lengthDistribution[len] (counter.h:406), ReadError (common.h:465), first-base flag.
// Fixed-length read sets put every increment on ONE histogram bin, so lengths are aggregated per warp
// (match_any) and then per CTA before touching global memory.
__global__ void __launch_bounds__(256)
read_marks_kernel(const u64 *__restrict__ off, u64 n_reads, u64 *len_hist, u32 *rflag, Counters *ctr)
{
    __shared__ u64 w_len[8];
    __shared__ u32 w_cnt[8];
    const u64 stride = (u64)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u64 n_round = (n_reads + blockDim.x - 1) / blockDim.x * blockDim.x;
    for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < n_round; r += stride) {
        const bool active = r < n_reads;
        u64 start = 0, len = 0;
        if (active) { start = off[r]; len = off[r + 1] - start; }
        if (active && len >= 500000ull) { atomicOr(&ctr->error_flags, ERR_READ_TOO_LONG); len = 500000ull; }
        const unsigned act = __ballot_sync(0xffffffffu, active);
        if (threadIdx.x < 8) w_cnt[threadIdx.x] = 0;
        __syncthreads();
        if (active) {
            const unsigned peers = __match_any_sync(act, len);
            if (lane == __ffs(peers) - 1) {
                if (lane == __ffs(act) - 1 && warp < 8) { w_len[warp] = len; w_cnt[warp] = (u32)__popc(peers); }
                else atomicAdd(&len_hist[len], (u64)__popc(peers));
            }
            if (len > 0) atomicOr(&rflag[start >> 5], 1u << (start & 31));
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const int nw = min((int)((blockDim.x + 31) >> 5), 8);
            for (int a = 0; a < nw; ++a) {
                if (!w_cnt[a]) continue;
                u64 total = w_cnt[a];
                for (int b = a + 1; b < nw; ++b)
                    if (w_cnt[b] && w_len[b] == w_len[a]) { total += w_cnt[b]; w_cnt[b] = 0; }
                atomicAdd(&len_hist[w_len[a]], total);
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
npos_scatter_kernel(const u64 *__restrict__ off, const int32_t *__restrict__ n_pos,
                    const u64 *__restrict__ n_pos_off, u64 n_reads, u32 *nflag)
{
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += stride) {
        const u64 base = off[r], len = off[r + 1] - base;
        for (u64 i = n_pos_off[r]; i < n_pos_off[r + 1]; ++i) {
            const int32_t rel = n_pos[i];
            if (rel < 0 || (u64)rel >= len) continue;
            const u64 p = base + (u64)rel;
            atomicOr(&nflag[p >> 5], 1u << (p & 31));
        }
    }
}


// PBK_ENC_PACKED2: N positions given as stream positions of the batch; positions past the last base (the zero padding of the
// last word) are flagged too, so that no window can end there (pack_kernel does the same for its tail)
__global__ void __launch_bounds__(256)
npos_abs_scatter_kernel(const u64 *__restrict__ n_positions, u64 n_n, u64 n_bases, u32 *nflag)
{
    const u64 stride = (u64)gridDim.x * blockDim.x, gtid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    for (u64 i = gtid; i < n_n; i += stride) {
        const u64 p = n_positions[i];
        if (p < n_bases) atomicOr(&nflag[p >> 5], 1u << (p & 31));
    }
    if (gtid == 0 && (n_bases & 31)) atomicOr(&nflag[n_bases >> 5], ~0u << (n_bases & 31));
}

// =================================================================================================
// counting: rolling canonical k-mers (counter.h:413-429) + table insert (counter.h:459-476)
// =================================================================================================

template <int W>
__device__ __forceinline__ void spill_key(const u64 *key, Counters *ctr, u64 *ovf, u64 ovf_cap)
{
    u64 at = atomicAdd(&ctr->overflow_n, 1ull);
    if (at < ovf_cap) {
#pragma unroll
        for (int j = 0; j < W; ++j) ovf[at * (W + 1) + j] = key[j];
        ovf[at * (W + 1) + W] = 1;
    } else {
        atomicOr(&ctr->error_flags, ERR_OVERFLOW_LOST);
    }
}

// One thread owns one stream word: the 32 windows that END at its 32 positions.  Because the stream
// is packed MSB-first and the word boundary is a window end, the k-1 bases before the word are simply
// the previous W stream words -- the rolling state is primed without a single shift.
template <int W>
__global__ void __launch_bounds__(256)
count_kernel(const u64 *__restrict__ stream, const u32 *__restrict__ nflag, const u32 *__restrict__ rflag,
             u64 word_begin, u64 word_end, int k, Table<W> table, Table<W> remote,
             u32 n_shards_in, u32 rank, Counters *ctr, u64 *ovf, u64 ovf_cap)
{
    // PASS_ONLY_BIT: a hash-range pass over the whole input on one GPU (pbk_config.n_passes): keys of the other ranges are
    // skipped -- neither counted nor staged -- because a later pass over the same reads picks them up
    const bool pass_only = (n_shards_in & PASS_ONLY_BIT) != 0;
    const u32 n_shards = n_shards_in & ~PASS_ONLY_BIT;
    const int top_shift = 2 * ((k - 1) & 31);
    const u64 top_mask = (k & 31) ? ((1ull << (2 * (k & 31))) - 1ull) : ~0ull;
    const int s = 2 * (32 * W - k);                   // 0..62
    const int nb = (k + 30) >> 5;                     // flag words that cover the k-1 previous positions
    const u64 stride = (u64)gridDim.x * blockDim.x;
    u64 inst = 0;
    u32 newk = 0, newr = 0;

    for (u64 wi = word_begin + (u64)blockIdx.x * blockDim.x + threadIdx.x; wi < word_end; wi += stride) {
        const u64 cur = stream[wi];
        const u32 nf = nflag[wi], rf = rflag[wi];

        // run = number of consecutive usable bases (same read, no N) ending just before this word
        int run = k;
        for (int j = 1; j <= nb; ++j) {
            const u32 a = nflag[wi - j], b = rflag[wi - j];
            if (a | b) {
                const int pn = a ? 32 - __clz(a) : 0;  // first position after the last N
                const int pr = b ? 31 - __clz(b) : 0;  // the last read start itself is usable
                run = 32 * j - max(pn, pr);
                break;
            }
        }

        u64 fwd[W], rev[W];
#pragma unroll
        for (int j = 0; j < W; ++j) fwd[j] = stream[wi - 1 - j];
        {
            u64 y[W + 1];
#pragma unroll
            for (int j = 0; j < W; ++j) y[j] = pair_reverse64(~fwd[W - 1 - j]);
            y[W] = 0;
#pragma unroll
            for (int j = 0; j < W; ++j) rev[j] = s ? ((y[j] >> s) | (y[j + 1] << (64 - s))) : y[j];
        }

#pragma unroll 2
        for (int i = 0; i < 32; ++i) {
            const u32 b = (u32)(cur >> (62 - 2 * i)) & 3u;
            if ((rf >> i) & 1u) run = 0;
            run = ((nf >> i) & 1u) ? 0 : run + 1;
#pragma unroll
            for (int j = W - 1; j > 0; --j) fwd[j] = (fwd[j] << 2) | (fwd[j - 1] >> 62);
            fwd[0] = (fwd[0] << 2) | b;
            fwd[W - 1] &= top_mask;
#pragma unroll
            for (int j = 0; j < W - 1; ++j) rev[j] = (rev[j] >> 2) | (rev[j + 1] << 62);
            rev[W - 1] = (rev[W - 1] >> 2) | ((u64)(3u - b) << top_shift);

            if (run >= k) {
                const bool use_rev = key_less<W>(rev, fwd);        // key = min(forward, reverse), counter.h:429
                u64 key[W];
#pragma unroll
                for (int j = 0; j < W; ++j) key[j] = use_rev ? rev[j] : fwd[j];
                const u64 h = hash_key<W>(key);
                const bool foreign = n_shards > 1 && shard_of_hash(h, n_shards) != rank;
                if (foreign && pass_only) continue;
                ++inst;
                int r;
                if (foreign) {
                    r = remote.insert(key, h, 1u, true);
                    newr += (r > 0);
                } else {
                    r = table.insert(key, h, 1u, true);
                    newk += (r > 0);
                }
                if (r < 0) spill_key<W>(key, ctr, ovf, ovf_cap);
            }
        }
    }
    inst = warp_sum_u64(inst);
    newk = warp_sum_u32(newk);
    newr = warp_sum_u32(newr);
    if ((threadIdx.x & 31) == 0) {
        if (inst) atomicAdd(&ctr->instances, inst);
        if (newk) atomicAdd(&ctr->new_keys, (u64)newk);
        if (newr) atomicAdd(&ctr->new_keys_remote, (u64)newr);
    }
}

// Occurrence lookup (ContigDivider::getOccurrenceArray, kmer_divide.cpp:151-197; the table probe of
// divideKmerUsedMakingPreviousContig, counter.h:828-861): same window enumeration as count_kernel, but the canonical key
// is only looked up.  occ[p] = clamped count of the window that ENDS at stream position p, 0 if the window is not usable
// (contains an N, crosses a read start) or its k-mer is not in the table.  A thread writes its 32 results as eight
// 8-byte words.
template <int W>
__global__ void __launch_bounds__(256)
lookup_kernel(const u64 *__restrict__ stream, const u32 *__restrict__ nflag, const u32 *__restrict__ rflag,
              u64 word_begin, u64 word_end, int k, Table<W> table, u64 *__restrict__ occ4)
{
    const int top_shift = 2 * ((k - 1) & 31);
    const u64 top_mask = (k & 31) ? ((1ull << (2 * (k & 31))) - 1ull) : ~0ull;
    const int s = 2 * (32 * W - k);
    const int nb = (k + 30) >> 5;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 wi = word_begin + (u64)blockIdx.x * blockDim.x + threadIdx.x; wi < word_end; wi += stride) {
        const u64 cur = stream[wi];
        const u32 nf = nflag[wi], rf = rflag[wi];
        int run = k;
        for (int j = 1; j <= nb; ++j) {
            const u32 a = nflag[wi - j], b = rflag[wi - j];
            if (a | b) {
                const int pn = a ? 32 - __clz(a) : 0;
                const int pr = b ? 31 - __clz(b) : 0;
                run = 32 * j - max(pn, pr);
                break;
            }
        }
        u64 fwd[W], rev[W];
#pragma unroll
        for (int j = 0; j < W; ++j) fwd[j] = stream[wi - 1 - j];
        {
            u64 y[W + 1];
#pragma unroll
            for (int j = 0; j < W; ++j) y[j] = pair_reverse64(~fwd[W - 1 - j]);
            y[W] = 0;
#pragma unroll
            for (int j = 0; j < W; ++j) rev[j] = s ? ((y[j] >> s) | (y[j + 1] << (64 - s))) : y[j];
        }
#pragma unroll 1
        for (int g = 0; g < 8; ++g) {
            u64 packed = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int i = 4 * g + q;
                const u32 b = (u32)(cur >> (62 - 2 * i)) & 3u;
                if ((rf >> i) & 1u) run = 0;
                run = ((nf >> i) & 1u) ? 0 : run + 1;
#pragma unroll
                for (int j = W - 1; j > 0; --j) fwd[j] = (fwd[j] << 2) | (fwd[j - 1] >> 62);
                fwd[0] = (fwd[0] << 2) | b;
                fwd[W - 1] &= top_mask;
#pragma unroll
                for (int j = 0; j < W - 1; ++j) rev[j] = (rev[j] >> 2) | (rev[j + 1] << 62);
                rev[W - 1] = (rev[W - 1] >> 2) | ((u64)(3u - b) << top_shift);
                if (run >= k) {
                    const bool use_rev = key_less<W>(rev, fwd);
                    u64 key[W];
#pragma unroll
                    for (int j = 0; j < W; ++j) key[j] = use_rev ? rev[j] : fwd[j];
                    packed |= (u64)table.find(key, hash_key<W>(key)) << (16 * q);
                }
            }
            occ4[wi * 8 + g] = packed;
        }
    }
}

template <int W>
__global__ void __launch_bounds__(256)
insert_records_kernel(const u64 *__restrict__ rec, u64 n, int weighted, Table<W> table, Table<W> remote,
                      u32 n_shards, u32 rank, Counters *ctr, u64 *ovf, u64 ovf_cap)
{
    const int rw = weighted ? W + 1 : W;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    u32 newk = 0, newr = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        u64 key[W];
#pragma unroll
        for (int j = 0; j < W; ++j) key[j] = rec[i * rw + j];
        u64 wgt = weighted ? rec[i * rw + W] : 1ull;
        if (wgt == 0) continue;
        const u32 add = wgt > COUNT_SAT ? COUNT_SAT : (u32)wgt;
        const u64 h = hash_key<W>(key);
        int r;
        if (n_shards > 1 && shard_of_hash(h, n_shards) != rank) {
            r = remote.insert(key, h, add, false);
            newr += (r > 0);
        } else {
            r = table.insert(key, h, add, false);
            newk += (r > 0);
        }
        if (r < 0) {
            u64 at = atomicAdd(&ctr->overflow_n, 1ull);
            if (at < ovf_cap) {
#pragma unroll
                for (int j = 0; j < W; ++j) ovf[at * (W + 1) + j] = key[j];
                ovf[at * (W + 1) + W] = add;
            } else {
                atomicOr(&ctr->error_flags, ERR_OVERFLOW_LOST);
            }
        }
    }
    newk = warp_sum_u32(newk);
    newr = warp_sum_u32(newr);
    if ((threadIdx.x & 31) == 0) {
        if (newk) atomicAdd(&ctr->new_keys, (u64)newk);
        if (newr) atomicAdd(&ctr->new_keys_remote, (u64)newr);
    }
}

// Table from contigs (makeKmerReadDistributionFromContig, counter.h:511-593): every k-mer window of a contig gets
// max(its value so far, value of the contig), the contig's value = max(coverage, minOccurrence) supplied per stream position
// (val[p] for the window ENDING at p).  Same window enumeration as count_kernel.  Windows with an N are skipped (the
// reference does not skip them and feeds code 4 into the 2-bit fields; its contigs never contain N at this point).
template <int W>
__global__ void __launch_bounds__(256)
contig_max_kernel(const u64 *__restrict__ stream, const u32 *__restrict__ nflag, const u32 *__restrict__ rflag,
                  u64 word_begin, u64 word_end, int k, Table<W> table, const uint16_t *__restrict__ val, Counters *ctr)
{
    const int top_shift = 2 * ((k - 1) & 31);
    const u64 top_mask = (k & 31) ? ((1ull << (2 * (k & 31))) - 1ull) : ~0ull;
    const int s = 2 * (32 * W - k);
    const int nb = (k + 30) >> 5;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    u32 newk = 0;
    for (u64 wi = word_begin + (u64)blockIdx.x * blockDim.x + threadIdx.x; wi < word_end; wi += stride) {
        const u64 cur = stream[wi];
        const u32 nf = nflag[wi], rf = rflag[wi];
        int run = k;
        for (int j = 1; j <= nb; ++j) {
            const u32 a = nflag[wi - j], b = rflag[wi - j];
            if (a | b) {
                const int pn = a ? 32 - __clz(a) : 0;
                const int pr = b ? 31 - __clz(b) : 0;
                run = 32 * j - max(pn, pr);
                break;
            }
        }
        u64 fwd[W], rev[W];
#pragma unroll
        for (int j = 0; j < W; ++j) fwd[j] = stream[wi - 1 - j];
        {
            u64 y[W + 1];
#pragma unroll
            for (int j = 0; j < W; ++j) y[j] = pair_reverse64(~fwd[W - 1 - j]);
            y[W] = 0;
#pragma unroll
            for (int j = 0; j < W; ++j) rev[j] = s ? ((y[j] >> s) | (y[j + 1] << (64 - s))) : y[j];
        }
#pragma unroll 2
        for (int i = 0; i < 32; ++i) {
            const u32 b = (u32)(cur >> (62 - 2 * i)) & 3u;
            if ((rf >> i) & 1u) run = 0;
            run = ((nf >> i) & 1u) ? 0 : run + 1;
#pragma unroll
            for (int j = W - 1; j > 0; --j) fwd[j] = (fwd[j] << 2) | (fwd[j - 1] >> 62);
            fwd[0] = (fwd[0] << 2) | b;
            fwd[W - 1] &= top_mask;
#pragma unroll
            for (int j = 0; j < W - 1; ++j) rev[j] = (rev[j] >> 2) | (rev[j + 1] << 62);
            rev[W - 1] = (rev[W - 1] >> 2) | ((u64)(3u - b) << top_shift);
            if (run >= k) {
                const u32 v = val[wi * 32 + i];
                if (v == 0) continue;                               // a value of 0 is "no entry" (counter.h:490)
                const bool use_rev = key_less<W>(rev, fwd);
                u64 key[W];
#pragma unroll
                for (int j = 0; j < W; ++j) key[j] = use_rev ? rev[j] : fwd[j];
                const int r = table.insert_max(key, hash_key<W>(key), v);
                newk += (r > 0);
                if (r < 0) atomicAdd(&ctr->overflow_n, 1ull);       // the caller sized the table for every window: treated as an error
            }
        }
    }
    newk = warp_sum_u32(newk);
    if ((threadIdx.x & 31) == 0 && newk) atomicAdd(&ctr->new_keys, (u64)newk);
}

// The eight neighbour probes BruijnGraph::makeInitialBruijnGraph makes per k-mer of sortedKeyFP (graph.h:337-375; SURVEY.md
// section 8f row 4): for key i in its stored (canonical = forward) orientation, bit b of the high nibble says that the k-mer
// "b + first k-1 bases" is in the table, bit b of the low nibble that "last k-1 bases + b" is -- each looked up in canonical
// form with a count >= min_count (the table loadKmer would build holds only those).  out[i] = (leftFlags << 4) | rightFlags,
// the layout of Junction::out (graph.h:398).  One thread per key, eight independent read-only probes.
template <int W>
__global__ void __launch_bounds__(256)
neighbor_flags_kernel(const u64 *__restrict__ keys, u64 n, int k, Table<W> table, u32 min_count, uint8_t *__restrict__ out)
{
    const int top_w = (k - 1) >> 5, top_sh = 2 * ((k - 1) & 31);       // word and shift of key position k-1 (the leftmost base)
    const u64 top_mask = (k & 31) ? ((1ull << (2 * (k & 31))) - 1ull) : ~0ull;
    const int s = 2 * (32 * W - k);
    const u64 stride = (u64)gridDim.x * blockDim.x;
    if (min_count == 0) min_count = 1;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        u64 fwd[W], rev[W];
#pragma unroll
        for (int j = 0; j < W; ++j) fwd[j] = keys[i * W + j];
        {
            u64 y[W + 1];
#pragma unroll
            for (int j = 0; j < W; ++j) y[j] = pair_reverse64(~fwd[W - 1 - j]);
            y[W] = 0;
#pragma unroll
            for (int j = 0; j < W; ++j) rev[j] = s ? ((y[j] >> s) | (y[j + 1] << (64 - s))) : y[j];
        }
        // the k-1 shared bases, shifted into place once per side
        u64 lf[W], lr[W], rf[W], rr[W];
#pragma unroll
        for (int j = 0; j < W - 1; ++j) { lf[j] = (fwd[j] >> 2) | (fwd[j + 1] << 62); rr[j] = (rev[j] >> 2) | (rev[j + 1] << 62); }
        lf[W - 1] = fwd[W - 1] >> 2; rr[W - 1] = rev[W - 1] >> 2;
#pragma unroll
        for (int j = W - 1; j > 0; --j) { lr[j] = (rev[j] << 2) | (rev[j - 1] >> 62); rf[j] = (fwd[j] << 2) | (fwd[j - 1] >> 62); }
        lr[0] = rev[0] << 2; rf[0] = fwd[0] << 2;
        lr[W - 1] &= top_mask; rf[W - 1] &= top_mask;
        u32 left = 0, right = 0;
#pragma unroll
        for (u32 b = 0; b < 4; ++b) {
            u64 a[W], c[W];
            // left neighbour: b becomes the leftmost base of forward, its complement the rightmost of reverse
#pragma unroll
            for (int j = 0; j < W; ++j) { a[j] = lf[j]; c[j] = lr[j]; }
            a[top_w] |= (u64)b << top_sh; c[0] |= (u64)(3u - b);
            {
                const u64 *key = key_less<W>(c, a) ? c : a;
                if (table.find(key, hash_key<W>(key)) >= min_count) left |= 1u << b;
            }
            // right neighbour: b becomes the rightmost base of forward, its complement the leftmost of reverse
#pragma unroll
            for (int j = 0; j < W; ++j) { a[j] = rf[j]; c[j] = rr[j]; }
            a[0] |= (u64)b; c[top_w] |= (u64)(3u - b) << top_sh;
            {
                const u64 *key = key_less<W>(c, a) ? c : a;
                if (table.find(key, hash_key<W>(key)) >= min_count) right |= 1u << b;
            }
        }
        out[i] = (uint8_t)((left << 4) | right);
    }
}

// Seeded entries (makeKmerReadDistributionConsideringPreviousGraph, counter.h:663-750): a k-mer that was in the table
// before the reads were counted keeps its seeded value -- the reference never counts read windows that hit the table
// (divideKmerUsedMakingPreviousContig, counter.h:828-861) and dumps the table as it was (counter.h:695-705).  Counting
// everything and then SETTING the seeded keys gives the same table.  Records: W key words + value; value 0 = no entry.
template <int W>
__global__ void __launch_bounds__(256)
override_records_kernel(const u64 *__restrict__ rec, u64 n, Table<W> table, Counters *ctr, u64 *ovf, u64 ovf_cap)
{
    const u64 stride = (u64)gridDim.x * blockDim.x;
    u32 newk = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        u64 key[W];
#pragma unroll
        for (int j = 0; j < W; ++j) key[j] = rec[i * (W + 1) + j];
        const u64 wgt = rec[i * (W + 1) + W];
        if (wgt == 0) continue;
        const u32 w = wgt > COUNT_SAT ? COUNT_SAT : (u32)wgt;
        const u64 h = hash_key<W>(key);
        if (table.set_count(key, h, w)) continue;
        const int r = table.insert(key, h, w, false);              // no read contains it: a new entry
        newk += (r > 0);
        if (r < 0) {
            const u64 at = atomicAdd(&ctr->overflow_n, 1ull);
            if (at < ovf_cap) {
#pragma unroll
                for (int j = 0; j < W; ++j) ovf[at * (W + 1) + j] = key[j];
                ovf[at * (W + 1) + W] = w;
            } else {
                atomicOr(&ctr->error_flags, ERR_OVERFLOW_LOST);
            }
        }
    }
    newk = warp_sum_u32(newk);
    if ((threadIdx.x & 31) == 0 && newk) atomicAdd(&ctr->new_keys, (u64)newk);
}

// Counter::pickupReadMatchedEdgeKmer (counter.h:870-910): matched[r] = 1 iff some usable window of read r has its
// k-mer in the table.  occ = lookup_kernel's output (u16 per stream position, indexed by window END).
__global__ void __launch_bounds__(256)
read_match_kernel(const u64 *__restrict__ off, u64 n_reads, const uint16_t *__restrict__ occ, int k, uint8_t *matched)
{
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += stride) {
        const u64 b = off[r], e = off[r + 1];
        uint8_t m = 0;
        for (u64 p = b + (u64)k - 1; p < e; ++p)
            if (occ[p]) { m = 1; break; }
        matched[r] = m;
    }
}

// =================================================================================================
// partitioned counting: Pass A scatters canonical k-mers into P hash-range buckets, Pass B inserts one
// bucket at a time.  home slot = mulhi(hash, capacity) and bucket = mulhi(hash, P) are both monotone in
// the hash, so bucket b only touches the contiguous table region [b, b+1) * capacity / P, which is
// small enough to stay in the 126 MB L2 while the bucket is processed: DRAM sees two sequential
// streams (keys in, table region in/out) instead of one random sector per k-mer.
// =================================================================================================

#ifndef PBK_CPU_EMUL
#define PBK_DYN_SMEM(type, name) extern __shared__ __align__(16) type name[]
#else
#define PBK_DYN_SMEM(type, name) static type name[32768]
#endif


// a bucket-store entry (W == 1: the hash of the key) back to a key, then spilled
template <int W>
__device__ __forceinline__ void spill_stored(const u64 *stored, Counters *ctr, u64 *ovf, u64 ovf_cap)
{
    u64 key[W];
#pragma unroll
    for (int j = 0; j < W; ++j) key[j] = stored[j];
    if (W == 1) key[0] = fmix64_inverse(stored[0]);
    spill_key<W>(key, ctr, ovf, ovf_cap);
}

template <int W>
__device__ __forceinline__ void bucket_append_direct(const u64 *key, u32 b, u64 *bkt_keys, u64 seg_cap,
                                                     u64 *bkt_cursor, u64 *ovf, u64 ovf_cap, Counters *ctr)
{
    const u64 at = atomicAdd(&bkt_cursor[b], 1ull);
    if (at < seg_cap) {
#pragma unroll
        for (int j = 0; j < W; ++j) bkt_keys[((u64)b * seg_cap + at) * W + j] = key[j];
    } else {
        spill_stored<W>(key, ctr, ovf, ovf_cap);    // bucket segment full: goes through the overflow list
    }
}

// windows of one stream word a Pass A thread handles per tile
#ifndef PBK_PART_WIN3
#define PBK_PART_WIN3 8
#endif
template <int W> struct PART_WIN { static constexpr int value = W == 1 ? PART_WIN1 : (W == 2 ? 16 : PBK_PART_WIN3); };

// OR of x << j for j in [0, w), 0 <= w <= 32 (doubling: at most five shift/or steps)
__device__ __forceinline__ u64 smear_up(u64 x, int w)
{
    if (w <= 0) return 0;
    u64 r = x;
    int s = 1;
    while (2 * s <= w) { r |= r << s; s *= 2; }
    return r | (r << (w - s));
}

// Pass A.  One thread owns one stream word (32 window ends), exactly like count_kernel, but instead of
// touching the table it drops the canonical key into its bucket's shared-memory bin; full tiles are
// then flushed warp-per-bucket with coalesced stores.
// KEYX (key exchange between GPUs, pbk_keyx_*): the n_buckets = n_dest x R buckets are ordered by owner shard first
// (shard_of_hash, the low hash bits) and by table region second (the top hash bits, R a power of two), so that the
// segments of one destination are contiguous -- the bucket store is the all-to-all send buffer as it stands.
// PASSF (hash-range passes on one GPU, pbk_config.n_passes): n_dest carries (n_passes << 16) | pass_index and only the keys of
// that range are bucketed (and counted as instances); the other instantiations are unchanged.
template <int W, bool KEYX = false, bool PASSF = false>
__global__ void __launch_bounds__(512)
partition_kernel(const u64 *__restrict__ stream, const u32 *__restrict__ nflag, const u32 *__restrict__ rflag,
                 u64 word_begin, u64 word_end, int k, u32 n_buckets, u32 bin_cap, u64 *bkt_keys, u64 seg_cap,
                 u64 *bkt_cursor, Counters *ctr, u64 *ovf, u64 ovf_cap, u32 n_dest)
{
    PBK_DYN_SMEM(u64, bins);                          // n_buckets * bin_cap * W words
    __shared__ u32 scount[PART_MAX_BUCKETS];
    __shared__ u64 sbase[PART_MAX_BUCKETS];
    const int top_shift = 2 * ((k - 1) & 31);
    const u64 top_mask = (k & 31) ? ((1ull << (2 * (k & 31))) - 1ull) : ~0ull;
    const int s = 2 * (32 * W - k);
    const int nb = (k + 30) >> 5;
    // (wsize is 32 on the GPU; tests/cpu_emul runs the kernel as a single one-lane "warp")
    const int wsize = blockDim.x < 32 ? (int)blockDim.x : 32;
    const int lane = threadIdx.x % wsize;
    // bucket = top log2(P) hash bits when P is a power of two (0 = use the generic mulhi form)
    const u32 n_part = KEYX ? n_buckets / n_dest : n_buckets;     // buckets that the top hash bits select among
    const int pshift = (n_part > 1 && (n_part & (n_part - 1)) == 0) ? 64 - (31 - __clz(n_part)) : 0;
    u64 inst = 0;

    // A thread takes PART_WIN<W> consecutive windows of one stream word per tile (16, 16, 8 for 1, 2, >= 3 key words):
    // the shared-memory bins take 8 W bytes per staged instance, so a whole word per thread leaves room for only 16
    // (one-word keys) or 2-4 (multi-word keys) warps per SM.
    constexpr int WIN = PART_WIN<W>::value, SUBS = 32 / WIN;
    const u64 n_items = (word_end - word_begin) * SUBS;
    for (u64 tile = (u64)blockIdx.x * blockDim.x; tile < n_items; tile += (u64)gridDim.x * blockDim.x) {
        for (u32 b = threadIdx.x; b < n_buckets; b += blockDim.x) scount[b] = 0;
        __syncthreads();
        const u64 item = tile + threadIdx.x;
        const bool live = item < n_items;
        const u64 wi = word_begin + item / SUBS;
        const int i0 = (int)(item % SUBS) * WIN;
        if (W == 1 && pshift && live) {
            // One-word keys and a power-of-two bucket count: the hot configuration (k <= 32).  Fully unrolled so that every shift is an immediate;
            // which of the 32 windows are usable is worked out once per word with bit smears instead of a running
            // counter: the window ending at position i is usable iff none of its k positions is an N and none of
            // its last k-1 positions starts a read (flags of this and the previous word; bit j = position j).
            const u64 cur = stream[wi];
            const u64 nn = ((u64)nflag[wi] << 32) | nflag[wi - 1], rr = ((u64)rflag[wi] << 32) | rflag[wi - 1];
            const u32 valid = (~(u32)((smear_up(nn, k) | smear_up(rr, k - 1)) >> 32) >> i0) & (WIN < 32 ? (1u << (WIN & 31)) - 1u : ~0u);
            if (valid) {
                const u64 topmul = 1ull << top_shift;
                u32 n_here = 0;
                // rolling state = the 32 bases that end just before position i0 of this word
                u64 f = i0 ? ((stream[wi - 1] << (2 * i0)) | (cur >> (64 - 2 * i0))) : stream[wi - 1];
                u64 r = pair_reverse64(~f);
                r = s ? (r >> s) : r;
                u64 c = i0 ? cur << (2 * i0) : cur;
                // (interleaving four windows' hash chains by hand was measured: slower -- more registers, no gain)
#pragma unroll
                for (int i = 0; i < WIN; ++i) {
                    const u32 b = (u32)(c >> 62);
                    c <<= 2;
                    f = ((f << 2) | b) & top_mask;
                    r = (r >> 2) | ((u64)(3u - b) * topmul);
                    if ((valid >> i) & 1u) {
                        const u64 hh = fmix64(r < f ? r : f);        // key = min(forward, reverse), counter.h:429
                        if constexpr (PASSF) { if (shard_of_hash(hh, n_dest >> 16) != (n_dest & 0xFFFFu)) continue; }
                        u32 bkt = (u32)(hh >> pshift);                // == mulhi(hh, n_buckets) for a power of two
                        if constexpr (KEYX) bkt += shard_of_hash(hh, n_dest) * n_part;
                        ++n_here;
                        const u32 pos = atomicAdd(&scount[bkt], 1u);
                        if (pos < bin_cap) bins[bkt * bin_cap + pos] = hh;
                        else bucket_append_direct<W>(&hh, bkt, bkt_keys, seg_cap, bkt_cursor, ovf, ovf_cap, ctr);
                    }
                }
                inst += n_here;
            }
        } else if (live) {
            const u64 cur = stream[wi];
            const u32 nf = nflag[wi], rf = rflag[wi];
            // run = usable bases (same read, no N) ending just before window position i0 of this word
            int run = k;
            for (int j = 1; j <= nb; ++j) {
                const u32 a = nflag[wi - j], b = rflag[wi - j];
                if (a | b) {
                    const int pn = a ? 32 - __clz(a) : 0;
                    const int pr = b ? 31 - __clz(b) : 0;
                    run = 32 * j - max(pn, pr);
                    break;
                }
            }
            if (i0) {                                           // the positions of this word that other threads own
                const u32 a = nf & ((1u << i0) - 1u), b = rf & ((1u << i0) - 1u);
                if (a | b) {
                    const int pn = a ? 32 - __clz(a) : 0;
                    const int pr = b ? 31 - __clz(b) : 0;
                    run = i0 - max(pn, pr);
                } else {
                    run += i0;
                }
            }
            // Nothing to do if no window ending in [i0, i0 + WIN) can be usable: flags inside the range only shorten the run, so
            // run + WIN < k settles it.  With k = 75 on 150-base reads the first 74 positions of every read -- half of all
            // work items -- end here, before any key word is built.
            if (run + WIN >= k) {
            // rolling state = the 32 W bases that end just before position i0
            u64 fwd[W], rev[W];
            if (i0 == 0) {
#pragma unroll
                for (int j = 0; j < W; ++j) fwd[j] = stream[wi - 1 - j];
            } else {
                u64 hi = cur;
#pragma unroll
                for (int j = 0; j < W; ++j) {
                    const u64 lo = stream[wi - 1 - j];
                    fwd[j] = (lo << (2 * i0)) | (hi >> (64 - 2 * i0));
                    hi = lo;
                }
            }
            {
                u64 y[W + 1];
#pragma unroll
                for (int j = 0; j < W; ++j) y[j] = pair_reverse64(~fwd[W - 1 - j]);
                y[W] = 0;
#pragma unroll
                for (int j = 0; j < W; ++j) rev[j] = s ? ((y[j] >> s) | (y[j + 1] << (64 - s))) : y[j];
            }
            u64 c = i0 ? cur << (2 * i0) : cur;
            const u32 nfs = nf >> i0, rfs = rf >> i0;
#pragma unroll 2
            for (int i = 0; i < WIN; ++i) {
                const u32 b = (u32)(c >> 62);
                c <<= 2;
                if ((rfs >> i) & 1u) run = 0;
                run = ((nfs >> i) & 1u) ? 0 : run + 1;
#pragma unroll
                for (int j = W - 1; j > 0; --j) fwd[j] = (fwd[j] << 2) | (fwd[j - 1] >> 62);
                fwd[0] = (fwd[0] << 2) | b;
                fwd[W - 1] &= top_mask;
#pragma unroll
                for (int j = 0; j < W - 1; ++j) rev[j] = (rev[j] >> 2) | (rev[j + 1] << 62);
                rev[W - 1] = (rev[W - 1] >> 2) | ((u64)(3u - b) << top_shift);
                if (run >= k) {
                    const bool use_rev = key_less<W>(rev, fwd);
                    u64 key[W];
#pragma unroll
                    for (int j = 0; j < W; ++j) key[j] = use_rev ? rev[j] : fwd[j];
                    const u64 hh = hash_key<W>(key);
                    if constexpr (PASSF) { if (shard_of_hash(hh, n_dest >> 16) != (n_dest & 0xFFFFu)) continue; }
                    u32 bkt = (u32)__umul64hi(hh, (u64)n_part);
                    if constexpr (KEYX) bkt += shard_of_hash(hh, n_dest) * n_part;
                    if (W == 1) key[0] = hh;          // one-word keys travel as their (bijective) hash
                    ++inst;
                    const u32 pos = atomicAdd(&scount[bkt], 1u);
                    if (pos < bin_cap) {
                        u64 *dst = bins + (bkt * bin_cap + pos) * (u32)W;      // (bins hold < 2^15 words: 32-bit index arithmetic)
#pragma unroll
                        for (int j = 0; j < W; ++j) dst[j] = key[j];
                    } else {
                        bucket_append_direct<W>(key, bkt, bkt_keys, seg_cap, bkt_cursor, ovf, ovf_cap, ctr);
                    }
                }
            }
            }   // run + WIN >= k
        }
        __syncthreads();
        // flush, step 1: one thread per bucket reserves the bin's place in the bucket segment (all atomics of the
        // tile in flight together); step 2: a group of 16 lanes per bucket copies the bin with coalesced stores
        // (a bin holds ~27 keys: groups of 16 keep most lanes busy, a whole warp per bucket did not)
        for (u32 b = threadIdx.x; b < n_buckets; b += blockDim.x) {
            const u32 n = min(scount[b], bin_cap);
            scount[b] = n;
            sbase[b] = n ? atomicAdd(&bkt_cursor[b], (u64)n) : 0ull;
        }
        __syncthreads();
        {
            const u32 gs = blockDim.x < 16u ? blockDim.x : 16u;
            const u32 gl = threadIdx.x % gs, n_groups = blockDim.x / gs;
            for (u32 b = threadIdx.x / gs; b < n_buckets; b += n_groups) {
                const u32 n = scount[b];
                if (n == 0) continue;
                const u64 g = sbase[b];
                const u64 *src = bins + (u64)b * bin_cap * W;
                u64 *dst = bkt_keys + ((u64)b * seg_cap + g) * W;
                if (g + n <= seg_cap) {                              // the whole bin fits: plain coalesced copy
                    for (u32 i = gl; i < n * W; i += gs) st_stream_u64(dst + i, src[i]);
                } else {
                    for (u32 i = gl; i < n * W; i += gs) {
                        if (g + i / W < seg_cap) st_stream_u64(dst + i, src[i]);
                        else if (i % W == 0) spill_stored<W>(src + i, ctr, ovf, ovf_cap);   // bucket segment full
                    }
                }
            }
        }
        __syncthreads();
    }
    inst = warp_sum_u64(inst);
    if (lane == 0 && inst) atomicAdd(&ctr->instances, inst);
}

// Pass A for multi-word keys with the LIVE work items compacted across the CTA.  A work item (PART_WIN<W> consecutive window
// ends of one stream word) can only produce keys if the run of usable bases reaching into it is long enough (run + WIN >= k):
// with k = 75 on 150-base reads that excludes the first 74 positions of every read -- half of all items -- but a warp's 32
// consecutive items span 256 positions, so every warp holds dead and live items and the early exit of partition_kernel saves
// no issue slots.  Here the CTA first classifies candidate items (cheap: flag words and two clz) and queues the live ones
// (item index + run, 8 bytes) in shared memory; as soon as blockDim are queued, every thread takes one and does the expensive
// part -- rolling W-word forward / reverse keys, canonical minimum, hash, bin -- with all lanes busy; then the usual flush.
// Same bins, same bucket store, same results as partition_kernel<W>; only which thread handles which item differs.
constexpr int PARTC_MAX_THREADS = 512;
struct PartcEntry { u32 item; int run; };

template <int W, bool PASSF = false>
__global__ void __launch_bounds__(PARTC_MAX_THREADS)
partition_compact_kernel(const u64 *__restrict__ stream, const u32 *__restrict__ nflag, const u32 *__restrict__ rflag,
                         u64 word_begin, u64 word_end, int k, u32 n_buckets, u32 bin_cap, u64 *bkt_keys, u64 seg_cap,
                         u64 *bkt_cursor, Counters *ctr, u64 *ovf, u64 ovf_cap, u32 pass)     // pass: PASSF only, (n_passes << 16) | pass_index
{
    PBK_DYN_SMEM(u64, bins);                          // n_buckets * bin_cap * W words
    __shared__ u32 scount[PART_MAX_BUCKETS];
    __shared__ u64 sbase[PART_MAX_BUCKETS];
    __shared__ PartcEntry s_q[2 * PARTC_MAX_THREADS];
    __shared__ u32 s_qn;
    const int top_shift = 2 * ((k - 1) & 31);
    const u64 top_mask = (k & 31) ? ((1ull << (2 * (k & 31))) - 1ull) : ~0ull;
    const int s = 2 * (32 * W - k);
    const int nb = (k + 30) >> 5;
    const u32 nthreads = blockDim.x, tid = threadIdx.x;
    const u32 wsize = nthreads < 32u ? nthreads : 32u, lane = tid % wsize;
    constexpr int WIN = PART_WIN<W>::value, SUBS = 32 / WIN;
    const u64 n_items = (word_end - word_begin) * SUBS;           // < 2^32: a batch has at most 2^31 bases
    u64 inst = 0;
    u64 cursor = (u64)blockIdx.x * nthreads;                       // first candidate item of this CTA's next batch
    const u64 stride = (u64)gridDim.x * nthreads;
    if (tid == 0) s_qn = 0;
    __syncthreads();

    for (;;) {
        // ---- phase 1: classify candidates until a full round of live items is queued (or the input ends) ----------------
        while (s_qn < nthreads && cursor < n_items) {              // (uniform: s_qn is only changed between barriers)
            const u64 item = cursor + tid;
            int run = 0;
            bool live = item < n_items;
            if (live) {
                const u64 wi = word_begin + item / SUBS;
                const int i0 = (int)(item % SUBS) * WIN;
                run = k;
                for (int j = 1; j <= nb; ++j) {
                    const u32 a = nflag[wi - j], b = rflag[wi - j];
                    if (a | b) {
                        const int pn = a ? 32 - __clz(a) : 0;
                        const int pr = b ? 31 - __clz(b) : 0;
                        run = 32 * j - max(pn, pr);
                        break;
                    }
                }
                if (i0) {                                           // the positions of this word in front of the item
                    const u32 nf = nflag[wi], rf = rflag[wi];
                    const u32 a = nf & ((1u << i0) - 1u), b = rf & ((1u << i0) - 1u);
                    if (a | b) {
                        const int pn = a ? 32 - __clz(a) : 0;
                        const int pr = b ? 31 - __clz(b) : 0;
                        run = i0 - max(pn, pr);
                    } else {
                        run += i0;
                    }
                }
                live = run + WIN >= k;                              // flags inside the item only shorten the run
            }
            const unsigned bal = __ballot_sync(0xffffffffu, live);
            u32 base = 0;
            if (lane == 0 && bal) base = atomicAdd(&s_qn, (u32)__popc(bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (live) {
                PartcEntry e;
                e.item = (u32)item; e.run = run;
                s_q[base + (u32)__popc(bal & ((1u << lane) - 1u))] = e;
            }
            cursor += stride;
            __syncthreads();
        }
        const u32 qn = s_qn;
        if (qn == 0) break;
        const u32 take = qn < nthreads ? qn : nthreads;
        for (u32 b = tid; b < n_buckets; b += nthreads) scount[b] = 0;
        __syncthreads();

        // ---- phase 2: one live item per thread, all lanes of the active warps busy -------------------------------------------
        if (tid < take) {
            const PartcEntry e = s_q[qn - 1 - tid];
            const u64 wi = word_begin + e.item / SUBS;
            const int i0 = (int)(e.item % SUBS) * WIN;
            int run = e.run;
            const u64 cur = stream[wi];
            u64 fwd[W], rev[W];
            if (i0 == 0) {
#pragma unroll
                for (int j = 0; j < W; ++j) fwd[j] = stream[wi - 1 - j];
            } else {
                u64 hi = cur;
#pragma unroll
                for (int j = 0; j < W; ++j) {
                    const u64 lo = stream[wi - 1 - j];
                    fwd[j] = (lo << (2 * i0)) | (hi >> (64 - 2 * i0));
                    hi = lo;
                }
            }
            {
                u64 y[W + 1];
#pragma unroll
                for (int j = 0; j < W; ++j) y[j] = pair_reverse64(~fwd[W - 1 - j]);
                y[W] = 0;
#pragma unroll
                for (int j = 0; j < W; ++j) rev[j] = s ? ((y[j] >> s) | (y[j + 1] << (64 - s))) : y[j];
            }
            u64 c = i0 ? cur << (2 * i0) : cur;
            const u32 nfs = nflag[wi] >> i0, rfs = rflag[wi] >> i0;
#pragma unroll 2
            for (int i = 0; i < WIN; ++i) {
                const u32 b = (u32)(c >> 62);
                c <<= 2;
                if ((rfs >> i) & 1u) run = 0;
                run = ((nfs >> i) & 1u) ? 0 : run + 1;
#pragma unroll
                for (int j = W - 1; j > 0; --j) fwd[j] = (fwd[j] << 2) | (fwd[j - 1] >> 62);
                fwd[0] = (fwd[0] << 2) | b;
                fwd[W - 1] &= top_mask;
#pragma unroll
                for (int j = 0; j < W - 1; ++j) rev[j] = (rev[j] >> 2) | (rev[j + 1] << 62);
                rev[W - 1] = (rev[W - 1] >> 2) | ((u64)(3u - b) << top_shift);
                if (run >= k) {
                    const bool use_rev = key_less<W>(rev, fwd);
                    u64 key[W];
#pragma unroll
                    for (int j = 0; j < W; ++j) key[j] = use_rev ? rev[j] : fwd[j];
                    const u64 hh = hash_key<W>(key);
                    if constexpr (PASSF) { if (shard_of_hash(hh, pass >> 16) != (pass & 0xFFFFu)) continue; }
                    const u32 bkt = (u32)__umul64hi(hh, (u64)n_buckets);
                    ++inst;
                    const u32 pos = atomicAdd(&scount[bkt], 1u);
                    if (pos < bin_cap) {
                        u64 *dst = bins + (bkt * bin_cap + pos) * (u32)W;
#pragma unroll
                        for (int j = 0; j < W; ++j) dst[j] = key[j];
                    } else {
                        bucket_append_direct<W>(key, bkt, bkt_keys, seg_cap, bkt_cursor, ovf, ovf_cap, ctr);
                    }
                }
            }
        }
        __syncthreads();
        // ---- flush (as partition_kernel): reserve per bucket, then 16-lane groups copy the bins ------------------------------
        for (u32 b = tid; b < n_buckets; b += nthreads) {
            const u32 n = min(scount[b], bin_cap);
            scount[b] = n;
            sbase[b] = n ? atomicAdd(&bkt_cursor[b], (u64)n) : 0ull;
        }
        if (tid == 0) s_qn = qn - take;
        __syncthreads();
        {
            const u32 gs = nthreads < 16u ? nthreads : 16u;
            const u32 gl = tid % gs, n_groups = nthreads / gs;
            for (u32 b = tid / gs; b < n_buckets; b += n_groups) {
                const u32 n = scount[b];
                if (n == 0) continue;
                const u64 g = sbase[b];
                const u64 *src = bins + (u64)b * bin_cap * W;
                u64 *dst = bkt_keys + ((u64)b * seg_cap + g) * W;
                if (g + n <= seg_cap) {
                    for (u32 i = gl; i < n * W; i += gs) st_stream_u64(dst + i, src[i]);
                } else {
                    for (u32 i = gl; i < n * W; i += gs) {
                        if (g + i / W < seg_cap) st_stream_u64(dst + i, src[i]);
                        else if (i % W == 0) spill_stored<W>(src + i, ctr, ovf, ovf_cap);
                    }
                }
            }
        }
        __syncthreads();
    }
    inst = warp_sum_u64(inst);
    if (lane == 0 && inst) atomicAdd(&ctr->instances, inst);
}

// Pass B, ONE persistent launch for buckets [b_first, b_end).
//  * Order: tiles of PASSB_TILE_KEYS keys are handed out in bucket order through an atomic ticket, so
//    at any moment every CTA works inside the same one or two hash ranges and the table region they
//    touch stays in L2 (threads that are free to drift apart end up all over the table: 26 GB of HBM
//    reads instead of 4).
//  * Latency: the ticket and the keys of the NEXT tile are fetched while the current tile is inserted,
//    and a thread issues the first-probe atomics of all its 8 keys before looking at any result; the
//    few follow-up probes go in further rounds.  One atomic round trip is ~2.5 us under load, so
//    throughput is set by how many independent atomics are in flight.
//  * HBM: tile j of bucket b prefetches its share of the table region of a later bucket, which turns the
//    table's HBM traffic into one sequential pass.
constexpr int PASSB_THREADS = 256;
constexpr int PASSB_KPT = 8;
// threads per CTA as the HOST counts tiles (the kernels derive theirs from blockDim): tests/cpu_emul runs every kernel as
// a single thread
#ifndef PBK_CPU_EMUL
constexpr int PASSB_LAUNCH_THREADS = PASSB_THREADS;
#else
constexpr int PASSB_LAUNCH_THREADS = 1;
#endif
constexpr int PASSB_TILE_KEYS = PASSB_LAUNCH_THREADS * PASSB_KPT;
#ifndef PBK_PASSB_WIDE_CTAS
#define PBK_PASSB_WIDE_CTAS 4
#endif
constexpr int PASSB_WIDE_CTAS = PBK_PASSB_WIDE_CTAS;   // resident CTAs per SM of the k > 32 kernel: it is latency bound, occupancy is what it needs

struct PassBBucket {
    u64 tile_start;         // first ticket of this bucket (tiles of the launch are numbered in bucket order)
    u64 n_keys;             // keys in this bucket
    const char *pf_base;    // region to prefetch while this bucket is processed (main table), nullptr = none
    const char *pf_base2;   // same for the remote-staging table
    u32 pf_lines, pf_lines2;   // 128-byte lines in those regions
};

__device__ __forceinline__ void passb_prefetch(const char *base, u64 lines, u64 j, u64 nt, u32 tid, u32 nthreads)
{
#ifndef PBK_CPU_EMUL
    // this tile's slice of the region, one prefetch per 32-byte SECTOR: L2 fills sector by sector, a
    // prefetch per 128-byte line left three quarters of the region to be fetched by missing atomics
    const u64 s0 = 4 * lines * j / nt, s1 = 4 * lines * (j + 1) / nt;
    for (u64 l = s0 + tid; l < s1; l += nthreads) prefetch_keep(base + l * 32);
#endif
}

template <int W>
__global__ void __launch_bounds__(PASSB_THREADS, PASSB_WIDE_CTAS)
bucket_insert_kernel(const u64 *__restrict__ bkt_keys, u64 seg_cap, const PassBBucket *__restrict__ bk,
                     u32 b_first, u32 b_end, u64 *ticket, Table<W> table, Table<W> remote, u32 n_shards, u32 rank,
                     Counters *ctr, u64 *ovf, u64 ovf_cap)
{
    __shared__ PassBBucket s_bk[PART_MAX_BUCKETS + 1];
    __shared__ u64 s_ticket[2];
    const u32 nb = b_end - b_first, tid = threadIdx.x, nthreads = blockDim.x, tile_keys = blockDim.x * PASSB_KPT;
    for (u32 i = tid; i <= nb; i += nthreads) s_bk[i] = bk[i];
    if (tid == 0) { s_ticket[0] = atomicAdd(ticket, 1ull); s_ticket[1] = atomicAdd(ticket, 1ull); }
    __syncthreads();
    const u64 n_tiles = s_bk[nb].tile_start;
    u32 newk = 0, newr = 0;

    // software pipeline: `nxt` holds the keys of the tile this CTA processes next
    u64 nxt[PASSB_KPT][W];
    u32 lb = 0, lb_next = 0;                                     // bucket (relative to b_first) of the current / next tile
    u64 t = s_ticket[0];
    int par = 0;
    auto load_tile = [&](u64 tt, u32 &bb) {
        while (s_bk[bb + 1].tile_start <= tt) ++bb;
        const u64 base_i = (tt - s_bk[bb].tile_start) * tile_keys, n = s_bk[bb].n_keys;
        const u64 *src = bkt_keys + (u64)(b_first + bb) * seg_cap * W;
#pragma unroll
        for (int q = 0; q < PASSB_KPT; ++q) {
            const u64 i = base_i + (u64)q * nthreads + tid;
#pragma unroll
            for (int w = 0; w < W; ++w) nxt[q][w] = i < n ? ld_stream_u64(src + i * W + w) : 0;
        }
    };
    if (t < n_tiles) load_tile(t, lb_next);

    while (t < n_tiles) {
        lb = lb_next;
        u64 key[PASSB_KPT][W];
#pragma unroll
        for (int q = 0; q < PASSB_KPT; ++q)
#pragma unroll
            for (int w = 0; w < W; ++w) key[q][w] = nxt[q][w];
        const u64 j = t - s_bk[lb].tile_start, n = s_bk[lb].n_keys;
        const u64 nt = s_bk[lb + 1].tile_start - s_bk[lb].tile_start;
        const u64 base_i = j * tile_keys;

        const u64 t_next = s_ticket[par ^ 1];                    // fetched one iteration ago
        if (t_next < n_tiles) load_tile(t_next, lb_next);
        __syncthreads();                                         // everyone has read s_ticket[par]
        if (tid == 0) s_ticket[par] = atomicAdd(ticket, 1ull);   // ticket for the tile after next
        if (s_bk[lb].pf_base) passb_prefetch(s_bk[lb].pf_base, s_bk[lb].pf_lines, j, nt, tid, nthreads);
        if (s_bk[lb].pf_base2) passb_prefetch(s_bk[lb].pf_base2, s_bk[lb].pf_lines2, j, nt, tid, nthreads);

#pragma unroll 1
        for (int q = 0; q < PASSB_KPT; ++q) {
            if (base_i + (u64)q * nthreads + tid >= n) continue;
            const u64 h = hash_key<W>(key[q]);
            int r;
            if (n_shards > 1 && shard_of_hash(h, n_shards) != rank) { r = remote.insert(key[q], h, 1u, true); newr += (r > 0); }
            else { r = table.insert(key[q], h, 1u, true); newk += (r > 0); }
            if (r < 0) spill_key<W>(key[q], ctr, ovf, ovf_cap);
        }
        __syncthreads();                                         // s_ticket[par] written by thread 0
        t = t_next;
        par ^= 1;
    }
    newk = warp_sum_u32(newk);
    newr = warp_sum_u32(newr);
    if ((tid & 31) == 0) {
        if (newk) atomicAdd(&ctr->new_keys, (u64)newk);
        if (newr) atomicAdd(&ctr->new_keys_remote, (u64)newr);
    }
}

// Tile map of a Pass B launch built on the device from the bucket cursors, so that Pass A and Pass B of a
// sub-batch can be chained on the GPU without a host round trip (the host-built variant is
// launch_bucket_insert in pbk_kernels.cu).  One CTA.
__device__ __forceinline__ void passb_region(const char *slots, u64 cap, u32 slot_bytes, u32 pb, u32 n_buckets,
                                             const char **base, u32 *lines)
{
    *base = nullptr; *lines = 0;
    if (!slots || pb >= n_buckets) return;
    const u64 s0 = cap * pb / n_buckets;                       // cap < 2^40, pb < 2^9: no overflow
    u64 s1 = cap * (pb + 1) / n_buckets + 128;
    if (s1 > cap) s1 = cap;
    *lines = (u32)(((s1 - s0) * slot_bytes + 127) / 128);
    *base = slots + s0 * slot_bytes;
}

__global__ void __launch_bounds__(PART_MAX_BUCKETS)
passb_desc_kernel(const u64 *__restrict__ cursor, u64 seg_cap, u32 n_buckets, u32 tile_keys, const char *tab, u64 tab_cap,
                  const char *rtab, u64 rtab_cap, u32 slot_bytes, int pf_dist, u64 *ticket, PassBBucket *out)
{
    __shared__ u64 s_tiles[PART_MAX_BUCKETS + 1];
    for (u32 b = threadIdx.x; b < n_buckets; b += blockDim.x) {
        const u64 n = min(cursor[b], seg_cap);
        PassBBucket d;
        d.tile_start = 0; d.n_keys = n;
        d.pf_base = d.pf_base2 = nullptr; d.pf_lines = d.pf_lines2 = 0;
        if (pf_dist > 0 && n) {
            passb_region(tab, tab_cap, slot_bytes, b + pf_dist, n_buckets, &d.pf_base, &d.pf_lines);
            passb_region(rtab, rtab_cap, slot_bytes, b + pf_dist, n_buckets, &d.pf_base2, &d.pf_lines2);
        }
        out[b] = d;
        s_tiles[b] = (n + tile_keys - 1) / tile_keys;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 tiles = 0;
        for (u32 b = 0; b < n_buckets; ++b) { const u64 t = s_tiles[b]; s_tiles[b] = tiles; tiles += t; }
        s_tiles[n_buckets] = tiles;
        ticket[0] = 0; ticket[1] = 0;
    }
    __syncthreads();
    for (u32 b = threadIdx.x; b < n_buckets; b += blockDim.x) out[b].tile_start = s_tiles[b];
    if (threadIdx.x == 0) {
        PassBBucket e;
        e.tile_start = s_tiles[n_buckets]; e.n_keys = 0; e.pf_base = e.pf_base2 = nullptr; e.pf_lines = e.pf_lines2 = 0;
        out[n_buckets] = e;
    }
}

// Tile map for Pass B in gather mode (key exchange): descriptor i = (table region i / n_src, source i % n_src), its fill
// count is the cursor of that source for (this destination, region).  Built on the device so that a chunk can be
// inserted without a host round trip.
__global__ void __launch_bounds__(PART_MAX_BUCKETS)
passb_desc_gather_kernel(const __grid_constant__ KeyxSources srcs, u64 seg_cap, u32 n_src, u32 n_regions, u32 tile_keys, const char *tab,
                         u64 tab_cap, u32 slot_bytes, int pf_dist, u64 *ticket, PassBBucket *out)
{
    __shared__ u64 s_tiles[PART_MAX_BUCKETS + 1];
    const u32 n_desc = n_src * n_regions;
    for (u32 i = threadIdx.x; i < n_desc; i += blockDim.x) {
        const u64 n = min(ld_cg_u64(srcs.cursors[i % n_src] + i / n_src), seg_cap);
        PassBBucket d;
        d.tile_start = 0; d.n_keys = n;
        d.pf_base = d.pf_base2 = nullptr; d.pf_lines = d.pf_lines2 = 0;
        if (pf_dist > 0 && n) passb_region(tab, tab_cap, slot_bytes, i + (u32)pf_dist * n_src, n_desc, &d.pf_base, &d.pf_lines);
        out[i] = d;
        s_tiles[i] = (n + tile_keys - 1) / tile_keys;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 tiles = 0;
        for (u32 i = 0; i < n_desc; ++i) { const u64 t = s_tiles[i]; s_tiles[i] = tiles; tiles += t; }
        s_tiles[n_desc] = tiles;
        ticket[0] = 0; ticket[1] = 0;
    }
    __syncthreads();
    for (u32 i = threadIdx.x; i < n_desc; i += blockDim.x) out[i].tile_start = s_tiles[i];
    if (threadIdx.x == 0) {
        PassBBucket e;
        e.tile_start = s_tiles[n_desc]; e.n_keys = 0; e.pf_base = e.pf_base2 = nullptr; e.pf_lines = e.pf_lines2 = 0;
        out[n_desc] = e;
    }
}

// Pass B for one-word keys (k <= 32).  The bucket store holds h = fmix64(key), so a key costs one
// streamed 8-byte load, one 64-bit atomic add on its home slot and a three-instruction test of the
// returned word.  What does not finish there -- the home slot belongs to another key, or it is claimed
// but its tag is not visible yet -- is NOT resolved on the spot: a dependent chain of L2 round trips
// executed by one or two lanes stalls the whole warp (measured: the 1 % of such keys tripled the time
// of the pass).  Instead the key goes on a small per-warp list in shared memory and rides along with
// the next round's batch of atomics (as probe d+1, or as a plain re-read of the slot).  Every round
// therefore is: loads out, all atomics out, one look at the results, no waiting on other threads.
constexpr int PASSB1_KPT = 8;
constexpr int PASSB1_ROUNDS = 4;
constexpr int PASSB1_TILE_KEYS = PASSB_LAUNCH_THREADS * PASSB1_KPT * PASSB1_ROUNDS;
constexpr int PASSB1_DEF_CAP = 32 * (PASSB1_KPT + 2);   // deferred keys per warp: what one round can add, plus one batch
constexpr u32 PASSB1_DONE = 0xFFFFu;
constexpr u32 PASSB1_VERIFY = 1u << 8, PASSB1_REMOTE = 1u << 9;   // deferred-entry flags above the displacement

#ifdef PBK_CPU_EMUL
static inline void __syncwarp() {}
#endif

// Look at what probe `d` of key `h` returned.  `verify` = the +1 is already in the slot and `old` is a
// fresh read of it.  Returns PASSB1_DONE or the list entry (displacement + flags) of the follow-up.
__device__ __forceinline__ u32 passb1_classify(const Table<1> &tb, bool is_remote, u64 h, u32 d, bool verify, u64 old,
                                               u64 keep, u32 &newk, u32 &newr, Counters *ctr, u64 *ovf, u64 ovf_cap)
{
    const u64 tag = (h << (64 - tb.g.rbits)) | ((u64)(d + 1) << tb.g.cbits);
    const u64 x = old ^ tag;
    if (x < tb.g.dz) return PASSB1_DONE;                          // our tag and room in the count field: the common case
    u64 *s = tb.slots + (((h >> tb.g.rbits) + d) & tb.g.capmask);
    if (old == 0) {                                               // we saw the slot empty: it is ours, publish the tag
        if (keep) red_add_keep_u64(s, tag, keep); else red_add_u64(s, tag);
        if (is_remote) ++newr; else ++newk;
        return PASSB1_DONE;
    }
    if ((old >> tb.g.cbits) == 0)                                 // claimed, tag not visible yet: look again next round
        return d | PASSB1_VERIFY | (is_remote ? PASSB1_REMOTE : 0u);
    if (keep) red_add_keep_u64(s, ~0ull, keep); else red_add_u64(s, ~0ull);   // not ours, or ours but saturated for sure
    if ((x >> tb.g.cbits) == 0) return PASSB1_DONE;
    if (d >= (u32)CT_MAX_DISP) { u64 k0 = fmix64_inverse(h); spill_key<1>(&k0, ctr, ovf, ovf_cap); return PASSB1_DONE; }
    return (d + 1) | (is_remote ? PASSB1_REMOTE : 0u);
}

// `src` / `n_valid`: this WARP's keys of the round (KPT x warp-size consecutive entries of the bucket store, lane-strided);
// STAGED: they were brought into shared memory by a TMA bulk copy and `src` points there.
template <bool SHARDED, bool FULL, bool STAGED = false>
__device__ __forceinline__ void passb1_round(const u64 *__restrict__ src, u32 n_valid, u32 tid, u32 nthreads, u32 lane,
                                             const Table<1> &table, const Table<1> &remote, u32 n_shards, u32 rank,
                                             u64 keep, u32 &newk, u32 &newr, Counters *ctr, u64 *ovf, u64 ovf_cap,
                                             u64 *def_h, uint16_t *def_m, u32 &n_def, u32 opts)
{
    // deferred keys of earlier rounds: the top min(n_def, warp size) entries, one per lane
    const u32 wsize = nthreads < 32u ? nthreads : 32u;
    const u32 take = n_def < wsize ? n_def : wsize;
    u64 xh = 0;
    u32 xm = PASSB1_DONE;
    __syncwarp();
    if (lane < take) { xh = def_h[n_def - 1 - lane]; xm = def_m[n_def - 1 - lane]; }
    __syncwarp();
    n_def -= take;

    u64 h[PASSB1_KPT], old[PASSB1_KPT];
    u32 valid = 0, rem = 0;
    if constexpr (FULL && !STAGED) {
        // a full warp slice: four 16-byte loads per lane (512 contiguous bytes per warp request) instead of eight 8-byte ones --
        // half as many requests for the same keys, which is what counts when the bucket store is a peer's, behind NVLink
        // (the key exchange's pull form); which lane inserts which key is irrelevant.  Slices start 16-byte aligned (even seg_cap).
        static_assert(PASSB1_KPT % 2 == 0, "pairs of keys");
#pragma unroll
        for (int p = 0; p < PASSB1_KPT / 2; ++p) {
            const ulonglong2 v = ld_stream_u64x2(src + 2 * ((u32)p * wsize + lane));
            h[2 * p] = v.x; h[2 * p + 1] = v.y;
        }
        valid = (1u << PASSB1_KPT) - 1u;
    } else {
#pragma unroll
        for (int q = 0; q < PASSB1_KPT; ++q) {
            const u32 i = (u32)q * wsize + lane;
            if (FULL || i < n_valid) { h[q] = STAGED ? src[i] : ld_stream_u64(src + i); valid |= 1u << q; }
            else h[q] = 0;
        }
    }
#ifdef PBK_EXPERIMENT
    if (opts & 2u) {                     // experiment: no table traffic at all
        u64 acc = 0;
#pragma unroll
        for (int q = 0; q < PASSB1_KPT; ++q) acc ^= h[q];
        if (acc == 0x1234567ull) ++newk;
        return;
    }
#endif
#pragma unroll
    for (int q = 0; q < PASSB1_KPT; ++q) {
        if (SHARDED && shard_of_hash(h[q], n_shards) != rank) rem |= 1u << q;
        const Table<1> &tb = (SHARDED && (rem >> q & 1)) ? remote : table;
        if (FULL || (valid >> q & 1)) old[q] = keep ? atom_add_keep_u64(tb.slots + (h[q] >> tb.g.rbits), 1ull, keep)
                                                    : atomicAdd(tb.slots + (h[q] >> tb.g.rbits), 1ull);
        else old[q] = 0;
    }
    u64 xold = 0;
    const bool x_remote = SHARDED && (xm & PASSB1_REMOTE), x_verify = (xm & PASSB1_VERIFY) != 0;
    const u32 xd = xm & 0xFFu;
    if (xm != PASSB1_DONE) {
        const Table<1> &tb = x_remote ? remote : table;
        u64 *s = tb.slots + (((xh >> tb.g.rbits) + xd) & tb.g.capmask);
        xold = x_verify ? ld_cg_u64(s) : (keep ? atom_add_keep_u64(s, 1ull, keep) : atomicAdd(s, 1ull));
    }
#ifdef PBK_EXPERIMENT
    if (opts & 8u) {                     // experiment: atomics only, results ignored
        u64 acc = 0;
#pragma unroll
        for (int q = 0; q < PASSB1_KPT; ++q) acc ^= old[q];
        if (acc == 0x1234567ull) ++newk;
        return;
    }
#endif
    // one look at every result; follow-ups go on the list
#pragma unroll
    for (int q = 0; q <= PASSB1_KPT; ++q) {
        u32 m = PASSB1_DONE;
        u64 hq;
        if (q < PASSB1_KPT) {
            hq = h[q < PASSB1_KPT ? q : 0];
            if (FULL || (valid >> q & 1)) {
                const bool r = SHARDED && (rem >> q & 1);
                m = passb1_classify(r ? remote : table, r, hq, 0, false, old[q < PASSB1_KPT ? q : 0], keep, newk, newr, ctr, ovf, ovf_cap);
            }
        } else {
            hq = xh;
            if (xm != PASSB1_DONE) m = passb1_classify(x_remote ? remote : table, x_remote, hq, xd, x_verify, xold, keep, newk, newr, ctr, ovf, ovf_cap);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, m != PASSB1_DONE);
        if (bal == 0) continue;
        if (m != PASSB1_DONE) {
            const u32 pos = n_def + (u32)__popc(bal & ((1u << lane) - 1u));
            def_h[pos] = hq;                                     // pos < PASSB1_DEF_CAP: callers keep n_def <= one batch
            def_m[pos] = (uint16_t)m;                            // before a round that brings new keys
        }
        n_def += (u32)__popc(bal);
    }
}

// MODE 0: every key is local.  MODE 1: keys of other shards go to the remote-staging table (record exchange after
// counting).  MODE 2 (key exchange before counting, pbk_keyx_insert_device): every key is local and `bkt_hash` is the
// all-to-all receive buffer, laid out [source rank][region][seg_cap]; the descriptors are in region-major order
// (descriptor i = region i / G, source i % G, with G = opts bits 8-15 and the regions per source in bits 16-31), so
// all sources' keys of one table region are inserted while that region is L2-resident.
// STAGED: a warp's keys of a round (KPT x 32 entries = 2 KB, contiguous in the bucket store -- local HBM, or a peer's over
// NVLink in the key exchange) arrive through a TMA bulk copy (cp.async.bulk -> shared memory, completion on the warp's own
// mbarrier) that lane 0 issues ONE ROUND AHEAD, into the other half of the warp's two-stage ring: the threads never have the
// key loads in flight themselves, and a remote store is read in 2 KB bursts instead of 256-byte warp requests.  Warps stay
// independent (no CTA barrier per round).
constexpr int PASSB1_WARP_KEYS = 32 * PASSB1_KPT;
constexpr size_t PASSB1_RING_BYTES = (size_t)(PASSB_THREADS / 32) * 2 * PASSB1_WARP_KEYS * 8;     // 32 KB per CTA
template <int MODE, bool STAGED>
__device__ __forceinline__ void
bucket_insert_compact_body(const u64 *__restrict__ bkt_hash, u64 seg_cap, const PassBBucket *__restrict__ bk,
                           u32 b_first, u32 b_end, u64 *ticket, const Table<1> &table, const Table<1> &remote, u32 n_shards,
                           u32 rank, Counters *ctr, u64 *ovf, u64 ovf_cap, u32 opts, const KeyxSources *srcs)
{
    constexpr bool SHARDED = MODE == 1;
    // (the bucket descriptors are read straight from global memory: a few cached loads per 8192-key tile, and the
    //  shared memory they would take is what lets two of these CTAs share an SM with two Pass A CTAs)
    __shared__ u64 s_ticket[2];
    __shared__ u64 s_def_h[PASSB_THREADS / 32][PASSB1_DEF_CAP];
    __shared__ uint16_t s_def_m[PASSB_THREADS / 32][PASSB1_DEF_CAP];
    PBK_DYN_SMEM(u64, s_ring);                       // STAGED: [warp][half][PASSB1_WARP_KEYS], PASSB1_RING_BYTES of dynamic shared memory
    __shared__ u64 s_wbar[PASSB_THREADS / 32][2];
    const u32 nb = b_end - b_first, tid = threadIdx.x, nthreads = blockDim.x;
    const u32 lane = nthreads < 32u ? 0u : (tid & 31u), warp = nthreads < 32u ? 0u : (tid >> 5);
    const u32 wsize = nthreads < 32u ? nthreads : 32u, warp_keys = wsize * PASSB1_KPT;
    u64 *def_h = s_def_h[warp];
    uint16_t *def_m = s_def_m[warp];
    u32 n_def = 0;
    const u32 def_room = nthreads < 32u ? 1u : 32u;              // a round may only start with at most one batch listed
    const u32 round_keys = nthreads * PASSB1_KPT, tile_keys = round_keys * PASSB1_ROUNDS;
    if (tid == 0) { s_ticket[0] = atomicAdd(ticket, 1ull); s_ticket[1] = atomicAdd(ticket, 1ull); }
    __syncthreads();
    const u64 n_tiles = bk[nb].tile_start;
    const u64 keep = (opts & 1u) ? l2_keep_policy() : 0ull;      // 0 = plain atomics (a real policy word is never 0)
    u32 newk = 0, newr = 0, lb = 0;
    int par = 0;
    u64 t = s_ticket[0];
    // where tile `tt` starts in the (local or peer) bucket store, and how many keys it has; `bb` = monotone bucket cursor
    auto tile_src = [&](u64 tt, u32 &bb, u64 &n_in_tile) -> const u64 * {
        while (bk[bb + 1].tile_start <= tt) ++bb;
        const u64 jj = tt - bk[bb].tile_start;
        n_in_tile = min((u64)tile_keys, bk[bb].n_keys - jj * tile_keys);
        const u32 seg = b_first + bb;
        if constexpr (MODE == 2) {                               // descriptor = (region seg / G, source seg % G)
            const u32 G = (opts >> 8) & 0xFFu;
            return srcs->keys[seg % G] + (u64)(seg / G) * seg_cap + jj * tile_keys;
        } else {
            return bkt_hash + (u64)seg * seg_cap + jj * tile_keys;
        }
    };
    // STAGED: lane 0 starts the bulk copy of this warp's keys of round r of the tile at `base` into ring half `half`
    int ring = 0;
    u32 ring_phase = 0;                                       // bit h = parity the next wait on ring half h expects
    auto issue = [&](const u64 *base, u64 n_in_tile, u32 r, int half) {
        const u64 off = (u64)r * round_keys + (u64)warp * warp_keys;
        if (off >= n_in_tile) return;                            // nothing for this warp in that round: no copy, no wait
        const u32 cnt = (u32)min((u64)warp_keys, n_in_tile - off);
        const u32 bytes = (cnt * 8u + 15u) & ~15u;
        if (lane == 0) {
            mbar_expect_tx(&s_wbar[warp][half], bytes);
            bulk_copy_g2s(s_ring + (size_t)(warp * 2 + half) * PASSB1_WARP_KEYS, base + off, bytes, &s_wbar[warp][half]);
        }
    };
    u32 lb_next = 0;
    if constexpr (STAGED) {
        if (lane == 0) { mbar_init(&s_wbar[warp][0], 1); mbar_init(&s_wbar[warp][1], 1); mbar_fence_init(); }
        __syncwarp();
        if (t < n_tiles) { u64 n0; const u64 *b0 = tile_src(t, lb_next, n0); issue(b0, n0, 0, 0); }
    }
    while (t < n_tiles) {
        u64 n_tile;
        const u64 *src = tile_src(t, lb, n_tile);
        const u64 j = t - bk[lb].tile_start;
        const u64 nt = bk[lb + 1].tile_start - bk[lb].tile_start;
        const u64 t_next = s_ticket[par ^ 1];                    // fetched one iteration ago
        __syncthreads();                                         // everyone has read both tickets
        if (tid == 0) s_ticket[par] = atomicAdd(ticket, 1ull);   // ticket for the tile after next
        if (bk[lb].pf_base) passb_prefetch(bk[lb].pf_base, bk[lb].pf_lines, j, nt, tid, nthreads);
        if (SHARDED && bk[lb].pf_base2) passb_prefetch(bk[lb].pf_base2, bk[lb].pf_lines2, j, nt, tid, nthreads);
        const u64 *src_next = nullptr;
        u64 n_next = 0;
        if (STAGED && t_next < n_tiles) src_next = tile_src(t_next, lb_next, n_next);
        const u32 n_rounds = (u32)((n_tile + round_keys - 1) / round_keys);
#pragma unroll 1
        for (u32 r = 0; r < n_rounds; ++r) {
            if constexpr (STAGED) {                              // the warp's next round -- of this tile or of the next -- goes out first
                if (r + 1 < n_rounds) issue(src, n_tile, r + 1, ring ^ 1);
                else if (src_next) issue(src_next, n_next, 0, ring ^ 1);
            }
#pragma unroll 1
            while (n_def > def_room)
                passb1_round<SHARDED, false>(bkt_hash, 0u, tid, nthreads, lane, table, remote, n_shards, rank, keep, newk, newr,
                                             ctr, ovf, ovf_cap, def_h, def_m, n_def, opts);
            const u64 off = (u64)r * round_keys + (u64)warp * warp_keys;
            const u32 cnt = off < n_tile ? (u32)min((u64)warp_keys, n_tile - off) : 0u;
            const u64 *wsrc = src + off;
            if constexpr (STAGED) {
                if (cnt) { mbar_wait(&s_wbar[warp][ring], (ring_phase >> ring) & 1u); ring_phase ^= 1u << ring; }
                wsrc = s_ring + (size_t)(warp * 2 + ring) * PASSB1_WARP_KEYS;
            }
            if (cnt == warp_keys)
                passb1_round<SHARDED, true, STAGED>(wsrc, cnt, tid, nthreads, lane, table, remote, n_shards, rank, keep, newk, newr,
                                                    ctr, ovf, ovf_cap, def_h, def_m, n_def, opts);
            else if (cnt)
                passb1_round<SHARDED, false, STAGED>(wsrc, cnt, tid, nthreads, lane, table, remote, n_shards, rank, keep, newk, newr,
                                                     ctr, ovf, ovf_cap, def_h, def_m, n_def, opts);
            if constexpr (STAGED) { __syncwarp(); ring ^= 1; }   // every lane has its keys in registers before the half is refilled
        }
        __syncthreads();                                         // s_ticket[par] written by thread 0
        t = t_next;
        par ^= 1;
    }
    // drain the warp's list (n_def is warp-uniform); publishing never waits, so this terminates
#pragma unroll 1
    while (n_def)
        passb1_round<SHARDED, false>(bkt_hash, 0u, tid, nthreads, lane, table, remote, n_shards, rank, keep, newk, newr, ctr, ovf,
                                     ovf_cap, def_h, def_m, n_def, opts);
    newk = warp_sum_u32(newk);
    newr = warp_sum_u32(newr);
    if ((tid & 31) == 0) {
        if (newk) atomicAdd(&ctr->new_keys, (u64)newk);
        if (newr) atomicAdd(&ctr->new_keys_remote, (u64)newr);
    }
}

template <int MODE>
__global__ void __launch_bounds__(PASSB_THREADS, 3)
bucket_insert_compact_kernel(const u64 *__restrict__ bkt_hash, u64 seg_cap, const PassBBucket *__restrict__ bk,
                             u32 b_first, u32 b_end, u64 *ticket, Table<1> table, Table<1> remote, u32 n_shards,
                             u32 rank, Counters *ctr, u64 *ovf, u64 ovf_cap, u32 opts)
{
    static_assert(MODE == 0 || MODE == 1, "MODE 2 (key exchange) is bucket_insert_gather_kernel");
    bucket_insert_compact_body<MODE, false>(bkt_hash, seg_cap, bk, b_first, b_end, ticket, table, remote, n_shards, rank, ctr, ovf, ovf_cap,
                                            opts, nullptr);
}
// the same with the keys staged through per-warp TMA bulk copies (see bucket_insert_compact_body)
template <int MODE>
__global__ void __launch_bounds__(PASSB_THREADS, 3)
bucket_insert_compact_staged_kernel(const u64 *__restrict__ bkt_hash, u64 seg_cap, const PassBBucket *__restrict__ bk,
                                    u32 b_first, u32 b_end, u64 *ticket, Table<1> table, Table<1> remote, u32 n_shards,
                                    u32 rank, Counters *ctr, u64 *ovf, u64 ovf_cap, u32 opts)
{
    bucket_insert_compact_body<MODE, true>(bkt_hash, seg_cap, bk, b_first, b_end, ticket, table, remote, n_shards, rank, ctr, ovf, ovf_cap,
                                           opts, nullptr);
}

// MODE 2 as a kernel: Pass B over the keys every source rank holds for this shard (KeyxSources: receive buffer or peer HBM)
__global__ void __launch_bounds__(PASSB_THREADS, 3)
bucket_insert_gather_kernel(const __grid_constant__ KeyxSources srcs, u64 seg_cap, const PassBBucket *__restrict__ bk,
                            u32 b_first, u32 b_end, u64 *ticket, Table<1> table, Counters *ctr, u64 *ovf, u64 ovf_cap, u32 opts)
{
    bucket_insert_compact_body<2, false>(nullptr, seg_cap, bk, b_first, b_end, ticket, table, table, 1u, 0u, ctr, ovf, ovf_cap, opts, &srcs);
}
__global__ void __launch_bounds__(PASSB_THREADS, 3)
bucket_insert_gather_staged_kernel(const __grid_constant__ KeyxSources srcs, u64 seg_cap, const PassBBucket *__restrict__ bk,
                                   u32 b_first, u32 b_end, u64 *ticket, Table<1> table, Counters *ctr, u64 *ovf, u64 ovf_cap, u32 opts)
{
    bucket_insert_compact_body<2, true>(nullptr, seg_cap, bk, b_first, b_end, ticket, table, table, 1u, 0u, ctr, ovf, ovf_cap, opts, &srcs);
}

// =================================================================================================
// Pass B, second form for one-word keys on an unsharded table: split + build (PBK_PASSB2).
//
// bucket_insert_compact_kernel above sits at the rate of the L2's 64-bit atomics (one per k-mer instance, ~100 G/s).  The only
// way under that is not to send the increments to the L2 at all: here a CTA owns a SUB-REGION of the table -- BUILD_SLOTS
// consecutive slots, 64 KB -- outright, holds it in shared memory while every key whose home slot lies in it is inserted
// with shared-memory atomics, and writes it back with coalesced 16-byte stores.  Nothing is shared between CTAs, so there is
// no claim/publish protocol across the chip: a slot word changes from 0 to (tag | 1) in ONE 64-bit shared-memory
// compare-and-swap, and a hit is a 32-bit shared-memory add on its low half (the count field).
// That needs the keys sorted by sub-region.  Pass A's P buckets (16 MB table regions) stay as they are; split_kernel takes a
// bucket's keys tile by tile and scatters them into the F = region / sub-region segments of that bucket (second-level
// partition on the next log2 F hash bits: counting sort of the tile in shared memory, one reservation per segment and tile,
// coalesced copy-out), region_build_kernel then drains segment after segment.  DRAM sees the keys twice more (8 B written,
// 8 B read per instance) -- sequential traffic the HBM has room for -- and the table once, written only when the table was
// empty before (pbk_reset): `load_existing` = 0.
// A probe sequence that would leave the sub-region (or pass CT_MAX_DISP) ends on the overflow list like every other key that
// finds no slot; drain_overflow inserts those with the global protocol afterwards, which may place them in the next
// sub-region's first slots -- to later launches just occupied slots with a foreign tag.
// =================================================================================================
constexpr int SPLIT_THREADS = 512;
constexpr int SPLIT_KPT = 16;
constexpr int SPLIT_MAX_F = 512;                     // sub-regions per bucket
#ifndef PBK_CPU_EMUL
constexpr int SPLIT_LAUNCH_THREADS = SPLIT_THREADS;
#else
constexpr int SPLIT_LAUNCH_THREADS = 1;
#endif
constexpr int SPLIT_TILE_KEYS = SPLIT_LAUNCH_THREADS * SPLIT_KPT;
constexpr int BUILD_THREADS = 512;
constexpr int BUILD_LOG2_SLOTS = 13;
constexpr u32 BUILD_SLOTS = 1u << BUILD_LOG2_SLOTS;  // 64 KB of compact slots
constexpr int BUILD_KPT = 4;                         // new keys per lane and round (two 16-byte loads)
constexpr int BUILD_DEF_CAP = 32 * (BUILD_KPT + 2);  // list entries per warp: what one round can add, plus one batch
constexpr size_t BUILD_SMEM_BYTES = ((size_t)BUILD_SLOTS + (size_t)(BUILD_THREADS / 32) * BUILD_DEF_CAP) * 8;    // 88 KB: two CTAs per SM

// bk: tile map (tiles of blockDim * SPLIT_KPT keys) of ALL descriptors, built by passb_desc_kernel / passb_desc_gather_kernel; this
// launch splits descriptors [d_first, d_end).  Local form: descriptor = bucket, keys in bkt_hash.  GATHER (key exchange: the keys
// every source rank holds for this shard, read in place -- out of the receive buffer, or out of the peers' HBM over NVLink in
// the pull form): descriptor i = (table region i / n_src, source i % n_src), and "bucket" below is the table region.
// sub_shift = rbits + BUILD_LOG2_SLOTS: h >> sub_shift is the global sub-region of hash h, its low log2 F bits the
// sub-region within the bucket.  sub_keys: [global sub-region][sub_cap] hashes, sub_cursor: their fill counts (zeroed).
template <bool GATHER>
__device__ __forceinline__ void
split_body(const u64 *__restrict__ bkt_hash, const KeyxSources *srcs, u32 n_src, u64 seg_cap, const PassBBucket *__restrict__ bk,
           u32 d_first, u32 d_end, u32 F, int sub_shift, u64 *__restrict__ sub_keys, u64 sub_cap, u64 *sub_cursor, Counters *ctr,
           u64 *ovf, u64 ovf_cap)
{
    PBK_DYN_SMEM(u64, s_sorted);                     // the tile's keys in sub-region order: blockDim * SPLIT_KPT entries
    __shared__ u32 s_cnt[SPLIT_MAX_F], s_off[SPLIT_MAX_F], s_lim[SPLIT_MAX_F];
    __shared__ u64 s_dst[SPLIT_MAX_F];
    const u32 tid = threadIdx.x, nthreads = blockDim.x, tile_keys = nthreads * SPLIT_KPT;
    const u32 wsize = nthreads < 32u ? nthreads : 32u;
    const u64 t_end = bk[d_end].tile_start;
    u32 lb = d_first;
    for (u64 t = bk[d_first].tile_start + blockIdx.x; t < t_end; t += gridDim.x) {
        while (bk[lb + 1].tile_start <= t) ++lb;
        const u64 j = t - bk[lb].tile_start;
        const u32 n = (u32)min((u64)tile_keys, bk[lb].n_keys - j * tile_keys);
        const u32 bucket = GATHER ? lb / n_src : lb;
        const u64 *src = (GATHER ? srcs->keys[lb % n_src] + (u64)bucket * seg_cap : bkt_hash + (u64)lb * seg_cap) + j * tile_keys;
        for (u32 f = tid; f < F; f += nthreads) s_cnt[f] = 0;
        __syncthreads();
        // 1. keys into registers, each counted into its sub-region (the count's old value = the key's rank there)
        u64 h[SPLIT_KPT];
        u32 sp[SPLIT_KPT];                           // sub-region << 16 | rank (rank < tile_keys <= 8192)
        if (n == tile_keys) {                        // full tile: 16-byte loads (tiles start 16-byte aligned: even seg_cap)
#pragma unroll
            for (int p = 0; p < SPLIT_KPT / 2; ++p) {
                const ulonglong2 v = ld_stream_u64x2(src + 2 * ((u32)p * nthreads + tid));
                h[2 * p] = v.x; h[2 * p + 1] = v.y;
            }
        } else {
#pragma unroll
            for (int p = 0; p < SPLIT_KPT / 2; ++p) {
                const u32 i = 2 * ((u32)p * nthreads + tid);
                h[2 * p] = i < n ? ld_stream_u64(src + i) : 0;
                h[2 * p + 1] = i + 1 < n ? ld_stream_u64(src + i + 1) : 0;
            }
        }
#pragma unroll
        for (int q = 0; q < SPLIT_KPT; ++q) {
            const u32 i = 2 * ((u32)(q >> 1) * nthreads + tid) + (u32)(q & 1);
            sp[q] = 0;
            if (i < n) {
                const u32 sub = (u32)(h[q] >> sub_shift) & (F - 1u);
                sp[q] = (sub << 16) | atomicAdd(&s_cnt[sub], 1u);
            }
        }
        __syncthreads();
        // 2. one reservation per non-empty sub-region, issued first: the round trips of these global atomics (all in flight
        //    together) are covered by the prefix scan and the sort below -- the result is only needed for the copy-out
        const bool one_each = nthreads >= F;         // (the CPU emulation runs one thread: it reserves in a loop, below)
        u64 my_base = 0;
        if (one_each && tid < F && s_cnt[tid]) my_base = atomicAdd(&sub_cursor[(u64)bucket * F + tid], (u64)s_cnt[tid]);
        //    exclusive prefix of the counts (one warp, a run of consecutive sub-regions per lane)
        if (tid < wsize) {
            const u32 per = (F + wsize - 1) / wsize, f0 = tid * per;
            u32 sum = 0;
            for (u32 i = 0; i < per; ++i) if (f0 + i < F) sum += s_cnt[f0 + i];
            u32 incl = sum;
            for (u32 o = 1; o < wsize; o <<= 1) {
                const u32 v = __shfl_up_sync(0xffffffffu, incl, (int)o);
                if (tid >= o) incl += v;
            }
            u32 run = incl - sum;
            for (u32 i = 0; i < per; ++i) if (f0 + i < F) { s_off[f0 + i] = run; run += s_cnt[f0 + i]; }
        }
        __syncthreads();
        // 3. keys into sub-region order; per sub-region: where sorted position p goes (s_dst[f] + p, an index into sub_keys) and the
        //    first position that no longer fits the segment (s_lim[f])
#pragma unroll
        for (int q = 0; q < SPLIT_KPT; ++q) {
            const u32 i = 2 * ((u32)(q >> 1) * nthreads + tid) + (u32)(q & 1);
            if (i < n) s_sorted[s_off[sp[q] >> 16] + (sp[q] & 0xFFFFu)] = h[q];
        }
        for (u32 f = tid; f < F; f += nthreads) {
            const u32 c = s_cnt[f], off = s_off[f];
            const u64 g0 = one_each ? my_base : (c ? atomicAdd(&sub_cursor[(u64)bucket * F + f], (u64)c) : 0ull);
            const u64 room = g0 < sub_cap ? sub_cap - g0 : 0ull;
            s_dst[f] = ((u64)bucket * F + f) * sub_cap + g0 - off;
            s_lim[f] = off + (u32)min((u64)c, room);
        }
        __syncthreads();
        // 4. copy-out in sorted order: consecutive positions of a sub-region are consecutive addresses of its segment (a full tile
        //    brings tile_keys / F keys for each sub-region: 256 B at F = 256); the sub-region of a key is in its own hash bits.
        //    (The first version gave every sub-region's run to a group of 16 lanes: 41 % of the kernel's instructions went
        //    into the per-run address arithmetic, profiles/r2l_split_build_kernels.txt)
        for (u32 p = tid; p < n; p += nthreads) {
            const u64 hk = s_sorted[p];
            const u32 f = (u32)(hk >> sub_shift) & (F - 1u);
            if (p < s_lim[f]) st_stream_u64(sub_keys + (s_dst[f] + p), hk);
            else spill_stored<1>(&hk, ctr, ovf, ovf_cap);            // segment full: through the overflow list
        }
        __syncthreads();                             // s_cnt / s_off / s_sorted are rewritten by the next tile
    }
}

__global__ void __launch_bounds__(SPLIT_THREADS, 2)
split_kernel(const u64 *__restrict__ bkt_hash, u64 seg_cap, const PassBBucket *__restrict__ bk, u32 d_first, u32 d_end,
             u32 F, int sub_shift, u64 *__restrict__ sub_keys, u64 sub_cap, u64 *sub_cursor, Counters *ctr, u64 *ovf, u64 ovf_cap)
{
    split_body<false>(bkt_hash, nullptr, 1u, seg_cap, bk, d_first, d_end, F, sub_shift, sub_keys, sub_cap, sub_cursor, ctr, ovf, ovf_cap);
}
__global__ void __launch_bounds__(SPLIT_THREADS, 2)
split_gather_kernel(const __grid_constant__ KeyxSources srcs, u32 n_src, u64 seg_cap, const PassBBucket *__restrict__ bk, u32 d_first,
                    u32 d_end, u32 F, int sub_shift, u64 *__restrict__ sub_keys, u64 sub_cap, u64 *sub_cursor, Counters *ctr,
                    u64 *ovf, u64 ovf_cap)
{
    split_body<true>(nullptr, &srcs, n_src, seg_cap, bk, d_first, d_end, F, sub_shift, sub_keys, sub_cap, sub_cursor, ctr, ovf, ovf_cap);
}

// One probe of the shared-memory copy of a sub-region: slot (home + d) for the key with hash h.  Returns BUILD_DONE when the
// key is settled (counted, claimed, or -- probe sequence leaving the sub-region or passing CT_MAX_DISP -- handed to the
// overflow list), else the displacement to try next.
// The slot word is read and claimed as 64 bits but LOOKED AT as two 32-bit halves (the pass is bound by instruction issue, and
// 64-bit shifts / compares by a run-time amount cost two to three instructions each): with q = log2(capacity) <= 31 the low
// half holds count (q - 7 bits), displacement + 1 (7 bits) and the low 32 - q bits of the remainder, the high half the rest.
constexpr u32 BUILD_DONE = 0xFFFFFFFFu;
struct BuildGeom { int rbits, q, cbits; u32 cmask; };
__device__ __forceinline__ u32 build_probe(u64 *s_tab, const BuildGeom &bg, u64 h, u32 d, u32 &newk, Counters *ctr, u64 *ovf, u64 ovf_cap)
{
    const u32 idx = ((u32)(h >> bg.rbits) & (BUILD_SLOTS - 1u)) + d;
    if (idx >= BUILD_SLOTS || d > (u32)CT_MAX_DISP) {
        const u64 k0 = fmix64_inverse(h);
        spill_key<1>(&k0, ctr, ovf, ovf_cap);
        return BUILD_DONE;
    }
    u64 *s = s_tab + idx;
    const u32 tag_hi = (u32)(h >> (32 - bg.q));                   // bits 32..63 of h << q
    const u32 tag_lo = ((u32)h << bg.q) | ((d + 1u) << bg.cbits);
    const u64 cur = *reinterpret_cast<volatile u64 *>(s);         // ONE 64-bit read: a slot goes from 0 to (tag | 1) in one step
    u32 lo = (u32)cur, hi = (u32)(cur >> 32);
    if ((lo | hi) == 0u) {
        const u64 old = atomicCAS(s, 0ull, ((u64)tag_hi << 32) | (u64)(tag_lo | 1u));
        if (old == 0) { ++newk; return BUILD_DONE; }              // the slot was empty and is ours, count 1
        lo = (u32)old; hi = (u32)(old >> 32);
    }
    if (hi == tag_hi && ((lo ^ tag_lo) >> bg.cbits) == 0u) {      // our key: +1 on the count field, unless saturated
        if ((lo & bg.cmask) < COUNT_SAT) atomicAdd(reinterpret_cast<u32 *>(s), 1u);
        return BUILD_DONE;
    }
    return d + 1u;
}

// Sub-regions [g_first, g_end) of the table, one CTA at a time each; load_existing = 0: the table is known to be all zero
// (nothing has touched it since it was cleared), so the sub-region is not read.
// The first version probed in a loop per key: linear probing has a long tail (a hot k-mer deep in a cluster is hit by every one of
// its instances), and a warp runs as long as its slowest lane -- ncu counted 7.4 active lanes per instruction and 550 warp
// instructions per 32 keys (profiles/r2k_split_build_v1_kernels.txt; 6.5 ms).  So, as in bucket_insert_compact_kernel: every key
// gets ONE probe per round; what is not settled goes on a small per-warp list in shared memory (the key's hash with the
// displacement to try next in its top seven bits -- those bits are the same for all keys of a sub-region) and rides with a
// later round, one list entry per lane.  Warps never wait for each other between the CTA barriers around load and write-back.
// That version ran in 2.07 ms, bound by instruction issue (64 % issue-active, ~100 warp instructions per 32 probes) with 15 % of
// the samples waiting for the round's own key loads (profiles/r2l_split_build_kernels.txt): hence the 32-bit slot arithmetic above
// and the keys of the NEXT round being loaded before the current round's probes.
__global__ void __launch_bounds__(BUILD_THREADS, 2)
region_build_kernel(const u64 *__restrict__ sub_keys, u64 sub_cap, const u64 *__restrict__ sub_cursor, u64 g_first, u64 g_end,
                    int sub_shift, Table<1> table, int load_existing, Counters *ctr, u64 *ovf, u64 ovf_cap)
{
    PBK_DYN_SMEM(u64, s_mem);                        // BUILD_SLOTS slot words, then BUILD_DEF_CAP list entries per warp
    u64 *s_tab = s_mem;
    const u32 tid = threadIdx.x, nthreads = blockDim.x;
    const u32 wsize = nthreads < 32u ? nthreads : 32u;
    const u32 lane = tid % wsize, warp = tid / wsize;
    const u32 lt_mask = (1u << lane) - 1u;
    u64 *def = s_mem + BUILD_SLOTS + (size_t)warp * BUILD_DEF_CAP;
    const u32 warp_keys = wsize * BUILD_KPT, round_keys = nthreads * BUILD_KPT;
    BuildGeom bg;
    bg.rbits = table.g.rbits; bg.q = 64 - table.g.rbits; bg.cbits = table.g.cbits; bg.cmask = (u32)table.g.cmask;
    constexpr u64 DMASK = 0x7Full << 57;
    u32 newk = 0;
    for (u64 r = g_first + blockIdx.x; r < g_end; r += gridDim.x) {
        const u64 n = min(sub_cursor[r], sub_cap);
        if (n == 0) continue;                        // (same for every thread) nothing to add: the sub-region stays as it is
        u64 *region = table.slots + r * BUILD_SLOTS;
        const u64 *src = sub_keys + r * sub_cap;     // 16-byte aligned: sub_cap is even
        const u64 top7 = (r << sub_shift) & DMASK;   // bits 57..63 of every hash of this sub-region
        // the warp's slice of a round: BUILD_KPT keys per lane from key index o (two 16-byte loads per lane); bit q of the
        // returned mask = key q exists
        auto fetch = [&](u64 o, u64 *k) -> u32 {
            u32 live = 0;
#pragma unroll
            for (int p = 0; p < BUILD_KPT / 2; ++p) {
                const u64 i = o + 2 * ((u64)p * wsize + lane);
                k[2 * p] = k[2 * p + 1] = 0;
                if (i + 1 < n) {
                    const ulonglong2 v = ld_stream_u64x2(src + i);
                    k[2 * p] = v.x; k[2 * p + 1] = v.y;
                    live |= 3u << (2 * p);
                } else if (i < n) {
                    k[2 * p] = ld_stream_u64(src + i);
                    live |= 1u << (2 * p);
                }
            }
            return live;
        };
        u64 nx[BUILD_KPT];
        u32 nx_live = fetch((u64)warp * warp_keys, nx);          // in flight while the sub-region is loaded / cleared
        if (load_existing) {
            for (u32 i = tid; i < BUILD_SLOTS / 2; i += nthreads) {
                const ulonglong2 v = ld_stream_u64x2(region + 2 * i);
                s_tab[2 * i] = v.x; s_tab[2 * i + 1] = v.y;
            }
        } else {
            for (u32 i = tid; i < BUILD_SLOTS; i += nthreads) s_tab[i] = 0;
        }
        __syncthreads();
        u32 n_def = 0;
        // settle or re-list what one probe of (hq, dq) left
        auto settle = [&](bool live, u64 hq, u32 dq) {
            u32 m = BUILD_DONE;
            if (live) m = build_probe(s_tab, bg, hq, dq, newk, ctr, ovf, ovf_cap);
            const unsigned bal = __ballot_sync(0xffffffffu, m != BUILD_DONE);
            if (bal == 0) return;
            if (m != BUILD_DONE) def[n_def + (u32)__popc(bal & lt_mask)] = (hq & ~DMASK) | ((u64)m << 57);
            n_def += (u32)__popc(bal);
        };
        // one list entry per lane (the top min(n_def, warp size) entries)
        auto list_round = [&]() {
            const u32 take = n_def < wsize ? n_def : wsize;
            u64 xe = 0;
            __syncwarp();
            if (lane < take) xe = def[n_def - 1 - lane];
            __syncwarp();
            n_def -= take;
            settle(lane < take, (xe & ~DMASK) | top7, (u32)(xe >> 57));
        };
        for (u64 o = (u64)warp * warp_keys; o < n; o += round_keys) {
            u64 h[BUILD_KPT];
            const u32 live = nx_live;
#pragma unroll
            for (int q = 0; q < BUILD_KPT; ++q) h[q] = nx[q];
            nx_live = fetch(o + round_keys, nx);                 // the next round's keys, in flight during this round's probes
#pragma unroll 1
            while (n_def > wsize) list_round();                  // a round may only start with at most one batch listed
#pragma unroll
            for (int q = 0; q < BUILD_KPT; ++q) settle((live >> q) & 1u, h[q], 0u);
            list_round();
        }
#pragma unroll 1
        while (n_def) list_round();                  // nothing ever waits, so the list drains
        __syncthreads();
        for (u32 i = tid; i < BUILD_SLOTS / 2; i += nthreads)
            st_cg_u64x2(region + 2 * i, s_tab[2 * i], s_tab[2 * i + 1]);
        __syncthreads();                             // the next sub-region overwrites s_tab
    }
    newk = warp_sum_u32(newk);
    if ((tid & 31) == 0 && newk) atomicAdd(&ctr->new_keys, (u64)newk);
}

// Pass B for multi-word keys (k > 32), batched like the one-word kernel: a thread works on PASSBW_KPT<W> new keys plus
// one deferred key per round; the state words of all their slots are loaded together, then the claims (CAS) and the
// key words of the occupied slots, then the matching keys get their reduction.  A slot that is locked by a claimer, a
// lost claim or a slot held by another key is not waited for: the key goes on the per-warp list (as an index into the
// bucket store, 10 bytes for any W) and retries with the next round.  bucket_insert_kernel<W> above is the one-key-
// at-a-time form of the same protocol (kept selectable: PBK_WIDE_SERIAL=1).
// measured on C1 (scripts/tune_variants.sh, profiles/r2b_tune_variants.jsonl): 2 new keys + 1 deferred per thread and round at 3
// CTAs per SM (76 registers) beats 4 + 1 at 2 CTAs (126 registers): k = 75 Pass B 6.6 -> 5.8 ms, k = 42 7.7 -> 6.8 ms
#ifndef PBK_PASSBW_KPT
#define PBK_PASSBW_KPT 2
#endif
#ifndef PBK_PASSBW_MINCTAS
#define PBK_PASSBW_MINCTAS 3
#endif
template <int W> struct PASSBW_KPT { static constexpr int value = W <= 3 ? PBK_PASSBW_KPT : 2; };
constexpr u32 PASSBW_REMOTE = 1u << 9;               // deferred-entry flag above the probe count (MAX_PROBE < 256)

// STAGED: the keys of a tile do not come through the threads' own global loads but through TMA bulk copies (cp.async.bulk ->
// shared memory, completion on an mbarrier): one thread issues the copy of the NEXT tile (one contiguous piece of the bucket
// store: tile_keys x 8 W bytes) while the CTA inserts the current one, so the only global-memory latency left on a round's
// critical path is the table slot itself, and the bucket store is read in whole 128-byte lines instead of 8-byte loads with a
// stride of 8 W bytes.  Needs 2 x tile_keys x 8 W bytes of dynamic shared memory (96 KB at W = 3) and segments whose byte
// offsets are 16-byte aligned (plan_partition keeps seg_cap even).
template <int W, bool STAGED>
__global__ void __launch_bounds__(PASSB_THREADS, PBK_PASSBW_MINCTAS)
bucket_insert_wide_kernel(const u64 *__restrict__ bkt_keys, u64 seg_cap, const PassBBucket *__restrict__ bk,
                          u32 b_first, u32 b_end, u64 *ticket, Table<W> table, Table<W> remote, u32 n_shards, u32 rank,
                          Counters *ctr, u64 *ovf, u64 ovf_cap)
{
    constexpr int KPT = PASSBW_KPT<W>::value, ROUNDS = PASSB_KPT / KPT, E = KPT + 1;
    constexpr int DEF_CAP = 32 * (KPT + 2);
    PBK_DYN_SMEM(u64, s_stage);                      // STAGED: two tiles of keys
    __shared__ u64 s_mbar[2];
    __shared__ u64 s_ticket[2];
    __shared__ u64 s_def_i[PASSB_THREADS / 32][DEF_CAP];
    __shared__ uint16_t s_def_m[PASSB_THREADS / 32][DEF_CAP];
    const u32 nb = b_end - b_first, tid = threadIdx.x, nthreads = blockDim.x;
    const u32 wsize = nthreads < 32u ? nthreads : 32u;
    const u32 lane = tid % wsize, warp = tid / wsize;
    u64 *def_i = s_def_i[warp];
    uint16_t *def_m = s_def_m[warp];
    u32 n_def = 0, newk = 0, newr = 0, lb = 0;
    const u32 round_keys = nthreads * KPT, tile_keys = round_keys * ROUNDS;
    if (tid == 0) { s_ticket[0] = atomicAdd(ticket, 1ull); s_ticket[1] = atomicAdd(ticket, 1ull); }
    __syncthreads();
    const u64 n_tiles = bk[nb].tile_start;

    // one round: `n_new` keys starting at key index `first` (0 = only deferred keys); STAGED: the same keys sit in shared
    // memory at `staged`
    auto round = [&](u64 first, u32 n_new, const u64 *staged) {
        u64 key[E][W], kw[E][W], gi[E];
        Slot<W> *sp[E];
        u32 d[E], cs[E], act[E], fold[E];           // act: 0 none, 1 compare, 2 claim attempt, 3 locked (retry); fold: see key_fold32
        bool live[E], rem[E];
        // deferred key of an earlier round (top of the warp's list)
        const u32 take = n_def < wsize ? n_def : wsize;
        __syncwarp();
        live[KPT] = lane < take;
        gi[KPT] = live[KPT] ? def_i[n_def - 1 - lane] : 0;
        const u32 xm = live[KPT] ? def_m[n_def - 1 - lane] : 0u;
        __syncwarp();
        n_def -= take;
        d[KPT] = xm & 0xFFu;
#pragma unroll
        for (int q = 0; q < KPT; ++q) {
            const u32 i = (u32)q * nthreads + tid;
            live[q] = i < n_new;
            gi[q] = first + i;
            d[q] = 0;
        }
#pragma unroll
        for (int e = 0; e < E; ++e) {
            if (STAGED && e < KPT && staged != nullptr) {          // this round's new keys: already in shared memory
                const u32 i = (u32)e * nthreads + tid;
#pragma unroll
                for (int w = 0; w < W; ++w) key[e][w] = live[e] ? staged[(u64)i * W + w] : 0;
            } else if constexpr (W == 2) {            // 16-byte entries: one vector load
                ulonglong2 v = make_ulonglong2(0, 0);
                if (live[e]) v = ld_stream_u64x2(bkt_keys + gi[e] * 2);
                key[e][0] = v.x; key[e][1] = v.y;
            } else {
#pragma unroll
                for (int w = 0; w < W; ++w) key[e][w] = live[e] ? ld_stream_u64(bkt_keys + gi[e] * W + w) : 0;
            }
        }
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const u64 h = hash_key<W>(key[e]);
            rem[e] = n_shards > 1 && shard_of_hash(h, n_shards) != rank;
            const Table<W> &tb = rem[e] ? remote : table;
            u64 idx = __umul64hi(h, tb.cap) + d[e];
            if (idx >= tb.cap) idx -= tb.cap;
            sp[e] = tb.slots + idx;
            cs[e] = 0; fold[e] = 0;
            if constexpr (W <= 3) {                   // sector-sized slot: key words and state word in ONE request
                if (live[e]) {
                    u64 q[4];
                    ld_cg_256(sp[e], q[0], q[1], q[2], q[3]);
#pragma unroll
                    for (int w = 0; w < W; ++w) kw[e][w] = q[w];
                    cs[e] = (u32)q[W];
                    fold[e] = (u32)(q[W] >> 32);
                }
            } else {
                if (live[e]) cs[e] = ld_cg_u32(&sp[e]->cs);
            }
        }
#pragma unroll
        for (int e = 0; e < E; ++e) {
            act[e] = 0;
            if (!live[e]) continue;
            if (cs[e] == 0) { act[e] = 2; cs[e] = atomicCAS(&sp[e]->cs, 0u, CS_LOCKED); }
            else if (cs[e] == CS_LOCKED) act[e] = 3;
            else {
                act[e] = 1;
                if constexpr (W > 3) {
#pragma unroll
                    for (int w = 0; w < W; ++w) kw[e][w] = ld_cg_u64(&sp[e]->key[w]);
                }
            }
        }
#pragma unroll
        for (int e = 0; e < E; ++e) {
            u32 m = 0xFFFFu;                         // 0xFFFF = finished, else the list entry of the retry
            if (act[e] == 2) {
                if (cs[e] == 0) {                    // the slot is ours
                    if constexpr (W <= 3) {
                        // sector-sized slot: key words and count 1 in ONE 256-bit store (the release fence of the
                        // word-by-word form cost as many stall cycles as all the loads of the kernel together)
                        u64 q[4] = {0, 0, 0, 0};
#pragma unroll
                        for (int w = 0; w < W; ++w) q[w] = key[e][w];
                        q[W] = 1ull | ((u64)key_fold32<W>(key[e]) << 32);
                        st_cg_256(sp[e], q[0], q[1], q[2], q[3]);
                    } else {                         // key words first, then the count publishes them
#pragma unroll
                        for (int w = 0; w < W; ++w) st_cg_u64(&sp[e]->key[w], key[e][w]);
                        st_release_u32(&sp[e]->cs, 1u);
                    }
                    if (rem[e]) ++newr; else ++newk;
                } else {
                    m = d[e] | (rem[e] ? PASSBW_REMOTE : 0u);           // somebody else got it: look again next round
                }
            } else if (act[e] == 3) {
                m = d[e] | (rem[e] ? PASSBW_REMOTE : 0u);
            } else if (act[e] == 1) {
                bool eq = true;
#pragma unroll
                for (int w = 0; w < W; ++w) eq &= (kw[e][w] == key[e][w]);
                bool torn = false;                   // W <= 3: the words read do not fold to what the state word says -> look again
                if constexpr (W <= 3) torn = !eq && key_fold32<W>(kw[e]) != fold[e];
                if (eq) red_add_u32(&sp[e]->cs, 1u);
                else if (torn) m = d[e] | (rem[e] ? PASSBW_REMOTE : 0u);
                else if (d[e] + 1 >= (u32)MAX_PROBE) spill_key<W>(key[e], ctr, ovf, ovf_cap);
                else m = (d[e] + 1) | (rem[e] ? PASSBW_REMOTE : 0u);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, m != 0xFFFFu);
            if (bal == 0) continue;
            if (m != 0xFFFFu) {
                const u32 pos = n_def + (u32)__popc(bal & ((1u << lane) - 1u));
                def_i[pos] = gi[e];
                def_m[pos] = (uint16_t)m;
            }
            n_def += (u32)__popc(bal);
        }
    };

    // STAGED: thread 0 starts the bulk copy of tile `tt` into stage buffer `buf`
    u32 lb_issue = 0;
    auto issue_tile = [&](u64 tt, int buf) {
        while (bk[lb_issue + 1].tile_start <= tt) ++lb_issue;
        const u64 jj = tt - bk[lb_issue].tile_start;
        const u64 first_i = (u64)(b_first + lb_issue) * seg_cap + jj * tile_keys;
        const u64 n_i = min((u64)tile_keys, bk[lb_issue].n_keys - jj * tile_keys);
        const u32 bytes = (u32)((n_i * W * 8 + 15) & ~15ull);
        mbar_expect_tx(&s_mbar[buf], bytes);
        bulk_copy_g2s(s_stage + (size_t)buf * tile_keys * W, bkt_keys + first_i * W, bytes, &s_mbar[buf]);
    };
    int par = 0;
    u32 phase = 0;                                            // bit b = parity the next wait on stage buffer b expects
    u64 t = s_ticket[0];
    if constexpr (STAGED) {
        if (tid == 0) { mbar_init(&s_mbar[0], 1); mbar_init(&s_mbar[1], 1); mbar_fence_init(); }
        __syncthreads();
        if (tid == 0 && t < n_tiles) issue_tile(t, 0);
    }
    while (t < n_tiles) {
        while (bk[lb + 1].tile_start <= t) ++lb;
        const u64 j = t - bk[lb].tile_start, n = bk[lb].n_keys;
        const u64 nt = bk[lb + 1].tile_start - bk[lb].tile_start;
        const u64 t_next = s_ticket[par ^ 1];
        __syncthreads();                              // (also: everybody is done with stage buffer par ^ 1)
        if (tid == 0) {
            s_ticket[par] = atomicAdd(ticket, 1ull);
            if (STAGED && t_next < n_tiles) issue_tile(t_next, par ^ 1);
        }
        if (bk[lb].pf_base) passb_prefetch(bk[lb].pf_base, bk[lb].pf_lines, j, nt, tid, nthreads);
        if (bk[lb].pf_base2) passb_prefetch(bk[lb].pf_base2, bk[lb].pf_lines2, j, nt, tid, nthreads);
        const u64 first = (u64)(b_first + lb) * seg_cap + j * tile_keys;
        const u64 left = n - j * tile_keys;
        const u64 *stage = nullptr;
        if constexpr (STAGED) {
            mbar_wait(&s_mbar[par], (phase >> par) & 1u);      // this tile's keys have landed
            phase ^= 1u << par;
            stage = s_stage + (size_t)par * tile_keys * W;
        }
#pragma unroll 1
        for (u64 o = 0; o < left && o < tile_keys; o += round_keys) {
#pragma unroll 1
            while (n_def > wsize) round(0, 0u, nullptr);      // a round may only start with at most one batch listed
            round(first + o, (u32)min((u64)round_keys, left - o), STAGED ? stage + o * W : nullptr);
        }
        __syncthreads();
        t = t_next;
        par ^= 1;
    }
#pragma unroll 1
    while (n_def) round(0, 0u, nullptr);             // nobody ever waits, so the list drains
    newk = warp_sum_u32(newk);
    newr = warp_sum_u32(newr);
    if ((tid & 31) == 0) {
        if (newk) atomicAdd(&ctr->new_keys, (u64)newk);
        if (newr) atomicAdd(&ctr->new_keys_remote, (u64)newr);
    }
}

// =================================================================================================
// table maintenance
// =================================================================================================

template <int W>
__global__ void __launch_bounds__(256)
rehash_kernel(Table<W> from, Table<W> to, Counters *ctr)
{
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < from.cap; i += stride) {
        u64 key[W];
        u32 count;
        if (!from.load(i, key, &count)) continue;
        if (to.insert(key, hash_key<W>(key), count, false) < 0) atomicOr(&ctr->error_flags, ERR_OVERFLOW_LOST);
    }
}

template <int W>
__global__ void __launch_bounds__(256) clamp_kernel(Table<W> t)
{
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < t.cap; i += stride) t.clamp(i);
}

// clamp + ++occurrenceDistribution[count] (counter.h:496): one streaming pass over the table.  Low counts go through a
// shared-memory histogram, the long tail straight to global atomics; count 1 -- most of a read set's distinct k-mers are
// sequencing-error singletons, which would put nearly every shared-memory atomic of a warp on ONE address -- is tallied in a
// register.  A thread keeps HIST_ILP independent 16-byte loads in flight (the first version had one 8-byte load per thread
// outstanding and ran at a third of the HBM rate).
constexpr int HIST_SMEM_BINS = 4096;
constexpr int HIST_ILP = 4;
template <int W>
__global__ void __launch_bounds__(256) histogram_kernel(Table<W> t, u64 *occ_hist)
{
    __shared__ u32 sh[HIST_SMEM_BINS];
    for (int i = threadIdx.x; i < HIST_SMEM_BINS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const u64 stride = (u64)gridDim.x * blockDim.x, gtid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u32 ones = 0;
    auto tally = [&](u32 c) {
        if (c == 0) return;
        if (c == 1) { ++ones; return; }
        if (c > COUNT_SAT) c = COUNT_SAT;
        if (c < HIST_SMEM_BINS) atomicAdd(&sh[c], 1u);
        else atomicAdd(&occ_hist[c], 1ull);
    };
    if constexpr (W == 1) {                           // two 8-byte slots per load; the key is not needed here
        const u64 n2 = t.cap / 2;                     // capacity is a power of two >= 2^27 (tests/cpu_emul: any even number)
        const int cbits = t.g.cbits;
        const u64 cmask = t.g.cmask;
        for (u64 i0 = gtid; i0 < n2; i0 += stride * HIST_ILP) {
            ulonglong2 v[HIST_ILP];
#pragma unroll
            for (int j = 0; j < HIST_ILP; ++j) {
                const u64 i = i0 + (u64)j * stride;
                v[j] = i < n2 ? ld_stream_u64x2(t.slots + 2 * i) : make_ulonglong2(0, 0);
            }
#pragma unroll
            for (int j = 0; j < HIST_ILP; ++j) {
                if (v[j].x >> cbits) { const u64 c = v[j].x & cmask; tally(c > COUNT_SAT ? COUNT_SAT : (u32)c); }
                if (v[j].y >> cbits) { const u64 c = v[j].y & cmask; tally(c > COUNT_SAT ? COUNT_SAT : (u32)c); }
            }
        }
        if (t.cap & 1) { if (gtid == 0) { u64 key, c; if (ct_decode(t.slots[t.cap - 1], t.cap - 1, t.g, &key, &c)) tally(c > COUNT_SAT ? COUNT_SAT : (u32)c); } }
    } else {
        for (u64 i0 = gtid; i0 < t.cap; i0 += stride * HIST_ILP) {
            u32 c[HIST_ILP];
#pragma unroll
            for (int j = 0; j < HIST_ILP; ++j) {
                const u64 i = i0 + (u64)j * stride;
                c[j] = i < t.cap ? ld_cg_u32(&t.slots[i].cs) : 0u;
            }
#pragma unroll
            for (int j = 0; j < HIST_ILP; ++j) tally(c[j]);
        }
    }
    ones = warp_sum_u32(ones);
    if ((threadIdx.x & 31) == 0 && ones) atomicAdd(&sh[1], ones);
    __syncthreads();
    for (int i = threadIdx.x; i < HIST_SMEM_BINS; i += blockDim.x)
        if (sh[i]) atomicAdd(&occ_hist[i], (u64)sh[i]);
}

// Compaction of the entries with count >= min_count.  Output positions are reserved per CTA chunk (shared-memory
// counter, one global atomic per chunk): a warp-level reservation on the single global counter serialises millions of
// same-address atomics when the table is large.
template <int W>
__global__ void __launch_bounds__(256)
export_kernel(Table<W> t, u32 min_count, u64 *keys_out, uint16_t *counts_out, u64 capacity, u64 *d_n_out)
{
    __shared__ u32 s_n;
    __shared__ u64 s_base;
    constexpr int PER_THREAD = 8;
    const u64 chunk = (u64)blockDim.x * PER_THREAD;
    for (u64 c0 = (u64)blockIdx.x * chunk; c0 < t.cap; c0 += (u64)gridDim.x * chunk) {
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        u64 key[PER_THREAD][W];
        u32 count[PER_THREAD], rank[PER_THREAD];
        if constexpr (W == 1) {
            // compact slots: all eight loads go out first (streaming, nothing depends on them yet); a slot is decoded -- the
            // inverse of fmix64 -- only if it is occupied AND passes the count filter (1 % of the slots of a read set's table)
            u64 v[PER_THREAD];
#pragma unroll
            for (int q = 0; q < PER_THREAD; ++q) {
                const u64 i = c0 + (u64)q * blockDim.x + threadIdx.x;
                v[q] = i < t.cap ? ld_stream_u64(t.slots + i) : 0ull;
            }
#pragma unroll
            for (int q = 0; q < PER_THREAD; ++q) {
                rank[q] = 0xFFFFFFFFu; count[q] = 0;
                if ((v[q] >> t.g.cbits) == 0) continue;
                const u64 cc = v[q] & t.g.cmask;
                count[q] = cc > COUNT_SAT ? COUNT_SAT : (u32)cc;
                if (count[q] < min_count) continue;
                u64 dummy;
                ct_decode(v[q], c0 + (u64)q * blockDim.x + threadIdx.x, t.g, &key[q][0], &dummy);
                rank[q] = atomicAdd(&s_n, 1u);
            }
        } else {
#pragma unroll
            for (int q = 0; q < PER_THREAD; ++q) {
                const u64 i = c0 + (u64)q * blockDim.x + threadIdx.x;
                rank[q] = 0xFFFFFFFFu; count[q] = 0;
                if (i < t.cap && t.load(i, key[q], &count[q]) && count[q] >= min_count) rank[q] = atomicAdd(&s_n, 1u);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base = s_n ? atomicAdd(d_n_out, (u64)s_n) : 0ull;
        __syncthreads();
#pragma unroll
        for (int q = 0; q < PER_THREAD; ++q) {
            if (rank[q] == 0xFFFFFFFFu) continue;
            const u64 at = s_base + rank[q];
            if (at < capacity) {
#pragma unroll
                for (int j = 0; j < W; ++j) keys_out[at * W + j] = key[q][j];
                counts_out[at] = (uint16_t)count[q];
            }
        }
        __syncthreads();
    }
}

// Per-destination counts / record packing of the remote-staging table.  With few destinations nearly every warp
// would hit the same one or two global counters (measured at N = 2: 10 ms of serialised same-address atomics per
// step), so both kernels aggregate per CTA in shared memory first: cheap shared-memory atomics per entry, one global
// atomic per destination per CTA (count) or per CTA chunk (pack).
constexpr u32 SHARD_SMEM_MAX = 1024;                 // destinations handled through shared memory

template <int W>
__global__ void __launch_bounds__(256)
shard_count_kernel(Table<W> t, u32 n_shards, u64 *d_counts)
{
    __shared__ u32 s_cnt[SHARD_SMEM_MAX];
    const bool use_smem = n_shards <= SHARD_SMEM_MAX;
    if (use_smem) for (u32 i = threadIdx.x; i < n_shards; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    // contiguous range per CTA, so that a CTA's partial counts stay below 2^32
    const u64 per_cta = (t.cap + gridDim.x - 1) / gridDim.x;
    const u64 lo = (u64)blockIdx.x * per_cta, hi = min(t.cap, lo + per_cta);
    for (u64 i0 = lo; i0 < hi; i0 += (u64)blockDim.x << 20) {          // flush before a counter could wrap
        const u64 i1 = min(hi, i0 + ((u64)blockDim.x << 20));
        for (u64 i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
            u64 key[W];
            u32 count;
            if (!t.load(i, key, &count)) continue;
            const u32 dest = shard_of_hash(hash_key<W>(key), n_shards);
            if (use_smem) atomicAdd(&s_cnt[dest], 1u); else atomicAdd(&d_counts[dest], 1ull);
        }
        __syncthreads();
        if (use_smem)
            for (u32 d = threadIdx.x; d < n_shards; d += blockDim.x) {
                if (s_cnt[d]) atomicAdd(&d_counts[d], (u64)s_cnt[d]);
                s_cnt[d] = 0;
            }
        __syncthreads();
    }
}

template <int W>
__global__ void __launch_bounds__(256)
shard_pack_kernel(Table<W> t, u32 n_shards, u64 *d_cursors, u64 *rec_out)
{
    __shared__ u32 s_cnt[SHARD_SMEM_MAX];
    __shared__ u64 s_base[SHARD_SMEM_MAX];
    const bool use_smem = n_shards <= SHARD_SMEM_MAX;
    constexpr int PER_THREAD = 8;                                       // slots per thread and chunk
    const u64 chunk = (u64)blockDim.x * PER_THREAD;
    for (u64 c0 = (u64)blockIdx.x * chunk; c0 < t.cap; c0 += (u64)gridDim.x * chunk) {
        if (use_smem) for (u32 i = threadIdx.x; i < n_shards; i += blockDim.x) s_cnt[i] = 0;
        __syncthreads();
        u64 key[PER_THREAD][W];
        u32 count[PER_THREAD], dest[PER_THREAD], rank[PER_THREAD];
#pragma unroll
        for (int q = 0; q < PER_THREAD; ++q) {
            const u64 i = c0 + (u64)q * blockDim.x + threadIdx.x;
            dest[q] = 0xFFFFFFFFu; rank[q] = 0; count[q] = 0;
            if (i < t.cap && t.load(i, key[q], &count[q])) {
                dest[q] = shard_of_hash(hash_key<W>(key[q]), n_shards);
                if (use_smem) rank[q] = atomicAdd(&s_cnt[dest[q]], 1u);
            }
        }
        __syncthreads();
        if (use_smem)
            for (u32 d = threadIdx.x; d < n_shards; d += blockDim.x)
                s_base[d] = s_cnt[d] ? atomicAdd(&d_cursors[d], (u64)s_cnt[d]) : 0ull;
        __syncthreads();
#pragma unroll
        for (int q = 0; q < PER_THREAD; ++q) {
            if (dest[q] == 0xFFFFFFFFu) continue;
            const u64 at = use_smem ? s_base[dest[q]] + rank[q] : atomicAdd(&d_cursors[dest[q]], 1ull);
#pragma unroll
            for (int j = 0; j < W; ++j) rec_out[at * (W + 1) + j] = key[q][j];
            rec_out[at * (W + 1) + W] = count[q];
        }
        __syncthreads();
    }
}

}  // namespace pbk
