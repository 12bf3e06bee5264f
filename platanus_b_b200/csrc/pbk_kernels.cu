// pbk_kernels.cu -- hand-written sm_100a kernels of the k-mer counting path.
//
// None of this is GEMM-shaped: it is byte/integer streaming plus scattered read-modify-write, so
// there is deliberately no tensor-core code.  What matters here is coalesced 128-bit loads, one HBM
// sector per table access, enough resident warps to cover DRAM latency, and grids sized from the
// SM count (148 on B200).
#include "pbk_kernels.cuh"
#include "pbk_kernels_impl.cuh"

#include <algorithm>
#include <cstdlib>
#ifndef PBK_CPU_EMUL                    // (tests/cpu_emul compiles this file for the host: no CUB, no microbenchmark)
#include <cub/device/device_radix_sort.cuh>
#endif

namespace pbk {

static inline int grid_for(u64 n_threads, int block, int sm_count, int blocks_per_sm)
{
    u64 need = (n_threads + block - 1) / block;
    u64 cap = (u64)sm_count * blocks_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

#define PBK_DISPATCH_W(words, CALL)                                                                  \
    switch (words) {                                                                                \
    case 1: { constexpr int W = 1; CALL; } break;                                                    \
    case 2: { constexpr int W = 2; CALL; } break;                                                    \
    case 3: { constexpr int W = 3; CALL; } break;                                                    \
    case 4: { constexpr int W = 4; CALL; } break;                                                    \
    case 5: { constexpr int W = 5; CALL; } break;                                                    \
    case 6: { constexpr int W = 6; CALL; } break;                                                    \
    case 7: { constexpr int W = 7; CALL; } break;                                                    \
    case 8: { constexpr int W = 8; CALL; } break;                                                    \
    default: break;                                                                                  \
    }

void launch_pack(const uint8_t *bases, u64 n_valid, u64 n_words, int encoding,
                 u64 *stream, u32 *nflag, u64 word0, Counters *ctr, cudaStream_t st)
{
    if (n_words == 0) return;
    int grid = grid_for(n_words, 256, 148, 16);
    bool aligned = (reinterpret_cast<uintptr_t>(bases) & 15) == 0;
    if (aligned) pack_kernel<true><<<grid, 256, 0, st>>>(bases, n_valid, n_words, encoding, stream, nflag, word0, ctr);
    else pack_kernel<false><<<grid, 256, 0, st>>>(bases, n_valid, n_words, encoding, stream, nflag, word0, ctr);
}

void launch_read_marks(const u64 *offsets, u64 n_reads, u64 *len_hist, u32 *rflag, Counters *ctr, cudaStream_t st)
{
    if (n_reads == 0) return;
    read_marks_kernel<<<grid_for(n_reads, 256, 148, 8), 256, 0, st>>>(offsets, n_reads, len_hist, rflag, ctr);
}

void launch_npos_abs_scatter(const u64 *n_positions, u64 n_n, u64 n_bases, u32 *nflag, cudaStream_t st)
{
    npos_abs_scatter_kernel<<<grid_for(n_n + 1, 256, 148, 4), 256, 0, st>>>(n_positions, n_n, n_bases, nflag);
}

void launch_npos_scatter(const u64 *offsets, const int32_t *n_pos, const u64 *n_pos_offsets, u64 n_reads,
                         u32 *nflag, cudaStream_t st)
{
    if (n_reads == 0) return;
    npos_scatter_kernel<<<grid_for(n_reads, 256, 148, 8), 256, 0, st>>>(offsets, n_pos, n_pos_offsets, n_reads, nflag);
}

// =================================================================================================
// dispatch on W = ceil(k/32)
// =================================================================================================


void launch_count(const u64 *stream, const u32 *nflag, const u32 *rflag, u64 word_begin, u64 word_end, int k,
                  TableView table, TableView remote, ShardInfo shard, Counters *ctr,
                  u64 *overflow_keys, u64 overflow_cap, int sm_count, cudaStream_t st, u32 pass)
{
    if (word_end <= word_begin) return;
    const int grid = grid_for(word_end - word_begin, 256, sm_count, 8);
    // hash-range pass on an unsharded context: the ownership test of the sharded form, with foreign keys skipped
    const u32 n_sh = pass ? ((pass >> 16) | PASS_ONLY_BIT) : shard.n_shards, rank = pass ? (pass & 0xFFFFu) : shard.rank;
    PBK_DISPATCH_W(table.words,
        (count_kernel<W><<<grid, 256, 0, st>>>(stream, nflag, rflag, word_begin, word_end, k,
            Table<W>(table.slots, table.cap), Table<W>(remote.slots, remote.cap),
            n_sh, rank, ctr, overflow_keys, overflow_cap)));
}

void launch_lookup(const u64 *stream, const u32 *nflag, const u32 *rflag, u64 word_begin, u64 word_end, int k,
                   TableView table, uint16_t *occ, int sm_count, cudaStream_t st)
{
    if (word_end <= word_begin) return;
    const int grid = grid_for(word_end - word_begin, 256, sm_count, 8);
    PBK_DISPATCH_W(table.words,
        (lookup_kernel<W><<<grid, 256, 0, st>>>(stream, nflag, rflag, word_begin, word_end, k,
            Table<W>(table.slots, table.cap), (u64 *)occ)));
}

void launch_contig_max(const u64 *stream, const u32 *nflag, const u32 *rflag, u64 word_begin, u64 word_end, int k,
                       TableView table, const uint16_t *val, Counters *ctr, int sm_count, cudaStream_t st)
{
    if (word_end <= word_begin) return;
    const int grid = grid_for(word_end - word_begin, 256, sm_count, 8);
    PBK_DISPATCH_W(table.words,
        (contig_max_kernel<W><<<grid, 256, 0, st>>>(stream, nflag, rflag, word_begin, word_end, k,
            Table<W>(table.slots, table.cap), val, ctr)));
}

void launch_override_records(const u64 *records, u64 n, TableView table, Counters *ctr, u64 *overflow_keys, u64 overflow_cap,
                             int sm_count, cudaStream_t st)
{
    if (n == 0) return;
    const int grid = grid_for(n, 256, sm_count, 8);
    PBK_DISPATCH_W(table.words,
        (override_records_kernel<W><<<grid, 256, 0, st>>>(records, n, Table<W>(table.slots, table.cap), ctr, overflow_keys, overflow_cap)));
}

void launch_neighbor_flags(const u64 *keys, u64 n, int k, TableView table, u32 min_count, uint8_t *out, int sm_count, cudaStream_t st)
{
    if (n == 0) return;
    const int grid = grid_for(n, 256, sm_count, 8);
    PBK_DISPATCH_W(table.words, (neighbor_flags_kernel<W><<<grid, 256, 0, st>>>(keys, n, k, Table<W>(table.slots, table.cap), min_count, out)));
}

void launch_read_match(const u64 *offsets, u64 n_reads, const uint16_t *occ, int k, uint8_t *matched, int sm_count, cudaStream_t st)
{
    if (n_reads == 0) return;
    read_match_kernel<<<grid_for(n_reads, 256, sm_count, 8), 256, 0, st>>>(offsets, n_reads, occ, k, matched);
}

void launch_insert_records(const u64 *records, u64 n, bool weighted, TableView table, TableView remote,
                           ShardInfo shard, Counters *ctr, u64 *overflow_keys, u64 overflow_cap, int sm_count,
                           cudaStream_t st)
{
    if (n == 0) return;
    const int grid = grid_for(n, 256, sm_count, 8);
    PBK_DISPATCH_W(table.words,
        (insert_records_kernel<W><<<grid, 256, 0, st>>>(records, n, weighted ? 1 : 0, Table<W>(table.slots, table.cap),
            Table<W>(remote.slots, remote.cap), shard.n_shards, shard.rank,
            ctr, overflow_keys, overflow_cap)));
}

// ---- partitioned counting -------------------------------------------------------------------------
constexpr size_t PART_SMEM_BUDGET = 100 * 1024;      // bins; two CTAs per SM (227 KB) so one computes while one flushes
constexpr u64 PART_REGION_BYTES = 16ull << 20;       // table bytes per bucket: Pass B is flat from 4 to 32 MB, Pass A prefers few, large bins

PartitionPlan plan_partition(u64 est_table_bytes, u64 windows_ub, int words, size_t smem_budget)
{
    PartitionPlan p{};
    if (smem_budget == 0) smem_budget = getenv("PBK_PART_SMEM_KB") ? (size_t)atoi(getenv("PBK_PART_SMEM_KB")) * 1024 : PART_SMEM_BUDGET;
    if (words > 1) smem_budget -= 8 * 1024;                  // partition_compact_kernel's queue of live work items (static shared memory)
    const u64 entries = smem_budget / (8 * (size_t)words);
    const u64 region_bytes = getenv("PBK_REGION_MB") ? (u64)atoi(getenv("PBK_REGION_MB")) << 20 : PART_REGION_BYTES;
    u64 P = (est_table_bytes + region_bytes - 1) / region_bytes;
    P = std::max<u64>(P, 8);
    P = std::min<u64>(P, std::min<u64>((u64)PART_MAX_BUCKETS, entries / 24));
    p.n_buckets = (u32)P;
    p.bin_cap = (u32)(entries / P);
    // one tile = threads * win window ends (win = PART_WIN1, 16, 8 for 1, 2, >= 3 key words), ~85 % of them valid; keep the
    // mean bin fill at 60 % of the bin
    const double win = words == 1 ? (double)PART_WIN1 : words == 2 ? 16.0 : (double)PBK_PART_WIN3;
    u64 threads = (u64)(p.bin_cap * 0.6 * (double)P / (win * 0.85));
    threads = std::min<u64>(512, threads / 32 * 32);
    p.threads = (int)std::max<u64>(64, threads);
    p.smem = (size_t)p.n_buckets * p.bin_cap * 8 * words;
    p.seg_cap = (windows_ub / P + windows_ub / (P * 16) + 8192 + 1) & ~1ull;    // even: segment byte offsets stay 16-byte aligned for any W
    return p;
}

// key-exchange layout (pbk_keyx_*): n_dest x n_regions buckets, destination-major.  n_regions is a function of the
// number of shards alone and seg_cap of the largest batch of any rank, so every rank derives the same layout.
PartitionPlan plan_partition_keyx(u32 n_dest, u64 max_windows_any_rank, int words, size_t smem_budget)
{
    u32 R = getenv("PBK_KEYX_REGIONS") ? (u32)atoi(getenv("PBK_KEYX_REGIONS")) : 64u;
    while (R > 8 && (u64)R * n_dest > 256) R >>= 1;                 // 2, 4 shards: 64 regions; 8: 32; 16: 16
    while (R & (R - 1)) R &= R - 1;                                 // power of two
    R = std::max<u32>(R, 2);
    PartitionPlan p{};
    if (smem_budget == 0) smem_budget = getenv("PBK_PART_SMEM_KB") ? (size_t)atoi(getenv("PBK_PART_SMEM_KB")) * 1024 : PART_SMEM_BUDGET;
    const u64 entries = smem_budget / (8 * (size_t)words);
    const u64 P = std::min<u64>((u64)R * n_dest, (u64)PART_MAX_BUCKETS);
    p.n_buckets = (u32)P;
    p.bin_cap = (u32)std::max<u64>(1, entries / P);
    const double win = words == 1 ? (double)PART_WIN1 : words == 2 ? 16.0 : 8.0;
    u64 threads = (u64)(p.bin_cap * 0.6 * (double)P / (win * 0.85));
    threads = std::min<u64>(512, threads / 32 * 32);
    p.threads = (int)std::max<u64>(64, threads);
    p.smem = (size_t)p.n_buckets * p.bin_cap * 8 * words;
    p.seg_cap = (max_windows_any_rank / P + max_windows_any_rank / (P * 16) + 8192 + 1) & ~1ull;   // even: 16-byte aligned segments
    return p;
}

template <int W, bool KEYX, bool PASSF = false>
static void partition_launch_w(const u64 *stream, const u32 *nflag, const u32 *rflag, u64 word_begin, u64 word_end, int k,
                               const PartitionPlan &plan, u64 *bkt_keys, u64 *bkt_cursor, Counters *ctr,
                               u64 *overflow_keys, u64 overflow_cap, int grid, u32 n_dest, cudaStream_t st)
{
    cudaFuncSetAttribute(partition_kernel<W, KEYX, PASSF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
    partition_kernel<W, KEYX, PASSF><<<grid, plan.threads, plan.smem, st>>>(stream, nflag, rflag, word_begin, word_end, k,
        plan.n_buckets, plan.bin_cap, bkt_keys, plan.seg_cap, bkt_cursor, ctr, overflow_keys, overflow_cap, n_dest);
}

// pass != 0: (n_passes << 16) | pass_index -- only the keys of that hash range are bucketed (pbk_config.n_passes; not with keyx_dest)
void launch_partition(const u64 *stream, const u32 *nflag, const u32 *rflag, u64 word_begin, u64 word_end, int k,
                      int words, const PartitionPlan &plan, u64 *bkt_keys, u64 *bkt_cursor, Counters *ctr,
                      u64 *overflow_keys, u64 overflow_cap, int sm_count, cudaStream_t st, u32 keyx_dest, u32 pass)
{
    if (word_end <= word_begin) return;
    const u64 subs = words == 1 ? 32 / PART_WIN1 : words == 2 ? 2 : 32 / PBK_PART_WIN3;       // 32 / PART_WIN<W>: work items per stream word
    const u64 tiles = ((word_end - word_begin) * subs + plan.threads - 1) / plan.threads;
    const int ctas = std::max(1, std::min(8, (int)((220 * 1024) / (plan.smem + 7 * 1024))));     // resident CTAs per SM
    const int grid = (int)std::min<u64>(tiles, (u64)sm_count * ctas);
    if (keyx_dest > 1) {                                        // one-word keys only (pbk_keyx_plan refuses k > 32)
        if (words == 1)
            partition_launch_w<1, true>(stream, nflag, rflag, word_begin, word_end, k, plan, bkt_keys, bkt_cursor, ctr,
                                        overflow_keys, overflow_cap, grid, keyx_dest, st);
        return;
    }
    // multi-word keys: live work items compacted across the CTA (partition_compact_kernel) unless PBK_PART_COMPACT=0; its queue takes
    // 8 KB of static shared memory, which plan_partition leaves free for these key widths
    static const bool compact = !(getenv("PBK_PART_COMPACT") && atoi(getenv("PBK_PART_COMPACT")) == 0);
    if (words > 1 && compact && plan.threads <= PARTC_MAX_THREADS) {
        switch (words) {
#define PBK_CASE_W(Wv) case Wv:                                                                                                        \
            if (pass) {                                                                                                                \
                cudaFuncSetAttribute(partition_compact_kernel<Wv, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem); \
                partition_compact_kernel<Wv, true><<<grid, plan.threads, plan.smem, st>>>(stream, nflag, rflag, word_begin, word_end, k, \
                    plan.n_buckets, plan.bin_cap, bkt_keys, plan.seg_cap, bkt_cursor, ctr, overflow_keys, overflow_cap, pass);         \
            } else {                                                                                                                   \
                cudaFuncSetAttribute(partition_compact_kernel<Wv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);       \
                partition_compact_kernel<Wv><<<grid, plan.threads, plan.smem, st>>>(stream, nflag, rflag, word_begin, word_end, k,      \
                    plan.n_buckets, plan.bin_cap, bkt_keys, plan.seg_cap, bkt_cursor, ctr, overflow_keys, overflow_cap, 0u);           \
            } break;
        PBK_CASE_W(2) PBK_CASE_W(3) PBK_CASE_W(4) PBK_CASE_W(5) PBK_CASE_W(6) PBK_CASE_W(7) PBK_CASE_W(8)
#undef PBK_CASE_W
        default: break;
        }
        return;
    }
    if (pass) {
        PBK_DISPATCH_W(words,
            (partition_launch_w<W, false, true>(stream, nflag, rflag, word_begin, word_end, k, plan, bkt_keys, bkt_cursor, ctr,
                                                overflow_keys, overflow_cap, grid, pass, st)));
        return;
    }
    PBK_DISPATCH_W(words,
        (partition_launch_w<W, false>(stream, nflag, rflag, word_begin, word_end, k, plan, bkt_keys, bkt_cursor, ctr,
                                      overflow_keys, overflow_cap, grid, 1u, st)));
}

// one-word Pass B with the keys staged through per-warp TMA bulk copies: opt-in (PBK_PASSB_STAGED=1).  Measured on one B200
// (profiles/r2b_tune_variants.jsonl) it is SLOWER than the threads' own streaming loads when the bucket store is local
// (Pass B 4.47 vs 3.71 ms on C1: the pass is bound by its atomics, and the per-round mbarrier waits only add latency).
// Needs even seg_cap (16-byte aligned copies).
static bool passb1_staged(u64 seg_cap)
{
    static const bool on = getenv("PBK_PASSB_STAGED") && atoi(getenv("PBK_PASSB_STAGED")) != 0;
    return on && (seg_cap % 2 == 0);
}
template <typename K> static void passb1_staged_attr(K kernel)
{
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PASSB1_RING_BYTES);
}

size_t passb_desc_bytes(u32 n_buckets) { return 16 + (size_t)(n_buckets + 1) * sizeof(PassBBucket); }

static void prefetch_region(const TableView &t, u32 pb, u32 n_buckets, const char **base, u32 *lines)
{
    *base = nullptr; *lines = 0;
    if (!t.slots || pb >= n_buckets) return;
    const unsigned __int128 cap = t.cap;
    const u64 s0 = (u64)((cap * pb) / n_buckets);
    u64 s1 = (u64)((cap * (pb + 1)) / n_buckets) + 128;
    if (s1 > t.cap) s1 = t.cap;
    *lines = (u32)(((s1 - s0) * t.slot_bytes() + 127) / 128);
    *base = (const char *)t.slots + s0 * t.slot_bytes();
}

void launch_bucket_insert(const u64 *bkt_keys, u64 seg_cap, const u64 *counts, void *h_desc, void *d_desc,
                          u32 b_first, u32 b_end, u32 n_buckets, TableView table, TableView remote, ShardInfo shard,
                          Counters *ctr, u64 *overflow_keys, u64 overflow_cap, int sm_count, cudaStream_t st)
{
    if (b_end <= b_first) return;
    const int pf_dist = getenv("PBK_PF_DIST") ? atoi(getenv("PBK_PF_DIST")) : 1;   // buckets of look-ahead
    // layout of the descriptor buffer: [ticket (u64, padded to 16 bytes)][PassBBucket x (nb + 1)]
    u64 *h_ticket = (u64 *)h_desc;
    PassBBucket *h = (PassBBucket *)((char *)h_desc + 16);
    h_ticket[0] = 0; h_ticket[1] = 0;
    const u64 tile_keys = table.words == 1 ? PASSB1_TILE_KEYS : PASSB_TILE_KEYS;
    u64 tiles = 0;
    for (u32 b = b_first; b < b_end; ++b) {
        PassBBucket &d = h[b - b_first];
        d.tile_start = tiles;
        d.n_keys = counts[b];
        tiles += (counts[b] + tile_keys - 1) / tile_keys;
        d.pf_base = d.pf_base2 = nullptr; d.pf_lines = d.pf_lines2 = 0;
        if (pf_dist > 0 && counts[b]) {
            prefetch_region(table, b + pf_dist, n_buckets, &d.pf_base, &d.pf_lines);
            if (shard.n_shards > 1) prefetch_region(remote, b + pf_dist, n_buckets, &d.pf_base2, &d.pf_lines2);
        }
    }
    h[b_end - b_first] = PassBBucket{tiles, 0, nullptr, nullptr, 0, 0};
    cudaMemcpyAsync(d_desc, h_desc, passb_desc_bytes(b_end - b_first), cudaMemcpyHostToDevice, st);
    if (tiles == 0) return;
    const PassBBucket *d_bk = (const PassBBucket *)((const char *)d_desc + 16);
    if (table.words == 1) {
        // persistent: every CTA resident (3 per SM)
        const int ctas = getenv("PBK_PASSB_CTAS") ? atoi(getenv("PBK_PASSB_CTAS")) : 3;
        const u32 opts = getenv("PBK_PASSB_HINT") ? (u32)atoi(getenv("PBK_PASSB_HINT")) : 1u;
        const int grid = (int)std::min<u64>(tiles, (u64)sm_count * ctas);
        if (passb1_staged(seg_cap)) {
            if (shard.n_shards > 1) {
                passb1_staged_attr(bucket_insert_compact_staged_kernel<1>);
                bucket_insert_compact_staged_kernel<1><<<grid, PASSB_THREADS, PASSB1_RING_BYTES, st>>>(bkt_keys, seg_cap, d_bk, b_first, b_end,
                    (u64 *)d_desc, Table<1>(table.slots, table.cap), Table<1>(remote.slots, remote.cap), shard.n_shards,
                    shard.rank, ctr, overflow_keys, overflow_cap, opts);
            } else {
                passb1_staged_attr(bucket_insert_compact_staged_kernel<0>);
                bucket_insert_compact_staged_kernel<0><<<grid, PASSB_THREADS, PASSB1_RING_BYTES, st>>>(bkt_keys, seg_cap, d_bk, b_first, b_end,
                    (u64 *)d_desc, Table<1>(table.slots, table.cap), Table<1>(remote.slots, remote.cap), 1, 0, ctr,
                    overflow_keys, overflow_cap, opts);
            }
            return;
        }
        if (shard.n_shards > 1)
            bucket_insert_compact_kernel<1><<<grid, PASSB_THREADS, 0, st>>>(bkt_keys, seg_cap, d_bk, b_first, b_end,
                (u64 *)d_desc, Table<1>(table.slots, table.cap), Table<1>(remote.slots, remote.cap), shard.n_shards,
                shard.rank, ctr, overflow_keys, overflow_cap, opts);
        else
            bucket_insert_compact_kernel<0><<<grid, PASSB_THREADS, 0, st>>>(bkt_keys, seg_cap, d_bk, b_first, b_end,
                (u64 *)d_desc, Table<1>(table.slots, table.cap), Table<1>(remote.slots, remote.cap), 1, 0, ctr,
                overflow_keys, overflow_cap, opts);
        return;
    }
    if (getenv("PBK_WIDE_SERIAL")) {                             // the one-key-at-a-time kernel
        const int grid = (int)std::min<u64>(tiles, (u64)sm_count * PASSB_WIDE_CTAS);
        switch (table.words) {
#define PBK_CASE_W(Wv) case Wv: bucket_insert_kernel<Wv><<<grid, PASSB_THREADS, 0, st>>>(bkt_keys, seg_cap, d_bk, b_first,   \
                b_end, (u64 *)d_desc, Table<Wv>(table.slots, table.cap), Table<Wv>(remote.slots, remote.cap), shard.n_shards,    \
                shard.rank, ctr, overflow_keys, overflow_cap); break;
        PBK_CASE_W(2) PBK_CASE_W(3) PBK_CASE_W(4) PBK_CASE_W(5) PBK_CASE_W(6) PBK_CASE_W(7) PBK_CASE_W(8)
#undef PBK_CASE_W
        default: break;
        }
        return;
    }
    const int grid = (int)std::min<u64>(tiles, (u64)sm_count * PBK_PASSBW_MINCTAS);
    // keys staged through TMA bulk copies (two tiles in shared memory): opt-in, PBK_WIDE_STAGED=1 (measured 6 % slower than the
    // threads' own loads at k = 75, profiles/r2b_tune_variants.jsonl); needs even seg_cap for odd W
    static const bool staged_on = getenv("PBK_WIDE_STAGED") && atoi(getenv("PBK_WIDE_STAGED")) != 0;
    const size_t stage_bytes = 2 * (size_t)PASSB_TILE_KEYS * table.words * 8;
    const bool staged = staged_on && stage_bytes * PBK_PASSBW_MINCTAS <= 200 * 1024 && ((seg_cap * table.words) % 2 == 0);
    switch (table.words) {
#define PBK_CASE_W(Wv) case Wv:                                                                                                   \
        if (staged) {                                                                                                          \
            cudaFuncSetAttribute(bucket_insert_wide_kernel<Wv, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes);   \
            bucket_insert_wide_kernel<Wv, true><<<grid, PASSB_THREADS, stage_bytes, st>>>(bkt_keys, seg_cap, d_bk, b_first,       \
                b_end, (u64 *)d_desc, Table<Wv>(table.slots, table.cap), Table<Wv>(remote.slots, remote.cap), shard.n_shards,  \
                shard.rank, ctr, overflow_keys, overflow_cap);                                                                 \
        } else {                                                                                                               \
            bucket_insert_wide_kernel<Wv, false><<<grid, PASSB_THREADS, 0, st>>>(bkt_keys, seg_cap, d_bk, b_first,               \
                b_end, (u64 *)d_desc, Table<Wv>(table.slots, table.cap), Table<Wv>(remote.slots, remote.cap), shard.n_shards,  \
                shard.rank, ctr, overflow_keys, overflow_cap);                                                                 \
        } break;
    PBK_CASE_W(2) PBK_CASE_W(3) PBK_CASE_W(4) PBK_CASE_W(5) PBK_CASE_W(6) PBK_CASE_W(7) PBK_CASE_W(8)
#undef PBK_CASE_W
    default: break;
    }
}

// Key exchange (k <= 32): Pass B over the keys every source holds for this shard (KeyxSources: slices of the all-to-all
// receive buffer, or the peers' bucket stores read in place over NVLink).  `counts` are in descriptor order (descriptor
// i = region i / n_src, source i % n_src); descriptors [d_first, d_end) of n_src x n_regions are inserted by this launch.
// While (region j, source s) is worked on, slice s of region j + 1 is prefetched.
void launch_bucket_insert_gathered(const KeyxSources &srcs, u64 seg_cap, const u64 *counts, void *h_desc, void *d_desc,
                                   u32 d_first, u32 d_end, u32 n_src, u32 n_regions, TableView table, Counters *ctr,
                                   u64 *overflow_keys, u64 overflow_cap, int sm_count, cudaStream_t st)
{
    if (d_end <= d_first || table.words != 1) return;
    const int pf_dist = getenv("PBK_PF_DIST") ? atoi(getenv("PBK_PF_DIST")) : 1;
    const u32 n_desc = n_src * n_regions;
    u64 *h_ticket = (u64 *)h_desc;
    PassBBucket *h = (PassBBucket *)((char *)h_desc + 16);
    h_ticket[0] = 0; h_ticket[1] = 0;
    u64 tiles = 0;
    for (u32 i = d_first; i < d_end; ++i) {
        PassBBucket &d = h[i - d_first];
        d.tile_start = tiles;
        d.n_keys = counts[i];
        tiles += (counts[i] + PASSB1_TILE_KEYS - 1) / PASSB1_TILE_KEYS;
        d.pf_base = d.pf_base2 = nullptr; d.pf_lines = d.pf_lines2 = 0;
        if (pf_dist > 0 && counts[i]) prefetch_region(table, i + (u32)pf_dist * n_src, n_desc, &d.pf_base, &d.pf_lines);
    }
    h[d_end - d_first] = PassBBucket{tiles, 0, nullptr, nullptr, 0, 0};
    cudaMemcpyAsync(d_desc, h_desc, passb_desc_bytes(d_end - d_first), cudaMemcpyHostToDevice, st);
    if (tiles == 0) return;
    const PassBBucket *d_bk = (const PassBBucket *)((const char *)d_desc + 16);
    const int ctas = getenv("PBK_PASSB_CTAS") ? atoi(getenv("PBK_PASSB_CTAS")) : 3;
    const u32 opts = (getenv("PBK_PASSB_HINT") ? ((u32)atoi(getenv("PBK_PASSB_HINT")) & 0xFFu) : 1u) | (n_src << 8) | (n_regions << 16);
    const int grid = (int)std::min<u64>(tiles, (u64)sm_count * ctas);
    if (passb1_staged(seg_cap)) {
        passb1_staged_attr(bucket_insert_gather_staged_kernel);
        bucket_insert_gather_staged_kernel<<<grid, PASSB_THREADS, PASSB1_RING_BYTES, st>>>(srcs, seg_cap, d_bk, d_first, d_end, (u64 *)d_desc,
            Table<1>(table.slots, table.cap), ctr, overflow_keys, overflow_cap, opts);
        return;
    }
    bucket_insert_gather_kernel<<<grid, PASSB_THREADS, 0, st>>>(srcs, seg_cap, d_bk, d_first, d_end, (u64 *)d_desc,
        Table<1>(table.slots, table.cap), ctr, overflow_keys, overflow_cap, opts);
}

// the same without a host round trip: tile map built on the device from the sources' cursors, Pass B right behind it
void launch_bucket_insert_gathered_chained(const KeyxSources &srcs, u64 seg_cap, void *d_desc, u32 n_src,
                                           u32 n_regions, TableView table, Counters *ctr, u64 *overflow_keys, u64 overflow_cap,
                                           int sm_count, cudaStream_t st)
{
    if (table.words != 1) return;
    const int pf_dist = getenv("PBK_PF_DIST") ? atoi(getenv("PBK_PF_DIST")) : 1;
    const u32 n_desc = n_src * n_regions;
    passb_desc_gather_kernel<<<1, PART_MAX_BUCKETS, 0, st>>>(srcs, seg_cap, n_src, n_regions, (u32)PASSB1_TILE_KEYS,
        (const char *)table.slots, table.cap, (u32)table.slot_bytes(), pf_dist, (u64 *)d_desc, (PassBBucket *)((char *)d_desc + 16));
    const PassBBucket *d_bk = (const PassBBucket *)((const char *)d_desc + 16);
    const int ctas = getenv("PBK_PASSB_CTAS") ? atoi(getenv("PBK_PASSB_CTAS")) : 3;
    const u32 opts = (getenv("PBK_PASSB_HINT") ? ((u32)atoi(getenv("PBK_PASSB_HINT")) & 0xFFu) : 1u) | (n_src << 8) | (n_regions << 16);
    if (passb1_staged(seg_cap)) {
        passb1_staged_attr(bucket_insert_gather_staged_kernel);
        bucket_insert_gather_staged_kernel<<<sm_count * ctas, PASSB_THREADS, PASSB1_RING_BYTES, st>>>(srcs, seg_cap, d_bk, 0, n_desc, (u64 *)d_desc,
            Table<1>(table.slots, table.cap), ctr, overflow_keys, overflow_cap, opts);
        return;
    }
    bucket_insert_gather_kernel<<<sm_count * ctas, PASSB_THREADS, 0, st>>>(srcs, seg_cap, d_bk, 0, n_desc, (u64 *)d_desc,
        Table<1>(table.slots, table.cap), ctr, overflow_keys, overflow_cap, opts);    // CTAs that find no tile left leave at once
}

// device-chained variant for k <= 32: tile map built by a kernel from the cursors, Pass B launched on `st_insert`
// right behind it (the caller orders the two streams with events)
void launch_passb_desc(const u64 *d_cursor, u64 seg_cap, u32 n_buckets, TableView table, TableView remote, ShardInfo shard,
                       void *d_desc, cudaStream_t st)
{
    const int pf_dist = getenv("PBK_PF_DIST") ? atoi(getenv("PBK_PF_DIST")) : 1;
    const u32 tile_keys = table.words == 1 ? (u32)PASSB1_TILE_KEYS : (u32)PASSB_TILE_KEYS;
    passb_desc_kernel<<<1, PART_MAX_BUCKETS, 0, st>>>(d_cursor, seg_cap, n_buckets, tile_keys, (const char *)table.slots,
        table.cap, shard.n_shards > 1 ? (const char *)remote.slots : nullptr, remote.cap, (u32)table.slot_bytes(), pf_dist,
        (u64 *)d_desc, (PassBBucket *)((char *)d_desc + 16));
}

void launch_bucket_insert_chained(const u64 *bkt_keys, u64 seg_cap, const void *d_desc, u32 n_buckets, TableView table,
                                  TableView remote, ShardInfo shard, Counters *ctr, u64 *overflow_keys, u64 overflow_cap,
                                  int sm_count, int ctas_per_sm, cudaStream_t st)
{
    const PassBBucket *d_bk = (const PassBBucket *)((const char *)d_desc + 16);
    if (table.words > 1) {                                      // multi-word keys: the same chaining, the wide kernel (its own streaming loads)
        const int wgrid = sm_count * PBK_PASSBW_MINCTAS;
        switch (table.words) {
#define PBK_CASE_W(Wv) case Wv: bucket_insert_wide_kernel<Wv, false><<<wgrid, PASSB_THREADS, 0, st>>>(bkt_keys, seg_cap, d_bk, 0, n_buckets, \
                (u64 *)d_desc, Table<Wv>(table.slots, table.cap), Table<Wv>(remote.slots, remote.cap), shard.n_shards, shard.rank, ctr,     \
                overflow_keys, overflow_cap); break;
        PBK_CASE_W(2) PBK_CASE_W(3) PBK_CASE_W(4) PBK_CASE_W(5) PBK_CASE_W(6) PBK_CASE_W(7) PBK_CASE_W(8)
#undef PBK_CASE_W
        default: break;
        }
        return;
    }
    const u32 opts = getenv("PBK_PASSB_HINT") ? (u32)atoi(getenv("PBK_PASSB_HINT")) : 1u;
    if (getenv("PBK_PASSB_CTAS")) ctas_per_sm = atoi(getenv("PBK_PASSB_CTAS"));
    const int grid = sm_count * ctas_per_sm;                    // CTAs that find no tile left leave at once
    if (passb1_staged(seg_cap)) {
        if (shard.n_shards > 1) {
            passb1_staged_attr(bucket_insert_compact_staged_kernel<1>);
            bucket_insert_compact_staged_kernel<1><<<grid, PASSB_THREADS, PASSB1_RING_BYTES, st>>>(bkt_keys, seg_cap, d_bk, 0, n_buckets, (u64 *)d_desc,
                Table<1>(table.slots, table.cap), Table<1>(remote.slots, remote.cap), shard.n_shards, shard.rank, ctr,
                overflow_keys, overflow_cap, opts);
        } else {
            passb1_staged_attr(bucket_insert_compact_staged_kernel<0>);
            bucket_insert_compact_staged_kernel<0><<<grid, PASSB_THREADS, PASSB1_RING_BYTES, st>>>(bkt_keys, seg_cap, d_bk, 0, n_buckets, (u64 *)d_desc,
                Table<1>(table.slots, table.cap), Table<1>(remote.slots, remote.cap), 1, 0, ctr, overflow_keys, overflow_cap, opts);
        }
        return;
    }
    if (shard.n_shards > 1)
        bucket_insert_compact_kernel<1><<<grid, PASSB_THREADS, 0, st>>>(bkt_keys, seg_cap, d_bk, 0, n_buckets, (u64 *)d_desc,
            Table<1>(table.slots, table.cap), Table<1>(remote.slots, remote.cap), shard.n_shards, shard.rank, ctr,
            overflow_keys, overflow_cap, opts);
    else
        bucket_insert_compact_kernel<0><<<grid, PASSB_THREADS, 0, st>>>(bkt_keys, seg_cap, d_bk, 0, n_buckets, (u64 *)d_desc,
            Table<1>(table.slots, table.cap), Table<1>(remote.slots, remote.cap), 1, 0, ctr, overflow_keys, overflow_cap, opts);
}

// ---- Pass B, second form (k <= 32, unsharded): split_kernel + region_build_kernel (see pbk_kernels_impl.cuh) ----------------
// The route needs Pass A's buckets to be whole groups of sub-regions: a power-of-two bucket count (its bucket function is then
// the top hash bits) and a table of at least one sub-region per bucket.
bool passb2_geom(TableView table, u32 n_buckets, Passb2Geom *out)
{
    if (table.words != 1 || !table.slots || n_buckets == 0 || (n_buckets & (n_buckets - 1))) return false;
    if ((table.cap & (table.cap - 1)) || table.cap > (1ull << 31)) return false;   // (region_build_kernel looks at 32-bit halves of a slot)
    const u64 n_sub = table.cap >> BUILD_LOG2_SLOTS;
    if (n_sub < n_buckets || n_sub / n_buckets > (u64)SPLIT_MAX_F) return false;
    out->n_sub = n_sub;
    out->F = (u32)(n_sub / n_buckets);
    out->sub_shift = ct_geom(table.cap).rbits + BUILD_LOG2_SLOTS;
    return true;
}

// entries per sub-region segment for a bucket store of up to `windows_ub` keys: the mean plus a half (a sub-region holds a few
// thousand distinct keys, so its share of a read set's instances varies by several per cent; keys beyond a full segment go
// through the overflow list), even (16-byte aligned segments)
u64 passb2_sub_cap(u64 windows_ub, u64 n_sub)
{
    const u64 mean = windows_ub / n_sub + 1;
    return (mean + mean / 2 + 1024 + 1) & ~1ull;
}

// CTA size of split_kernel: 512 threads (tiles of 8192 keys, two CTAs per SM) unless PBK_SPLIT_THREADS says 256 (tiles of 4096
// keys, four CTAs per SM: smaller barrier domains, shorter runs per sub-region) -- a tuning switch
static int split_threads()
{
#ifdef PBK_CPU_EMUL
    return SPLIT_LAUNCH_THREADS;
#else
    static const int t = (getenv("PBK_SPLIT_THREADS") && atoi(getenv("PBK_SPLIT_THREADS")) == 256) ? 256 : SPLIT_THREADS;
    return t;
#endif
}
static u32 split_tile_keys() { return (u32)split_threads() * SPLIT_KPT; }

// tile map of ALL n_buckets buckets for split_kernel, built on the device from the cursors
void launch_passb2_desc(const u64 *d_cursor, u64 seg_cap, u32 n_buckets, void *d_desc, cudaStream_t st)
{
    passb_desc_kernel<<<1, PART_MAX_BUCKETS, 0, st>>>(d_cursor, seg_cap, n_buckets, split_tile_keys(), nullptr, 0, nullptr, 0,
        8u, 0, (u64 *)d_desc, (PassBBucket *)((char *)d_desc + 16));
}
// the same for the key exchange: descriptor i = (table region i / n_src, source i % n_src), fill counts read from the sources
void launch_passb2_desc_gather(const KeyxSources &srcs, u64 seg_cap, u32 n_src, u32 n_regions, void *d_desc, cudaStream_t st)
{
    passb_desc_gather_kernel<<<1, PART_MAX_BUCKETS, 0, st>>>(srcs, seg_cap, n_src, n_regions, split_tile_keys(), nullptr, 0, 8u, 0,
        (u64 *)d_desc, (PassBBucket *)((char *)d_desc + 16));
}

// descriptors [d_first, d_end) of the tile map; srcs == nullptr: the context's own bucket store (descriptor = bucket)
void launch_passb2_split(const u64 *bkt_keys, const KeyxSources *srcs, u32 n_src, u64 seg_cap, const void *d_desc, u32 d_first, u32 d_end,
                         const Passb2Geom &geom, u64 *d_sub_keys, u64 sub_cap, u64 *d_sub_cursor, Counters *ctr, u64 *overflow_keys,
                         u64 overflow_cap, int sm_count, cudaStream_t st)
{
    const PassBBucket *d_bk = (const PassBBucket *)((const char *)d_desc + 16);
    const size_t smem = (size_t)split_tile_keys() * 8;
    const int threads = split_threads();
    const int ctas = getenv("PBK_SPLIT_CTAS") ? std::max(1, atoi(getenv("PBK_SPLIT_CTAS"))) : (threads == 256 ? 4 : 2);
    if (srcs) {
        cudaFuncSetAttribute(split_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        split_gather_kernel<<<sm_count * ctas, threads, smem, st>>>(*srcs, n_src, seg_cap, d_bk, d_first, d_end, geom.F, geom.sub_shift,
            d_sub_keys, sub_cap, d_sub_cursor, ctr, overflow_keys, overflow_cap);
        return;
    }
    cudaFuncSetAttribute(split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    split_kernel<<<sm_count * ctas, threads, smem, st>>>(bkt_keys, seg_cap, d_bk, d_first, d_end, geom.F, geom.sub_shift,
        d_sub_keys, sub_cap, d_sub_cursor, ctr, overflow_keys, overflow_cap);
}

// buckets (table regions) [b_first, b_end)
void launch_passb2_build(const u64 *d_sub_keys, u64 sub_cap, const u64 *d_sub_cursor, u32 b_first, u32 b_end, const Passb2Geom &geom,
                         TableView table, bool load_existing, Counters *ctr, u64 *overflow_keys, u64 overflow_cap, int sm_count,
                         cudaStream_t st)
{
    const size_t smem = BUILD_SMEM_BYTES;
    cudaFuncSetAttribute(region_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int ctas = getenv("PBK_BUILD_CTAS") ? std::max(1, atoi(getenv("PBK_BUILD_CTAS"))) : 2;
    region_build_kernel<<<sm_count * ctas, BUILD_THREADS, smem, st>>>(d_sub_keys, sub_cap, d_sub_cursor, (u64)b_first * geom.F,
        (u64)b_end * geom.F, geom.sub_shift, Table<1>(table.slots, table.cap), load_existing ? 1 : 0, ctr, overflow_keys, overflow_cap);
}

void launch_table_init(TableView t, cudaStream_t st)
{
    cudaMemsetAsync(t.slots, 0, t.bytes(), st);        // both slot formats use all-zero for "empty"
}

void launch_table_rehash(TableView from, TableView to, Counters *ctr, cudaStream_t st)
{
    const int grid = grid_for(from.capacity(), 256, 148, 8);
    PBK_DISPATCH_W(from.words,
        (rehash_kernel<W><<<grid, 256, 0, st>>>(Table<W>(from.slots, from.cap), Table<W>(to.slots, to.cap), ctr)));
}

void launch_table_clamp(TableView t, cudaStream_t st)
{
    const int grid = grid_for(t.capacity(), 256, 148, 8);
    PBK_DISPATCH_W(t.words, (clamp_kernel<W><<<grid, 256, 0, st>>>(Table<W>(t.slots, t.cap))));
}

void launch_table_histogram(TableView t, u64 *occ_hist, cudaStream_t st)
{
    const int grid = grid_for(t.capacity() / 8 + 1, 256, 148, 8);      // 8 slots per thread and iteration, 8 resident CTAs per SM
    PBK_DISPATCH_W(t.words, (histogram_kernel<W><<<grid, 256, 0, st>>>(Table<W>(t.slots, t.cap), occ_hist)));
}

void launch_table_export(TableView t, u32 min_count, u64 *keys_out, uint16_t *counts_out, u64 capacity,
                         u64 *d_n_out, cudaStream_t st)
{
    const int grid = grid_for(t.capacity(), 256, 148, 8);
    PBK_DISPATCH_W(t.words,
        (export_kernel<W><<<grid, 256, 0, st>>>(Table<W>(t.slots, t.cap), min_count, keys_out,
            counts_out, capacity, d_n_out)));
}

void launch_shard_count(TableView remote, u32 n_shards, u64 *d_counts, cudaStream_t st)
{
    const int grid = grid_for(remote.capacity(), 256, 148, 8);
    PBK_DISPATCH_W(remote.words,
        (shard_count_kernel<W><<<grid, 256, 0, st>>>(Table<W>(remote.slots, remote.cap), n_shards, d_counts)));
}

void launch_shard_pack(TableView remote, u32 n_shards, u64 *d_cursors, u64 *records_out, cudaStream_t st)
{
    const int grid = grid_for(remote.capacity(), 256, 148, 8);
    PBK_DISPATCH_W(remote.words,
        (shard_pack_kernel<W><<<grid, 256, 0, st>>>(Table<W>(remote.slots, remote.cap), n_shards,
            d_cursors, records_out)));
}

// =================================================================================================
// sorted export (sortedKeyFromKmerFile, counter.h:917-951): LSD radix sort, least significant word
// first, with a permutation carried along for multi-word keys.  Radix sort itself is library code
// (CUB); it is not on the counting hot path.
// =================================================================================================

#ifndef PBK_CPU_EMUL
__global__ void iota_kernel(u32 *p, u64 n)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) p[i] = (u32)i;
}
__global__ void gather_word_kernel(const u64 *keys, const u32 *perm, u64 n, int words, int w, u64 *out)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
        out[i] = keys[(u64)perm[i] * words + w];
}
__global__ void gather_rows_kernel(const u64 *keys, const uint16_t *counts, const u32 *perm, u64 n, int words,
                                   u64 *keys_out, uint16_t *counts_out)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const u64 src = perm[i];
        for (int j = 0; j < words; ++j) keys_out[i * words + j] = keys[src * words + j];
        counts_out[i] = counts[src];
    }
}

namespace {
inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }
struct SortCarve {                                  // where each temporary lives inside the caller's scratch block
    size_t k2, c2, perm, perm2, kw, kw2, keys_tmp, counts_tmp, cub, cub_bytes, total;
};
SortCarve sort_carve(u64 n, int words)
{
    SortCarve c{};
    size_t at = 0;
    auto take = [&](size_t bytes) { const size_t o = at; at += al256(bytes); return o; };
    if (words == 1) {
        c.k2 = take(n * 8); c.c2 = take(n * 2);
        cub::DoubleBuffer<u64> dk(nullptr, nullptr);
        cub::DoubleBuffer<uint16_t> dc(nullptr, nullptr);
        cub::DeviceRadixSort::SortPairs(nullptr, c.cub_bytes, dk, dc, (long long)n, 0, 64, (cudaStream_t)0);
    } else {
        c.perm = take(n * 4); c.perm2 = take(n * 4); c.kw = take(n * 8); c.kw2 = take(n * 8);
        c.keys_tmp = take(n * 8 * words); c.counts_tmp = take(n * 2);
        cub::DoubleBuffer<u64> dk(nullptr, nullptr);
        cub::DoubleBuffer<u32> dp(nullptr, nullptr);
        cub::DeviceRadixSort::SortPairs(nullptr, c.cub_bytes, dk, dp, (long long)n, 0, 64, (cudaStream_t)0);
    }
    c.cub = take(c.cub_bytes ? c.cub_bytes : 1);
    c.total = at;
    return c;
}
}  // namespace

size_t sort_export_scratch_bytes(u64 n, int words) { return n < 2 ? 0 : sort_carve(n, words).total; }

// Everything is queued on `st`; no allocation, no synchronisation (the scratch block belongs to the context's HBM budget).
cudaError_t sort_export(u64 *keys, uint16_t *counts, u64 n, int words, int k, void *scratch, size_t scratch_bytes, cudaStream_t st)
{
    if (n < 2) return cudaSuccess;
    if (n >= (1ull << 32)) return cudaErrorInvalidValue;        // the permutation of the multi-word path is 32-bit
    const SortCarve c = sort_carve(n, words);
    if (!scratch || scratch_bytes < c.total) return cudaErrorInvalidValue;
    char *base = (char *)scratch;
    size_t tmp_bytes = c.cub_bytes;
    cudaError_t e;
    const int grid = grid_for(n, 256, 148, 8);
    const long long items = (long long)n;                        // 64-bit item count: n may exceed INT_MAX
    if (words == 1) {
        cub::DoubleBuffer<u64> dk(keys, (u64 *)(base + c.k2));
        cub::DoubleBuffer<uint16_t> dc(counts, (uint16_t *)(base + c.c2));
        const int end_bit = 2 * k > 64 ? 64 : 2 * k;
        e = cub::DeviceRadixSort::SortPairs(base + c.cub, tmp_bytes, dk, dc, items, 0, end_bit, st);
        if (!e && dk.Current() != keys) {
            cudaMemcpyAsync(keys, dk.Current(), n * 8, cudaMemcpyDeviceToDevice, st);
            cudaMemcpyAsync(counts, dc.Current(), n * 2, cudaMemcpyDeviceToDevice, st);
        }
        return e;
    }
    u32 *perm = (u32 *)(base + c.perm);
    u64 *keys_tmp = (u64 *)(base + c.keys_tmp);
    uint16_t *counts_tmp = (uint16_t *)(base + c.counts_tmp);
    cub::DoubleBuffer<u64> dk((u64 *)(base + c.kw), (u64 *)(base + c.kw2));
    cub::DoubleBuffer<u32> dp(perm, (u32 *)(base + c.perm2));
    iota_kernel<<<grid, 256, 0, st>>>(perm, n);
    e = cudaGetLastError();
    for (int w = 0; w < words && !e; ++w) {
        gather_word_kernel<<<grid, 256, 0, st>>>(keys, dp.Current(), n, words, w, dk.Current());
        int end_bit = 64;
        if (w == words - 1 && (k & 31)) end_bit = 2 * (k & 31);
        e = cub::DeviceRadixSort::SortPairs(base + c.cub, tmp_bytes, dk, dp, items, 0, end_bit, st);
    }
    if (!e) {
        gather_rows_kernel<<<grid, 256, 0, st>>>(keys, counts, dp.Current(), n, words, keys_tmp, counts_tmp);
        cudaMemcpyAsync(keys, keys_tmp, n * 8 * words, cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(counts, counts_tmp, n * 2, cudaMemcpyDeviceToDevice, st);
    }
    return e;
}

// =================================================================================================
// microbenchmark: random read-modify-write rate (R_atomic of SURVEY.md section 8d)
// =================================================================================================

struct MbSlot { u64 key; u32 cs; u32 pad; };     // 16-byte key + count slot of the first table design

__global__ void __launch_bounds__(256)
microbench_kernel(MbSlot *t, int log2slots, u64 n_ops, int mode, u64 seed)
{
    const u64 stride = (u64)gridDim.x * blockDim.x;
    u32 sink = 0;
    const CtGeom g = ct_geom(2ull << log2slots);      // modes 2-4 see the buffer as 2^(log2slots+1) 8-byte slots
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_ops; i += stride) {
        const u64 h = fmix64(i + seed);
        if (mode == 0) {
            red_add_u32(&t[h >> (64 - log2slots)].cs, 1u);
        } else if (mode == 1) {
            MbSlot *s = t + (h >> (64 - log2slots));
            const u64 cur = ld_cg_u64(&s->key);
            if (cur != 0x123456789ull) red_add_u32(&s->cs, 1u);
            else ++sink;
        } else if (mode == 2) {                      // compact-table inserts of distinct keys (claims)
            sink += (u32)ct_insert_unit(reinterpret_cast<u64 *>(t), g, h);
        } else if (mode == 3) {                      // 8-byte slots: one atomic with return per op
            u64 *s8 = reinterpret_cast<u64 *>(t);
            const u64 old = atomicAdd(&s8[h >> (63 - log2slots)], 1ull);
            sink += (u32)(old >> 40);
        } else if (mode == 5) {                      // mode 3 with C1's skew: 7/8 of the ops on 1/28 of the slots
            u64 *s8 = reinterpret_cast<u64 *>(t);
            const u64 n8 = 2ull << log2slots;
            u64 idx = h >> (63 - log2slots);
            if ((h & 7) != 0) idx = (fmix64(h >> 3) % (n8 / 28)) * 28 + 5;
            const u64 old = atomicAdd(&s8[idx], 1ull);
            sink += (u32)(old >> 40);
        } else if (mode >= 6 && mode <= 8) {
            // Pass B skeleton on an L2-resident table: 6 = key from an HBM stream + atomic; 7 = + result-dependent
            // follow-up (1 op in 8: take the +1 back, second atomic); 8 = 7 with eight keys per thread, atomics first
            u64 *s8 = reinterpret_cast<u64 *>(t);
            const u64 *stream = s8 + (2ull << log2slots);        // the key stream lives behind the table
            const u64 smask = (1ull << 27) - 1;                  // 1 GiB of keys
            if (mode < 8) {
                const u64 k = __ldcs(stream + (i & smask));
                const u64 hk = fmix64(k + i);
                u64 *q = &s8[hk >> (63 - log2slots)];
                const u64 old = atomicAdd(q, 1ull);
                if (mode == 7 && ((old ^ hk) & 7) == 0) {
                    red_add_u64(q, ~0ull);
                    sink += (u32)atomicAdd(&s8[(hk >> (63 - log2slots)) ^ 1], 1ull);
                }
                sink += (u32)(old >> 40);
            } else {
                if ((i / stride) & 7) continue;                    // every 8th iteration does 8 ops
                u64 hk[8], old[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) hk[j] = fmix64(__ldcs(stream + ((i + j * stride) & smask)) + i + j * stride);
#pragma unroll
                for (int j = 0; j < 8; ++j) old[j] = atomicAdd(&s8[hk[j] >> (63 - log2slots)], 1ull);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (((old[j] ^ hk[j]) & 7) == 0) {
                        red_add_u64(&s8[hk[j] >> (63 - log2slots)], ~0ull);
                        sink += (u32)atomicAdd(&s8[(hk[j] >> (63 - log2slots)) ^ 1], 1ull);
                    }
#pragma unroll
                for (int j = 0; j < 8; ++j) sink += (u32)(old[j] >> 40);
            }
        } else if (mode >= 9 && mode <= 11) {
            // mode 7 on a table that does NOT fit in L2: the ops sweep 128 consecutive regions, each op lands
            // in the region its position i belongs to.  9 = no prefetch; 10 = prefetch.global.L2 of the
            // region after the next, sector by sector; 11 = the same with plain loads (ld.cg) instead
            u64 *s8 = reinterpret_cast<u64 *>(t);
            const u64 n8 = 2ull << log2slots, rs = n8 >> 7;      // slots per region
            const u64 *stream = s8 + n8;
            const u64 smask = (1ull << 27) - 1;
            const u64 per_region = (n_ops + 127) >> 7;
            const u64 region = i / per_region, o = i - region * per_region;
            if (mode >= 10 && region + 2 < 128) {
                const u64 sectors = rs * 8 / 32;
                const u64 a = (u64)((double)o * (double)sectors / (double)per_region), b = (u64)((double)(o + 1) * (double)sectors / (double)per_region);
                if (b > a) {
                    const char *pa = (const char *)(s8 + (region + 2) * rs) + a * 32;
                    if (mode == 10) asm volatile("prefetch.global.L2 [%0];" :: "l"(pa));
                    else sink += (u32)ld_cg_u32((const u32 *)pa);
                }
            }
            const u64 k = __ldcs(stream + (i & smask));
            const u64 hk = fmix64(k + i);
            u64 *q = &s8[region * rs + (hk % rs)];
            const u64 old = atomicAdd(q, 1ull);
            if (((old ^ hk) & 7) == 0) {
                red_add_u64(q, ~0ull);
                sink += (u32)atomicAdd(&s8[region * rs + ((hk % rs) ^ 1)], 1ull);
            }
            sink += (u32)(old >> 40);
        } else if (mode >= 100) {
            // sweep with shift/mask index math only: mode = 100 + 4 * log2(regions) + variant.  The table is cut
            // into 2^rl consecutive regions, op i lands in region i >> log2(n_ops / regions) (n_ops must be a
            // power of two).  variant 0 = atomic with return; 1 = + prefetch.L2 of the next region, one per sector;
            // 2 = atomic + streamed 8-byte key load; 3 = 2 + claim-style follow-up reduction on 1 op in 8
            u64 *s8 = reinterpret_cast<u64 *>(t);
            const int rl = (mode - 100) >> 2, variant = (mode - 100) & 3;
            const u64 n8 = 2ull << log2slots;
            const int rs_log = log2slots + 1 - rl;                 // slots per region
            const int pr_log = 63 - __clzll((long long)n_ops) - rl;   // ops per region
            const u64 region = i >> pr_log, o = i & ((1ull << pr_log) - 1);
            if (variant == 1 && region + 1 < (1ull << rl)) {
                const int sec_log = rs_log - 2;                    // sectors per region (4 slots per sector)
                if (sec_log >= pr_log) {
                    for (u64 q = o << (sec_log - pr_log); q < ((o + 1) << (sec_log - pr_log)); ++q)
                        asm volatile("prefetch.global.L2 [%0];" :: "l"((const char *)(s8 + ((region + 1) << rs_log)) + q * 32));
                } else if ((o & ((1ull << (pr_log - sec_log)) - 1)) == 0) {
                    asm volatile("prefetch.global.L2 [%0];" :: "l"((const char *)(s8 + ((region + 1) << rs_log)) + (o >> (pr_log - sec_log)) * 32));
                }
            }
            u64 hk = h;
            if (variant >= 2) hk = fmix64(__ldcs(s8 + n8 + (i & ((1ull << 27) - 1))) + i);
            u64 *q = &s8[(region << rs_log) + (hk >> (64 - rs_log))];
            const u64 old = atomicAdd(q, 1ull);
            if (variant == 3 && ((old ^ hk) & 7) == 0) red_add_u64(q, 1ull << 40);
            sink += (u32)(old >> 40);
        } else {                                     // 8-byte slots: load, then reduction
            u64 *s8 = reinterpret_cast<u64 *>(t);
            u64 *q = &s8[h >> (63 - log2slots)];
            const u64 cur = ld_cg_u64(q);
            if (cur != 0x123456789ull) red_add_u64(q, 1ull);
            else ++sink;
        }
    }
    if (sink == 0xFFFFFFFFu) t[0].pad = sink;
}

void launch_microbench(void *table, int log2slots, u64 n_ops, int mode, u64 seed, int sm_count, cudaStream_t st)
{
    microbench_kernel<<<grid_for(n_ops, 256, sm_count, 8), 256, 0, st>>>((MbSlot *)table, log2slots, n_ops, mode, seed);
}

#endif  // PBK_CPU_EMUL

}  // namespace pbk
