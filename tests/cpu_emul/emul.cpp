// emul.cpp -- TEST ONLY: drives the product kernels (pbk_kernels_impl.cuh) sequentially on the host
// through the same stages pbk_api.cu launches on the GPU: read_marks -> pack -> count -> histogram ->
// export.  Used by tests/test_kernel_logic_cpu.py to compare the kernel logic with the oracle.
#include "cuda_shim.h"
#include "../../platanus_b_b200/csrc/pbk_kernels_impl.cuh"

#include <vector>

using namespace pbk;

static int g_partition = 0;
extern "C" void emul_set_partition(int on) { g_partition = on; }

template <int W>
static int run(const uint8_t *bases, const u64 *off, u64 n_reads, int k, int encoding, const int32_t *n_pos,
               const u64 *n_pos_off, u64 table_slots, u32 n_shards, u32 rank, u32 min_count,
               u64 *keys_out, uint16_t *counts_out, u64 cap_out, u64 *n_out, u64 *occ_hist, u64 *len_hist,
               u64 *n_inst, u32 *err_flags, u64 *remote_records, u64 *n_remote)
{
    const u64 n_bases = off[n_reads], words = (n_bases + 31) / 32;
    std::vector<u64> stream(words + STREAM_PAD_WORDS + 1, 0);
    std::vector<u32> nflag(words + STREAM_PAD_WORDS + 1, 0), rflag(words + STREAM_PAD_WORDS + 1, 0);
    Counters ctr{};
    typedef typename SlotType<W>::type slot_t;
    if (W == 1) {                                  // compact slots: power-of-two capacity, >= 2^24 here (17-bit counts)
        u64 p2 = 1ull << 24;
        while (p2 < table_slots) p2 <<= 1;
        if (table_slots < (1ull << 20)) p2 = 1ull << 24;
        table_slots = p2;
    }
    std::vector<slot_t> table_v(table_slots), remote_v(n_shards > 1 ? table_slots : 1);
    memset(table_v.data(), 0, table_v.size() * sizeof(slot_t));
    memset(remote_v.data(), 0, remote_v.size() * sizeof(slot_t));
    Table<W> table(table_v.data(), table_v.size()), remote(remote_v.data(), remote_v.size());
    const u64 OVF = 1 << 16;
    std::vector<u64> ovf((W + 1) * OVF);

    read_marks_kernel(off, n_reads, len_hist, rflag.data() + STREAM_PAD_WORDS, &ctr);
    // two chunks with an odd split to exercise word0 offsets (chunk boundaries are multiples of 32 bases)
    const u64 w_split = words / 3;
    pack_kernel<false>(bases, std::min<u64>(n_bases, w_split * 32), w_split, encoding, stream.data() + STREAM_PAD_WORDS,
                       nflag.data() + STREAM_PAD_WORDS, 0, &ctr);
    pack_kernel<false>(bases + w_split * 32, n_bases - w_split * 32, words - w_split, encoding,
                       stream.data() + STREAM_PAD_WORDS, nflag.data() + STREAM_PAD_WORDS, w_split, &ctr);
    if (encoding == 1) npos_scatter_kernel(off, n_pos, n_pos_off, n_reads, nflag.data() + STREAM_PAD_WORDS);
    if (!g_partition) {
        count_kernel<W>(stream.data() + STREAM_PAD_WORDS, nflag.data() + STREAM_PAD_WORDS, rflag.data() + STREAM_PAD_WORDS,
                        0, w_split, k, table, remote, n_shards, rank, &ctr,
                        ovf.data(), OVF);
        count_kernel<W>(stream.data() + STREAM_PAD_WORDS, nflag.data() + STREAM_PAD_WORDS, rflag.data() + STREAM_PAD_WORDS,
                        w_split, words, k, table, remote, n_shards, rank, &ctr,
                        ovf.data(), OVF);
    } else {
        // Pass A into P buckets with deliberately tiny bins/segments so the direct-append and spill paths run too
        const u32 P = (k & 1) ? 7 : 8, bin_cap = 3;     // odd k: generic bucket function, even k: the power-of-two shift
        const u64 seg_cap = std::max<u64>(4, n_bases / P * 3 / 4);
        std::vector<u64> bkt((u64)P * seg_cap * W, 0), cursor(P, 0);
        partition_kernel<W>(stream.data() + STREAM_PAD_WORDS, nflag.data() + STREAM_PAD_WORDS, rflag.data() + STREAM_PAD_WORDS,
                            0, w_split, k, P, bin_cap, bkt.data(), seg_cap, cursor.data(), &ctr, ovf.data(), OVF, 1u);
        partition_kernel<W>(stream.data() + STREAM_PAD_WORDS, nflag.data() + STREAM_PAD_WORDS, rflag.data() + STREAM_PAD_WORDS,
                            w_split, words, k, P, bin_cap, bkt.data(), seg_cap, cursor.data(), &ctr, ovf.data(), OVF, 1u);
        std::vector<u64> count(P);
        for (u32 b = 0; b < P; ++b) count[b] = std::min<u64>(cursor[b], seg_cap);
        const u64 tk = W == 1 ? PASSB1_KPT * PASSB1_ROUNDS : PASSB_KPT;     // blockDim = 1 in the emulation
        auto pass_b = [&](const PassBBucket *desc, u32 b0, u32 b1, u64 *ticket) {
            if constexpr (W == 1) {
                if (n_shards > 1) bucket_insert_compact_staged_kernel<1>(bkt.data(), seg_cap, desc, b0, b1, ticket, table, remote, n_shards, rank, &ctr, ovf.data(), OVF, 1u);
                else bucket_insert_compact_staged_kernel<0>(bkt.data(), seg_cap, desc, b0, b1, ticket, table, remote, n_shards, rank, &ctr, ovf.data(), OVF, 0u);
            } else if (k % 2) {            // odd k: the batched kernel, even k: the one-key-at-a-time form of the protocol
                bucket_insert_wide_kernel<W, true>(bkt.data(), seg_cap, desc, b0, b1, ticket, table, remote, n_shards, rank, &ctr, ovf.data(), OVF);
            } else {
                bucket_insert_kernel<W>(bkt.data(), seg_cap, desc, b0, b1, ticket, table, remote, n_shards, rank, &ctr, ovf.data(), OVF);
            }
        };
        if (W == 1 && (k % 3) == 0) {
            // chained route of pbk_api.cu (`Pipe`): the tile map is built by passb_desc_kernel from the cursors
            std::vector<PassBBucket> desc(P + 1);
            u64 ticket[2] = {7, 7};
            passb_desc_kernel(cursor.data(), seg_cap, P, (u32)tk, (const char *)table_v.data(), table_v.size(), nullptr, 0, 8u, 1,
                              ticket, desc.data());
            pass_b(desc.data(), 0, P, ticket);
        } else {
            // host-built tile map, Pass B in two launches (pilot + rest)
            const u32 cuts[3] = {0, 2, P};
            for (int part = 0; part < 2; ++part) {
                const u32 b0 = cuts[part], b1 = cuts[part + 1];
                std::vector<PassBBucket> desc(b1 - b0 + 1);
                u64 tiles = 0, ticket = 0;
                for (u32 b = b0; b < b1; ++b) { desc[b - b0] = PassBBucket{tiles, count[b], nullptr, nullptr, 0, 0}; tiles += (count[b] + tk - 1) / tk; }
                desc[b1 - b0] = PassBBucket{tiles, 0, nullptr, nullptr, 0, 0};
                pass_b(desc.data(), b0, b1, &ticket);
            }
        }
    }
    if (ctr.overflow_n) {          // grow + rehash + re-insert, as pbk_api.cu does
        std::vector<slot_t> bigger(table_slots * 4);
        memset(bigger.data(), 0, bigger.size() * sizeof(slot_t));
        rehash_kernel<W>(table, Table<W>(bigger.data(), bigger.size()), &ctr);
        table_v.swap(bigger);
        table_slots *= 4;
        table = Table<W>(table_v.data(), table_v.size());
        std::vector<u64> copy(ovf.begin(), ovf.begin() + std::min<u64>(ctr.overflow_n, OVF) * (W + 1));
        const u64 n = std::min<u64>(ctr.overflow_n, OVF);
        ctr.overflow_n = 0;
        insert_records_kernel<W>(copy.data(), n, 1, table, remote, n_shards, rank, &ctr,
                                 ovf.data(), OVF);
    }
    histogram_kernel<W>(table, occ_hist);
    *n_out = 0;
    export_kernel<W>(table, min_count, keys_out, counts_out, cap_out, n_out);
    *n_inst = ctr.instances;
    *err_flags = ctr.error_flags;
    *n_remote = 0;
    if (n_shards > 1 && remote_records) {
        std::vector<u64> cnt(n_shards, 0), cur(n_shards, 0);
        shard_count_kernel<W>(remote, n_shards, cnt.data());
        u64 tot = 0;
        for (u32 i = 0; i < n_shards; ++i) { cur[i] = tot; tot += cnt[i]; }
        shard_pack_kernel<W>(remote, n_shards, cur.data(), remote_records);
        *n_remote = tot;
    }
    return 0;
}

extern "C" int emul_count(const uint8_t *bases, const u64 *off, u64 n_reads, int k, int encoding,
                          const int32_t *n_pos, const u64 *n_pos_off, u64 table_slots, u32 n_shards, u32 rank,
                          u32 min_count, u64 *keys_out, uint16_t *counts_out, u64 cap_out, u64 *n_out,
                          u64 *occ_hist, u64 *len_hist, u64 *n_inst, u32 *err_flags, u64 *remote_records,
                          u64 *n_remote)
{
#define GO(Wv) case Wv: return run<Wv>(bases, off, n_reads, k, encoding, n_pos, n_pos_off, table_slots, n_shards, rank, \
                                      min_count, keys_out, counts_out, cap_out, n_out, occ_hist, len_hist, n_inst, \
                                      err_flags, remote_records, n_remote);
    switch ((k + 31) / 32) { GO(1) GO(2) GO(3) GO(4) GO(5) GO(6) GO(7) GO(8) default: return -1; }
}

// insert weighted records (the receive side of the shard exchange) into a fresh table and dump it
extern "C" int emul_insert_records(const u64 *records, u64 n, int k, u64 table_slots, u64 *keys_out,
                                   uint16_t *counts_out, u64 cap_out, u64 *n_out)
{
    const int Wd = (k + 31) / 32;
    Counters ctr{};
    std::vector<u64> ovf(16 * 9);
#define INS(Wv) case Wv: { if (Wv == 1) table_slots = 1ull << 24;                                                  \
        std::vector<SlotType<Wv>::type> tv(table_slots);                                                             \
        memset(tv.data(), 0, tv.size() * sizeof(SlotType<Wv>::type));                                                \
        Table<Wv> t(tv.data(), tv.size());                                                                           \
        insert_records_kernel<Wv>(records, n, 1, t, t, 1, 0, &ctr, ovf.data(), 16);                                  \
        *n_out = 0; export_kernel<Wv>(t, 1, keys_out, counts_out, cap_out, n_out);                                   \
        return ctr.overflow_n ? -2 : 0; }
    switch (Wd) { INS(1) INS(2) INS(3) INS(4) INS(5) INS(6) INS(7) INS(8) default: return -1; }
}

// ---- key exchange (pbk_keyx_*, k <= 32) -----------------------------------------------------------------------------
// Pass A of one rank into the all-to-all send buffer: [dest][region][seg_cap] hashes + [dest][region] cursors.  Keys that
// find their segment full come back as (key, weight) records -- pbk_api.cu routes those through the record exchange.
extern "C" int emul_keyx_partition(const uint8_t *bases, const u64 *off, u64 n_reads, int k, u32 n_dest, u32 n_regions,
                                   u64 seg_cap, u32 bin_cap, u64 *send, u64 *cursors, u64 *n_inst, u64 *spill_records,
                                   u64 spill_cap, u64 *n_spill, u32 *err_flags)
{
    if (k > 32) return -1;
    const u64 n_bases = off[n_reads], words = (n_bases + 31) / 32;
    std::vector<u64> stream(words + STREAM_PAD_WORDS + 1, 0), len_hist(500001, 0);
    std::vector<u32> nflag(words + STREAM_PAD_WORDS + 1, 0), rflag(words + STREAM_PAD_WORDS + 1, 0);
    Counters ctr{};
    read_marks_kernel(off, n_reads, len_hist.data(), rflag.data() + STREAM_PAD_WORDS, &ctr);
    pack_kernel<false>(bases, n_bases, words, 0, stream.data() + STREAM_PAD_WORDS, nflag.data() + STREAM_PAD_WORDS, 0, &ctr);
    const u32 P = n_dest * n_regions;
    for (u32 b = 0; b < P; ++b) cursors[b] = 0;
    const u64 w_split = words / 2;
    for (int part = 0; part < 2; ++part)
        partition_kernel<1, true>(stream.data() + STREAM_PAD_WORDS, nflag.data() + STREAM_PAD_WORDS, rflag.data() + STREAM_PAD_WORDS,
                                  part ? w_split : 0, part ? words : w_split, k, P, bin_cap, send, seg_cap, cursors, &ctr,
                                  spill_records, spill_cap, n_dest);
    *n_inst = ctr.instances;
    *n_spill = ctr.overflow_n;
    *err_flags = ctr.error_flags;
    return 0;
}

// Pass B of one rank over what the all-to-all delivered ([source][region][seg_cap], [source][region] cursors), descriptors
// built as launch_bucket_insert_gathered builds them; `extra` = (key, weight) records that arrived by the record route
extern "C" int emul_keyx_insert(const u64 *recv, const u64 *recv_cursors, u32 n_src, u32 n_regions, u64 seg_cap, int k,
                                const u64 *extra, u64 n_extra, u64 *keys_out, uint16_t *counts_out, u64 cap_out, u64 *n_out)
{
    if (k > 32) return -1;
    const u64 table_slots = 1ull << 24;
    std::vector<u64> tv(table_slots, 0);
    Table<1> t(tv.data(), tv.size());
    Counters ctr{};
    const u64 OVF = 1 << 12;
    std::vector<u64> ovf(2 * OVF);
    const u32 n_desc = n_src * n_regions;
    const u64 tk = PASSB1_KPT * PASSB1_ROUNDS;                     // blockDim = 1 in the emulation
    // two launches (pilot + rest), like pbk_keyx_insert_device on a first batch
    const u32 cuts[3] = {0, n_src * (n_regions > 1 ? 1u : 0u), n_desc};
    for (int part = 0; part < 2; ++part) {
        const u32 d0 = cuts[part], d1 = cuts[part + 1];
        if (d1 <= d0) continue;
        std::vector<PassBBucket> desc(d1 - d0 + 1);
        u64 tiles = 0, ticket[2] = {0, 0};
        for (u32 i = d0; i < d1; ++i) {
            const u64 n = std::min<u64>(recv_cursors[(i % n_src) * n_regions + i / n_src], seg_cap);
            desc[i - d0] = PassBBucket{tiles, n, nullptr, nullptr, 0, 0};
            tiles += (n + tk - 1) / tk;
        }
        desc[d1 - d0] = PassBBucket{tiles, 0, nullptr, nullptr, 0, 0};
        KeyxSources srcs{};
        for (u32 sr = 0; sr < n_src; ++sr) { srcs.keys[sr] = recv + (u64)sr * n_regions * seg_cap; srcs.cursors[sr] = recv_cursors + (u64)sr * n_regions; }
        if (tiles) bucket_insert_gather_staged_kernel(srcs, seg_cap, desc.data(), d0, d1, ticket, t, &ctr, ovf.data(), OVF,
                                               (n_src << 8) | (n_regions << 16));
    }
    if (n_extra) insert_records_kernel<1>(extra, n_extra, 1, t, t, 1, 0, &ctr, ovf.data(), OVF);
    *n_out = 0;
    export_kernel<1>(t, 1, keys_out, counts_out, cap_out, n_out);
    return ctr.overflow_n ? -2 : 0;
}

// ---- occurrence lookup (pbk_lookup) -----------------------------------------------------------------------------------
template <int W>
static int run_lookup(const uint8_t *bases, const u64 *off, u64 n_reads, int k, const u64 *keys, const uint16_t *counts,
                      u64 n, uint16_t *out)
{
    const u64 n_bases = off[n_reads], words = (n_bases + 31) / 32;
    std::vector<u64> stream(words + STREAM_PAD_WORDS + 1, 0), len_hist(500001, 0);
    std::vector<u32> nflag(words + STREAM_PAD_WORDS + 1, 0), rflag(words + STREAM_PAD_WORDS + 1, 0);
    Counters ctr{};
    typedef typename SlotType<W>::type slot_t;
    const u64 slots = W == 1 ? (1ull << 24) : 2 * n + 1024;
    std::vector<slot_t> tv(slots);
    memset(tv.data(), 0, tv.size() * sizeof(slot_t));
    Table<W> table(tv.data(), tv.size());
    std::vector<u64> rec(n * (W + 1) + 1), ovf(16 * (W + 1));
    for (u64 i = 0; i < n; ++i) {
        for (int j = 0; j < W; ++j) rec[i * (W + 1) + j] = keys[i * W + j];
        rec[i * (W + 1) + W] = counts[i];
    }
    insert_records_kernel<W>(rec.data(), n, 1, table, table, 1, 0, &ctr, ovf.data(), 16);
    if (ctr.overflow_n) return -2;
    read_marks_kernel(off, n_reads, len_hist.data(), rflag.data() + STREAM_PAD_WORDS, &ctr);
    pack_kernel<false>(bases, n_bases, words, 0, stream.data() + STREAM_PAD_WORDS, nflag.data() + STREAM_PAD_WORDS, 0, &ctr);
    std::vector<u64> occ4(words * 8 + 1, 0);
    const u64 w_split = words / 2;                     // two launches, like chunks of a larger batch
    lookup_kernel<W>(stream.data() + STREAM_PAD_WORDS, nflag.data() + STREAM_PAD_WORDS, rflag.data() + STREAM_PAD_WORDS, 0, w_split, k, table, occ4.data());
    lookup_kernel<W>(stream.data() + STREAM_PAD_WORDS, nflag.data() + STREAM_PAD_WORDS, rflag.data() + STREAM_PAD_WORDS, w_split, words, k, table, occ4.data());
    const uint16_t *by_end = (const uint16_t *)occ4.data();
    for (u64 i = 0; i < n_bases; ++i) out[i] = 0;
    if (n_bases >= (u64)k)
        for (u64 i = 0; i + k - 1 < n_bases; ++i) out[i] = by_end[i + k - 1];     // window END -> window START, as pbk_lookup does
    return (ctr.error_flags & ERR_BAD_BASE) ? -3 : 0;
}

extern "C" int emul_lookup(const uint8_t *bases, const u64 *off, u64 n_reads, int k, const u64 *keys, const uint16_t *counts,
                           u64 n, uint16_t *out)
{
#define LK(Wv) case Wv: return run_lookup<Wv>(bases, off, n_reads, k, keys, counts, n, out);
    switch ((k + 31) / 32) { LK(1) LK(2) LK(3) LK(4) LK(5) LK(6) LK(7) LK(8) default: return -1; }
}

// ---- iterative-k steps (pbk_seed_entries + pbk_finalize, pbk_match_reads) -----------------------------------------------
template <int W>
static int run_seeded(const uint8_t *bases, const u64 *off, u64 n_reads, int k, const u64 *seed_keys, const uint16_t *seed_counts,
                      u64 n_seed, u64 *keys_out, uint16_t *counts_out, u64 cap_out, u64 *n_out, u64 *n_inst, uint8_t *matched)
{
    const u64 n_bases = off[n_reads], words = (n_bases + 31) / 32;
    std::vector<u64> stream(words + STREAM_PAD_WORDS + 1, 0), len_hist(500001, 0);
    std::vector<u32> nflag(words + STREAM_PAD_WORDS + 1, 0), rflag(words + STREAM_PAD_WORDS + 1, 0);
    Counters ctr{};
    typedef typename SlotType<W>::type slot_t;
    const u64 slots = W == 1 ? (1ull << 24) : 2 * (n_bases + n_seed) + 1024;
    std::vector<slot_t> tv(slots);
    memset(tv.data(), 0, tv.size() * sizeof(slot_t));
    Table<W> table(tv.data(), tv.size());
    const u64 OVF = 1 << 12;
    std::vector<u64> ovf((W + 1) * OVF), rec(n_seed * (W + 1) + 1);
    for (u64 i = 0; i < n_seed; ++i) {
        for (int j = 0; j < W; ++j) rec[i * (W + 1) + j] = seed_keys[i * W + j];
        rec[i * (W + 1) + W] = seed_counts[i];
    }
    read_marks_kernel(off, n_reads, len_hist.data(), rflag.data() + STREAM_PAD_WORDS, &ctr);
    pack_kernel<false>(bases, n_bases, words, 0, stream.data() + STREAM_PAD_WORDS, nflag.data() + STREAM_PAD_WORDS, 0, &ctr);
    if (matched) {
        // pbk_match_reads: the table holds the given entries, nothing is counted
        insert_records_kernel<W>(rec.data(), n_seed, 1, table, table, 1, 0, &ctr, ovf.data(), OVF);
        std::vector<u64> occ4(words * 8 + 1, 0);
        lookup_kernel<W>(stream.data() + STREAM_PAD_WORDS, nflag.data() + STREAM_PAD_WORDS, rflag.data() + STREAM_PAD_WORDS, 0, words, k, table, occ4.data());
        read_match_kernel(off, n_reads, (const uint16_t *)occ4.data(), k, matched);
        return ctr.overflow_n ? -2 : 0;
    }
    // pbk_push_reads ... pbk_seed_entries ... pbk_finalize: count everything, then the seeded keys get their value
    count_kernel<W>(stream.data() + STREAM_PAD_WORDS, nflag.data() + STREAM_PAD_WORDS, rflag.data() + STREAM_PAD_WORDS, 0, words, k,
                    table, table, 1, 0, &ctr, ovf.data(), OVF);
    if (ctr.overflow_n) return -2;
    override_records_kernel<W>(rec.data(), n_seed, table, &ctr, ovf.data(), OVF);
    if (ctr.overflow_n) return -2;
    *n_out = 0;
    export_kernel<W>(table, 1, keys_out, counts_out, cap_out, n_out);
    *n_inst = ctr.instances;
    return 0;
}

extern "C" int emul_seeded(const uint8_t *bases, const u64 *off, u64 n_reads, int k, const u64 *seed_keys, const uint16_t *seed_counts,
                           u64 n_seed, u64 *keys_out, uint16_t *counts_out, u64 cap_out, u64 *n_out, u64 *n_inst, uint8_t *matched)
{
#define SD(Wv) case Wv: return run_seeded<Wv>(bases, off, n_reads, k, seed_keys, seed_counts, n_seed, keys_out, counts_out, cap_out, n_out, n_inst, matched);
    switch ((k + 31) / 32) { SD(1) SD(2) SD(3) SD(4) SD(5) SD(6) SD(7) SD(8) default: return -1; }
}

// ---- table from contigs (pbk_push_contigs) -------------------------------------------------------------------------------
template <int W>
static int run_contigs(const uint8_t *bases, const u64 *off, u64 n_seqs, int k, const uint16_t *coverage, u64 min_occ,
                       u64 *keys_out, uint16_t *counts_out, u64 cap_out, u64 *n_out)
{
    const u64 n_bases = off[n_seqs], words = (n_bases + 31) / 32;
    std::vector<u64> stream(words + STREAM_PAD_WORDS + 1, 0), len_hist(500001, 0);
    std::vector<u32> nflag(words + STREAM_PAD_WORDS + 1, 0), rflag(words + STREAM_PAD_WORDS + 1, 0);
    std::vector<uint16_t> val(words * 32, 0);
    for (u64 r = 0; r < n_seqs; ++r) {
        const u64 v = std::min<u64>(std::max<u64>(coverage[r], min_occ), COUNT_SAT);
        for (u64 p = off[r]; p < off[r + 1]; ++p) val[p] = (uint16_t)v;
    }
    Counters ctr{};
    typedef typename SlotType<W>::type slot_t;
    const u64 slots = W == 1 ? (1ull << 24) : 2 * n_bases + 1024;
    std::vector<slot_t> tv(slots);
    memset(tv.data(), 0, tv.size() * sizeof(slot_t));
    Table<W> table(tv.data(), tv.size());
    read_marks_kernel(off, n_seqs, len_hist.data(), rflag.data() + STREAM_PAD_WORDS, &ctr);
    pack_kernel<false>(bases, n_bases, words, 0, stream.data() + STREAM_PAD_WORDS, nflag.data() + STREAM_PAD_WORDS, 0, &ctr);
    contig_max_kernel<W>(stream.data() + STREAM_PAD_WORDS, nflag.data() + STREAM_PAD_WORDS, rflag.data() + STREAM_PAD_WORDS, 0, words, k,
                         table, val.data(), &ctr);
    if (ctr.overflow_n) return -2;
    *n_out = 0;
    export_kernel<W>(table, 1, keys_out, counts_out, cap_out, n_out);
    return 0;
}

extern "C" int emul_contigs(const uint8_t *bases, const u64 *off, u64 n_seqs, int k, const uint16_t *coverage, u64 min_occ,
                            u64 *keys_out, uint16_t *counts_out, u64 cap_out, u64 *n_out)
{
#define CT(Wv) case Wv: return run_contigs<Wv>(bases, off, n_seqs, k, coverage, min_occ, keys_out, counts_out, cap_out, n_out);
    switch ((k + 31) / 32) { CT(1) CT(2) CT(3) CT(4) CT(5) CT(6) CT(7) CT(8) default: return -1; }
}
