// pbk_kernels.cuh -- launch wrappers of the hand-written sm_100a kernels (defined in pbk_kernels.cu)
#pragma once

#include "pbk_device.cuh"

namespace pbk {

struct TableView {
    void *slots;       // u64[cap] (k <= 32, cap a power of two >= 2^27) or Slot<W>[cap] (k > 32, any cap)
    u64   cap;         // number of slots
    int   words;       // W
    u64   capacity() const { return cap; }
    size_t slot_bytes() const { return words == 1 ? 8 : words <= 3 ? 32 : 8 * (size_t)words + 8; }   // compact / sector-sized / wide slots
    size_t bytes() const { return slot_bytes() * cap; }
};

struct ShardInfo {
    u32 n_shards;      // <= 1: everything is local
    u32 rank;
};

// ---- ingest -----------------------------------------------------------------------------------
// ASCII (or platanus code bytes) -> 2-bit stream words + N flags for stream words [word0, word0 + n_words).
// `bases` points at the byte of stream position word0 * 32; n_valid = number of real bases from there.
void launch_pack(const uint8_t *bases, u64 n_valid, u64 n_words, int encoding,
                 u64 *stream, u32 *nflag, u64 word0, Counters *ctr, cudaStream_t st);
// per read: length histogram, read-start flags, MAX_READ_LEN check
void launch_read_marks(const u64 *offsets, u64 n_reads, u64 *len_hist, u32 *rflag, Counters *ctr, cudaStream_t st);
// PBK_ENC_PLATANUS: scatter the N position lists into nflag
void launch_npos_scatter(const u64 *offsets, const int32_t *n_pos, const u64 *n_pos_offsets, u64 n_reads,
                         u32 *nflag, cudaStream_t st);

// PBK_ENC_PACKED2: absolute N positions + the padding behind the last base into nflag
void launch_npos_abs_scatter(const u64 *n_positions, u64 n_n, u64 n_bases, u32 *nflag, cudaStream_t st);

// ---- counting ---------------------------------------------------------------------------------
// windows that END in stream words [word_begin, word_end) are canonicalised and inserted
void launch_count(const u64 *stream, const u32 *nflag, const u32 *rflag, u64 word_begin, u64 word_end, int k,
                  TableView table, TableView remote, ShardInfo shard, Counters *ctr,
                  u64 *overflow_keys, u64 overflow_cap, int sm_count, cudaStream_t st, u32 pass = 0);
// occ[p] (u16, one per stream position of words [word_begin, word_end); 8-byte aligned) = clamped count of the window
// that ENDS at p, 0 if unusable or absent.  Read-only on the table.
void launch_lookup(const u64 *stream, const u32 *nflag, const u32 *rflag, u64 word_begin, u64 word_end, int k,
                   TableView table, uint16_t *occ, int sm_count, cudaStream_t st);
// table[k-mer of the window ending at p] = max(itself, val[p]) over stream words [word_begin, word_end)
void launch_contig_max(const u64 *stream, const u32 *nflag, const u32 *rflag, u64 word_begin, u64 word_end, int k,
                       TableView table, const uint16_t *val, Counters *ctr, int sm_count, cudaStream_t st);
// SET the count of n (key words..., value) records (insert those that are absent); value 0 = skip
void launch_override_records(const u64 *records, u64 n, TableView table, Counters *ctr, u64 *overflow_keys, u64 overflow_cap,
                             int sm_count, cudaStream_t st);
// out[i] = (leftFlags << 4) | rightFlags of key i: which of its eight neighbours are in the table with count >= min_count
void launch_neighbor_flags(const u64 *keys, u64 n, int k, TableView table, u32 min_count, uint8_t *out, int sm_count, cudaStream_t st);
// matched[r] = any window of read r found (occ: launch_lookup's output)
void launch_read_match(const u64 *offsets, u64 n_reads, const uint16_t *occ, int k, uint8_t *matched, int sm_count, cudaStream_t st);
// insert n (key words..., weight) records; a record has W + 1 words when weighted, W otherwise.
// Overflow-list records always have W + 1 words.
void launch_insert_records(const u64 *records, u64 n, bool weighted, TableView table, TableView remote,
                           ShardInfo shard, Counters *ctr, u64 *overflow_keys, u64 overflow_cap, int sm_count,
                           cudaStream_t st);

// ---- partitioned counting (large batches) -------------------------------------------------------
struct PartitionPlan {
    u32    n_buckets;   // P hash-range buckets == P contiguous table regions
    u32    bin_cap;     // entries per shared-memory bin
    int    threads;     // CTA size of the partition kernel (one stream word per thread and tile)
    size_t smem;        // dynamic shared memory of the partition kernel
    u64    seg_cap;     // entries per bucket segment in the bucket store
};
// shared-memory budget -> bucket count / bin size / CTA size for keys of `words` words
PartitionPlan plan_partition(u64 est_table_bytes, u64 windows_ub, int words, size_t smem_budget = 0);
// Pass A over stream words [word_begin, word_end): keys go to bkt_keys (P segments of seg_cap entries)
void launch_partition(const u64 *stream, const u32 *nflag, const u32 *rflag, u64 word_begin, u64 word_end, int k,
                      int words, const PartitionPlan &plan, u64 *bkt_keys, u64 *bkt_cursor, Counters *ctr,
                      u64 *overflow_keys, u64 overflow_cap, int sm_count, cudaStream_t st, u32 keyx_dest = 1, u32 pass = 0);
// (pass != 0 in launch_count / launch_partition: (n_passes << 16) | pass_index -- a hash-range pass over the whole input on one
//  GPU, pbk_config.n_passes: only the keys with shard_of_hash(hash, n_passes) == pass_index are counted, the others skipped)
// Key exchange between GPUs (k <= 32, pbk_keyx_*): plan with n_dest x n_regions buckets in destination-major order
// (launch_partition with keyx_dest = n_dest fills it), and Pass B over the all-to-all receive buffer.
PartitionPlan plan_partition_keyx(u32 n_dest, u64 max_windows_any_rank, int words, size_t smem_budget = 0);
void launch_bucket_insert_gathered(const KeyxSources &srcs, u64 seg_cap, const u64 *counts, void *h_desc, void *d_desc,
                                   u32 d_first, u32 d_end, u32 n_src, u32 n_regions, TableView table, Counters *ctr,
                                   u64 *overflow_keys, u64 overflow_cap, int sm_count, cudaStream_t st);
void launch_bucket_insert_gathered_chained(const KeyxSources &srcs, u64 seg_cap, void *d_desc, u32 n_src,
                                           u32 n_regions, TableView table, Counters *ctr, u64 *overflow_keys, u64 overflow_cap,
                                           int sm_count, cudaStream_t st);
// Pass B over buckets [b_first, b_end) in one launch.  `h_desc` is scratch for b_end - b_first + 1 bucket
// descriptors (pinned host memory), `d_desc` the same on the device; `counts[b]` = keys in bucket b.
size_t passb_desc_bytes(u32 n_buckets);
void launch_bucket_insert(const u64 *bkt_keys, u64 seg_cap, const u64 *counts, void *h_desc, void *d_desc,
                          u32 b_first, u32 b_end, u32 n_buckets, TableView table, TableView remote, ShardInfo shard,
                          Counters *ctr, u64 *overflow_keys, u64 overflow_cap, int sm_count, cudaStream_t st);

// k <= 32, device-chained: build the tile map from the cursors on the GPU (stream `st`), then Pass B over all buckets
void launch_passb_desc(const u64 *d_cursor, u64 seg_cap, u32 n_buckets, TableView table, TableView remote, ShardInfo shard,
                       void *d_desc, cudaStream_t st);
void launch_bucket_insert_chained(const u64 *bkt_keys, u64 seg_cap, const void *d_desc, u32 n_buckets, TableView table,
                                  TableView remote, ShardInfo shard, Counters *ctr, u64 *overflow_keys, u64 overflow_cap,
                                  int sm_count, int ctas_per_sm, cudaStream_t st);

// k <= 32, unsharded table, second form of Pass B (PBK_PASSB2): every bucket's keys are split once more by SUB-REGION of the
// table (split_kernel), and a CTA then builds each 64 KB sub-region in shared memory (region_build_kernel) -- no L2 atomics.
struct Passb2Geom {
    u32 F;            // sub-regions per bucket
    u64 n_sub;        // sub-regions of the whole table
    int sub_shift;    // hash >> sub_shift = global sub-region
};
bool passb2_geom(TableView table, u32 n_buckets, Passb2Geom *out);        // false: this table / bucket count cannot take the route
u64  passb2_sub_cap(u64 windows_ub, u64 n_sub);                            // entries per sub-region segment
// tile map of all n_buckets buckets for split_kernel, built on the device from the cursors; _gather: of the n_src x n_regions
// descriptors of a key exchange (descriptor i = table region i / n_src, source i % n_src)
void launch_passb2_desc(const u64 *d_cursor, u64 seg_cap, u32 n_buckets, void *d_desc, cudaStream_t st);
void launch_passb2_desc_gather(const KeyxSources &srcs, u64 seg_cap, u32 n_src, u32 n_regions, void *d_desc, cudaStream_t st);
// descriptors [d_first, d_end); srcs == nullptr: the context's own bucket store (descriptor = bucket).
// d_sub_keys: geom.n_sub segments of sub_cap hashes, d_sub_cursor: geom.n_sub fill counts (zeroed by the caller)
void launch_passb2_split(const u64 *bkt_keys, const KeyxSources *srcs, u32 n_src, u64 seg_cap, const void *d_desc, u32 d_first, u32 d_end,
                         const Passb2Geom &geom, u64 *d_sub_keys, u64 sub_cap, u64 *d_sub_cursor, Counters *ctr, u64 *overflow_keys,
                         u64 overflow_cap, int sm_count, cudaStream_t st);
// buckets (table regions) [b_first, b_end).
// load_existing = false: nothing has been inserted since the table was zero-filled
void launch_passb2_build(const u64 *d_sub_keys, u64 sub_cap, const u64 *d_sub_cursor, u32 b_first, u32 b_end, const Passb2Geom &geom,
                         TableView table, bool load_existing, Counters *ctr, u64 *overflow_keys, u64 overflow_cap, int sm_count,
                         cudaStream_t st);

// ---- table ------------------------------------------------------------------------------------
void launch_table_init(TableView t, cudaStream_t st);
// move every entry of `from` into `to` (capacity change)
void launch_table_rehash(TableView from, TableView to, Counters *ctr, cudaStream_t st);
// clamp counts to 65534 (keeps the 32-bit counters far from overflow on > 4G-instance inputs)
void launch_table_clamp(TableView t, cudaStream_t st);
// clamp + occurrence histogram (writeKmerDistribution, counter.h:483-507)
void launch_table_histogram(TableView t, u64 *occ_hist, cudaStream_t st);
// compact entries with count >= min_count: keys (n x W, word 0 first), counts u16; *d_n_out += n
void launch_table_export(TableView t, u32 min_count, u64 *keys_out, uint16_t *counts_out, u64 capacity,
                         u64 *d_n_out, cudaStream_t st);
// sort exported entries ascending in reference order (top word first); queued on `st`, temporaries carved from `scratch`
// (sort_export_scratch_bytes(n, words) bytes, provided by the caller from the context's HBM budget)
size_t sort_export_scratch_bytes(u64 n, int words);
cudaError_t sort_export(u64 *keys, uint16_t *counts, u64 n, int words, int k, void *scratch, size_t scratch_bytes, cudaStream_t st);

// ---- sharding ---------------------------------------------------------------------------------
// per-destination number of entries in the remote-staging table
void launch_shard_count(TableView remote, u32 n_shards, u64 *d_counts, cudaStream_t st);
// write (key words, count) records grouped by destination; d_cursors = exclusive prefix of counts
void launch_shard_pack(TableView remote, u32 n_shards, u64 *d_cursors, u64 *records_out, cudaStream_t st);

// ---- microbenchmark ---------------------------------------------------------------------------
void launch_microbench(void *table, int log2slots, u64 n_ops, int mode, u64 seed, int sm_count, cudaStream_t st);

}  // namespace pbk
