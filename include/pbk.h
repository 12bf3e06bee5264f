/*
 * pbk.h -- C ABI of libpbk.so: B200-native (sm_100a) k-mer occurrence counting for Platanus_B.
 *
 * This is the drop-in boundary for ONE path of the reference: what
 *     Counter<KMER>::makeKmerReadDistributionMT            (reference counter.h:276-383)
 * and the functions it drives (countKmerPerThreadFirst counter.h:391-434, countKmerOrWriteTemporary
 * counter.h:459-476, writeKmerDistribution counter.h:483-507) compute, plus the consumers of its
 * result that run before the `-kmer_occ_only` return: getLeftLocalMinimalValue (counter.h:245-267),
 * sortedKeyFromKmerFile (counter.h:917-951), loadKmer (counter.h:600-640) and
 * outputOccurrenceTableBinary / DoubleHash::writeTable (counter.h:955-963, doubleHash.h:266-278).
 *
 * The reference has no FFI layer: the seam is the header-only template
 * Counter<KMER>.  The host C++ shim that re-implements those member functions on top of this ABI
 * is platanus_b_b200/host/pbk_counter.hpp; INTEGRATION.md shows how a maintainer wires it in.
 *
 * Conventions: plain C types only; every function returns 0 (PBK_OK) or a negative pbk_status;
 * no exceptions cross the boundary; the caller owns every host buffer it passes in; the context
 * owns all device memory and streams.  A context is used from one host thread at a time.
 * There is NO CPU fallback: without a usable CUDA device pbk_create fails with PBK_E_NO_DEVICE.
 */
#ifndef PBK_H
#define PBK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PBK_ABI_VERSION   1
#define PBK_OCC_BINS      65535u    /* counter.h:336: occurrenceDistribution.resize(UINT16_MAX)      */
#define PBK_LEN_BINS      500001u   /* counter.h:328: lengthDistribution, MAX_READ_LEN + 1 bins      */
#define PBK_COUNT_SAT     65534u    /* counter.h:468: counts saturate at UINT16_MAX - 1              */
#define PBK_MAX_K         256u      /* keys of up to 8 x 64-bit words (reference: unbounded binstr_t) */

typedef enum pbk_status {
    PBK_OK              =  0,
    PBK_E_ARG           = -1,   /* bad argument                                                    */
    PBK_E_NO_DEVICE     = -2,   /* no CUDA device / driver: there is no CPU fallback                */
    PBK_E_CUDA          = -3,   /* a CUDA call failed; pbk_last_error() has the text                */
    PBK_E_NOMEM         = -4,   /* device (HBM budget) or host memory exhausted                     */
    PBK_E_READ_TOO_LONG = -5,   /* a read has >= 500000 bases: platanus::ReadError (common.h:465)   */
    PBK_E_BAD_BASE      = -6,   /* a character whose Char2Bin code (common.h:256) is undefined      */
    PBK_E_KMER_DIST     = -7,   /* empty distribution: platanus::KmerDistError (counter.h:225-237)  */
    PBK_E_STATE         = -8,   /* call out of order (e.g. export before finalize)                  */
    PBK_E_IO            = -9,   /* file could not be written: platanus::FILEError                   */
    PBK_E_UNSUPPORTED_K = -10   /* k == 0 or k > PBK_MAX_K                                          */
} pbk_status;

typedef struct pbk_ctx pbk_ctx;

enum {
    PBK_ENC_ASCII    = 0,  /* raw characters as the parser hands them to SEQ::convertFromString     */
    PBK_ENC_PLATANUS = 1,  /* SEQ temp-file form (common.h:426-448): bytes 0..3, N positions listed */
    PBK_ENC_PACKED2  = 2   /* 2-bit words + absolute N positions: pbk_push_reads_packed only          */
};

enum {
    PBK_F_TIMING = 1u << 0,       /* bracket every kernel launch with CUDA events (see pbk_get_stats) */
    PBK_F_NO_PARTITION = 1u << 1, /* always insert straight into the table (no hash-range bucket pass)  */
    PBK_F_FORCE_PARTITION = 1u << 2, /* use the bucket pass even for tiny batches (tests)              */
    PBK_F_NO_PIPELINE = 1u << 3,  /* always size the table with a pilot launch and host round trips      */
    PBK_F_UNKNOWN_AS_N = 1u << 4  /* a character without a Char2Bin code (IUPAC ambiguity codes ...) counts as N instead of failing
                                     with PBK_E_BAD_BASE (the reference silently miscodes it, common.h:256); also: PBK_UNKNOWN_AS_N=1 */
};

typedef struct pbk_config {
    uint32_t struct_size;        /* sizeof(pbk_config), for ABI evolution                           */
    uint32_t k;                  /* k-mer length, 1..PBK_MAX_K                                      */
    int32_t  device;             /* CUDA device ordinal, -1 = current device                        */
    uint32_t flags;              /* PBK_F_*                                                         */
    uint32_t n_shards;           /* 0/1 = unsharded; >1: this context owns hash range `shard_rank`  */
    uint32_t shard_rank;
    uint64_t table_slots_hint;   /* initial table capacity in slots, 0 = derive from the first push */
    uint64_t hbm_budget_bytes;   /* cap on device memory used by the context, 0 = 85% of free HBM   */
    /* Hash-range passes on ONE GPU, for a table that does not fit the HBM budget -- the counterpart of the reference's
     * memory-limited mode, where k-mers that find no room are written to temporary files and counted in further rounds
     * (counter.h:340-364, 442-449): with n_passes >= 2 this context counts ONLY the k-mers whose hash falls into range
     * pass_index of n_passes (the ownership function of the sharded forms, pbk_shard_of_key); the caller pushes the same
     * reads once per pass (the reference re-reads its temp files, too) and adds up what pbk_finalize / pbk_export return:
     * the passes' key sets are disjoint.  n_instances counts the windows of this pass only.  Not together with n_shards > 1;
     * pbk_push_contigs is not available.  pbk::Counter (host/pbk_counter.hpp) drives the passes when a plain count ends
     * with PBK_E_NOMEM.  A config of the earlier, shorter layout (struct_size 40) is accepted: no passes.             */
    uint32_t n_passes;           /* 0/1 = everything in one pass                                    */
    uint32_t pass_index;
} pbk_config;
#define PBK_CONFIG_SIZE_V1 40u   /* sizeof(pbk_config) before n_passes / pass_index were added */

/* counters of the work done so far; durations are CUDA-event times on the context's own streams */
typedef struct pbk_stats {
    uint64_t n_reads;
    uint64_t n_bases;
    uint64_t n_instances;        /* k-mer windows inserted                                          */
    uint64_t n_distinct;
    uint64_t table_slots;
    uint64_t table_bytes;
    uint64_t n_grow;             /* table rebuilds                                                  */
    uint64_t launches_pack;      /* kernel launches by class                                        */
    uint64_t launches_count;
    uint64_t launches_other;
    double   ms_pack;            /* only filled with PBK_F_TIMING                                   */
    double   ms_count;
    double   ms_other;
    uint64_t h2d_bytes;
    uint64_t d2h_bytes;
    uint64_t launches_partition; /* Pass A (hash-range bucket scatter) launches, also counted in launches_count */
    uint64_t launches_insert;    /* Pass B (bucket -> table) launches, also counted in launches_count           */
    double   ms_partition;       /* parts of ms_count                                                          */
    double   ms_insert;
    double   ms_count_elapsed;   /* counting phase, begin to end (= ms_count: the launches are serial)          */
    uint64_t n_pipelined_batches;/* pushes whose Pass A / tile map / Pass B were chained on the GPU (no host sync) */
    uint64_t n_split_build;      /* Pass B runs that took the second form: keys split by 64 KB sub-region of the table, every
                                    sub-region built in shared memory (split_kernel + region_build_kernel; k <= 32, unsharded)      */
} pbk_stats;

const char *pbk_strerror(int status);
/* text of the last failure on this context (valid until the next call on it) */
const char *pbk_last_error(const pbk_ctx *ctx);
int  pbk_abi_version(void);

int  pbk_create(pbk_ctx **out, const pbk_config *cfg);
void pbk_destroy(pbk_ctx *ctx);

/* Pinned host memory for callers that want the H2D copies to run at full PCIe speed. */
int  pbk_host_alloc(void **out, size_t bytes);
void pbk_host_free(void *p);

/*
 * Count the k-mers of a batch of reads (may be called repeatedly before pbk_finalize).
 *   bases         concatenated reads, read_offsets[n_reads] bytes, no separators
 *   read_offsets  n_reads + 1 ascending byte offsets into `bases`
 *   encoding      PBK_ENC_ASCII: characters (upper or lower case ACGTN, Char2Bin semantics).
 *                 PBK_ENC_PLATANUS: codes 0..3; N positions in n_pos[n_pos_offsets[r] ..
 *                 n_pos_offsets[r+1]) (read-relative, as SEQ::positionUnknown); the byte under an N
 *                 is ignored exactly like the reference (common.h:468-476).
 * Replaces the per-thread loops of counter.h:322-325 over the read temp files.
 * Host pointers.  Returns after the batch has been queued and its inputs consumed.
 */
int  pbk_push_reads(pbk_ctx *ctx, const uint8_t *bases, const uint64_t *read_offsets, uint64_t n_reads,
                    int encoding, const int32_t *n_pos, const uint64_t *n_pos_offsets);

/*
 * Opt-in third input form: the bases already packed to 2 bits by the host parser as it copies each record -- exactly the
 * library's device layout (32 bases per u64, first base in the two most significant bits; reads concatenated without
 * separators; the last word zero-padded), so the H2D copy goes straight into the stream buffer, there is no pack kernel,
 * and PCIe carries 0.25 byte per base instead of 1 (host buffers through pbk_push_reads(PBK_ENC_ASCII) cap at
 * ~44 G k-mers/s on a 55 GB/s link whatever the kernels do).
 *   words         ceil(read_offsets[n_reads] / 32) u64
 *   n_positions   stream positions (0-based, over the whole batch) of every N, ascending or not; the two bits under an N
 *                 are ignored
 * pbk_pack_reads produces this form from ASCII with Char2Bin semantics (common.h:256): returns PBK_E_BAD_BASE for a
 * character without a code; *n_n_out = number of N found (n_positions_out filled up to n_cap; PBK_E_ARG if more).
 * Reads of this form may be mixed with the other pushes; ASCII stays the default form the reference's parser hands over.   */
int  pbk_pack_reads(const uint8_t *bases, uint64_t n_bases, uint64_t *words_out, uint64_t *n_positions_out,
                    uint64_t n_cap, uint64_t *n_n_out);
int  pbk_push_reads_packed(pbk_ctx *ctx, const uint64_t *words, const uint64_t *read_offsets, uint64_t n_reads,
                           const uint64_t *n_positions, uint64_t n_n);

/* Same with `bases` and `read_offsets` already resident in device memory (PBK_ENC_ASCII only). */
int  pbk_push_reads_device(pbk_ctx *ctx, const void *d_bases, const void *d_read_offsets,
                           uint64_t n_reads, uint64_t n_bases);

/*
 * Finish counting: clamp counts to 65534 and build the distributions
 * (writeKmerDistribution counter.h:483-507, counter.h:328-333, 371-376).
 *   occ_hist   PBK_OCC_BINS u64, [c] = number of distinct k-mers with clamped count c
 *   len_hist   PBK_LEN_BINS u64 read-length histogram over all reads; may be NULL
 * Any output pointer may be NULL.
 */
int  pbk_finalize(pbk_ctx *ctx, uint64_t *occ_hist, uint64_t *len_hist,
                  uint64_t *n_distinct, uint64_t *n_instances, uint64_t *max_occurrence);

/*
 * Dump the table: every distinct canonical k-mer with clamped count >= min_count.
 *   keys    n x ceil(k/32) u64, word 0 (the LAST 32 bases, kmer.h:119/237 order) first
 *   counts  n u16
 *   sorted  non-zero: ascending in the reference's numeric order (top word first,
 *           binstr.h:460-466) -- what sortedKeyFromKmerFile (counter.h:917-951) produces
 * With capacity == 0 only *n_out is written (size query).  Needs pbk_finalize first.
 */
int  pbk_export(pbk_ctx *ctx, uint32_t min_count, int sorted, uint64_t *keys, uint16_t *counts,
                uint64_t capacity, uint64_t *n_out);

int  pbk_get_stats(const pbk_ctx *ctx, pbk_stats *out);
/* switch the per-launch CUDA events of PBK_F_TIMING on or off (they cost 3-4 % of a step: measure throughput
 * without them, the per-kernel breakdown with them) */
int  pbk_set_timing(pbk_ctx *ctx, int on);

/* Device-side stopwatch on the context's own compute stream (the stream every kernel of this context
 * is launched on, and that waits for every H2D copy): pbk_timer_mark records CUDA event `slot`
 * (0..7); pbk_timer_elapsed_ms waits for event `stop` and returns the time between two marks. */
int  pbk_timer_mark(pbk_ctx *ctx, int slot);
int  pbk_timer_elapsed_ms(pbk_ctx *ctx, int start, int stop, double *ms);

/* Forget all counts but keep device buffers; optionally change k (k sweep over the same process). */
int  pbk_reset(pbk_ctx *ctx, uint32_t k);

/* ---- hash-range sharding across GPUs (one context per GPU) --------------
 * With cfg.n_shards > 1, pbk_push_reads inserts the k-mers this shard owns and stages the others,
 * pre-aggregated, per destination.  The caller moves them with its collective of choice
 * (torch.distributed / NCCL all-to-all) and feeds what it received to pbk_shard_insert.        */
/* bytes of one staged record: ceil(k/32) u64 key words + one u64 count                            */
uint32_t pbk_shard_record_bytes(const pbk_ctx *ctx);
/* number of staged records per destination shard (n_shards entries; own entry is 0)               */
int  pbk_shard_send_counts(pbk_ctx *ctx, uint64_t *counts);
/* copy the staged records, grouped by destination in shard order, into a device buffer of the
 * caller (device pointer, capacity in records) and clear the staging area                        */
int  pbk_shard_pack_device(pbk_ctx *ctx, void *d_records, uint64_t capacity_records);
/* insert n records (device pointer) received from other shards                                    */
int  pbk_shard_insert_device(pbk_ctx *ctx, const void *d_records, uint64_t n_records);
/* owner shard of a key (host helper, same function as on the device)                              */
uint32_t pbk_shard_of_key(const uint64_t *key_words, uint32_t k, uint32_t n_shards);

/* ---- hash-range sharding, second form (k <= 32): exchange the KEYS before counting --------------
 * Pass A of the counter already sorts a batch's k-mers into hash-range buckets.  Here the buckets are
 * ordered by owner shard first and table region second, in a buffer of the caller, so the bucket store
 * IS the all-to-all send buffer (equal splits, no packing pass, no host round trip for counts); every
 * shard then runs the ordinary Pass B over what it received.  Compared with the record exchange above
 * there is no remote-staging table, no scan/pack of it and no weighted inserts; the wire carries 8 bytes
 * per k-mer instance instead of 16 per pre-aggregated record (which form wins
 * depends on coverage).  Ownership is the same function (pbk_shard_of_key), results are identical.
 *   1. every rank: max over ranks of the batch's window count (n_bases - n_reads * (k - 1)) -> pbk_keyx_plan
 *   2. pbk_keyx_partition[_device]  (pack + Pass A into d_send / d_cursors; returns when they are complete)
 *   3. all-to-all of d_send (bytes_per_dest per rank) and d_cursors (cursors_per_dest u64 per rank)
 *   4. pbk_keyx_insert_device(d_recv, d_recv_cursors)  (the caller has waited for its collective)
 *   5. if pbk_shard_send_counts reports staged records on ANY rank (keys that found their segment full:
 *      a k-mer repeated millions of times), run the record exchange above as well
 *   6. pbk_finalize as usual.                                                                        */
typedef struct pbk_keyx_layout {
    uint32_t n_dest;            /* = n_shards                                                      */
    uint32_t n_regions;         /* table regions (top hash bits) per destination                   */
    uint64_t seg_cap;           /* entries per (destination, region) segment                       */
    uint64_t entry_bytes;       /* 8: a k <= 32 key travels as its 64-bit hash (a bijection)       */
    uint64_t bytes_per_dest;    /* n_regions * seg_cap * entry_bytes: the all-to-all split         */
    uint64_t cursors_per_dest;  /* n_regions u64 fill counts per destination                       */
} pbk_keyx_layout;
/* same inputs -> same layout on every rank; PBK_E_UNSUPPORTED_K for k > 32 (use the record exchange) */
int  pbk_keyx_plan(pbk_ctx *ctx, uint64_t max_windows_any_rank, pbk_keyx_layout *out);
/* d_send: n_dest * bytes_per_dest bytes, d_cursors: n_dest * cursors_per_dest u64 (device pointers of
 * the caller, both overwritten).  Other arguments as pbk_push_reads / pbk_push_reads_device.       */
int  pbk_keyx_partition(pbk_ctx *ctx, const uint8_t *bases, const uint64_t *read_offsets, uint64_t n_reads,
                        int encoding, const int32_t *n_pos, const uint64_t *n_pos_offsets,
                        void *d_send, void *d_cursors);
int  pbk_keyx_partition_device(pbk_ctx *ctx, const void *d_bases, const void *d_read_offsets,
                               uint64_t n_reads, uint64_t n_bases, void *d_send, void *d_cursors);
/* Device-side ordering between the context's compute stream and a stream of the caller (e.g. the one its
 * collective library enqueues on), so that the steps above can be chained without host synchronisation:
 *   pbk_stream_signal  the caller's stream waits for everything queued on the context so far
 *   pbk_stream_wait    the context waits for everything queued on the caller's stream so far
 * `stream` is a cudaStream_t passed as void* (torch: torch.cuda.current_stream().cuda_stream).
 * pbk_keyx_partition_device_async is pbk_keyx_partition_device without the read-back at its end: it returns
 * when its kernels are queued; error flags, the instance count and keys that found their segment full are
 * picked up by the next call that needs the host's view (pbk_finalize, pbk_shard_send_counts, ...).   */
int  pbk_stream_signal(pbk_ctx *ctx, void *stream);
int  pbk_stream_wait(pbk_ctx *ctx, void *stream);
int  pbk_keyx_partition_device_async(pbk_ctx *ctx, const void *d_bases, const void *d_read_offsets,
                                     uint64_t n_reads, uint64_t n_bases, void *d_send, void *d_cursors);
/* d_recv / d_recv_cursors: what the all-to-all delivered, [source rank][region][seg_cap] entries and
 * [source rank][region] fill counts                                                               */
int  pbk_keyx_insert_device(pbk_ctx *ctx, const void *d_recv, const void *d_recv_cursors);

/* ---- hash-range sharding, third form (k <= 32): the pull exchange -- no collective moves the keys ----------------------
 * Pass A fills the context's OWN owner-major bucket store (same layout as above); every shard's Pass B then reads the
 * segments addressed to it in place, out of its peers' HBM over NVLink / NVSwitch (P2P loads through peer-mapped
 * pointers).  The transfer is Pass B's own streamed key loads and overlaps its table atomics tile by tile; between the
 * two passes there is only a barrier.  Results are identical to the other two forms (same ownership function).
 *   1. every rank: pbk_keyx_pull_setup(max over ranks of the batch's window count)  -- plans the layout and allocates the
 *      store (two parities: a store is rewritten two partition calls later, when every peer has finished reading it)
 *   2. connect every peer once: same process -> pbk_keyx_pull_connect_local(ctx, r, ctx_of_rank_r);
 *      other processes -> exchange pbk_keyx_pull_handle() blobs by any means, pbk_keyx_pull_connect_ipc(ctx, r, blob_of_r)
 *   3. per batch: pbk_keyx_pull_partition[_device]  ->  BARRIER over all ranks that is ordered after every rank's
 *      partition on its GPU and before this rank's insert (any collective on a stream bracketed by pbk_stream_signal /
 *      pbk_stream_wait; CUDA events between the contexts of one process; or host synchronisation)  ->  pbk_keyx_pull_insert
 *   4. staged records (keys that found their segment full) as in the second form; pbk_finalize as usual.            */
#define PBK_KEYX_HANDLE_BYTES 64
int  pbk_keyx_pull_setup(pbk_ctx *ctx, uint64_t max_windows_any_rank, pbk_keyx_layout *out);
int  pbk_keyx_pull_handle(pbk_ctx *ctx, void *handle_out /* PBK_KEYX_HANDLE_BYTES */);
int  pbk_keyx_pull_connect_ipc(pbk_ctx *ctx, uint32_t src_rank, const void *handle);
int  pbk_keyx_pull_connect_local(pbk_ctx *ctx, uint32_t src_rank, pbk_ctx *peer);
int  pbk_keyx_pull_partition(pbk_ctx *ctx, const uint8_t *bases, const uint64_t *read_offsets, uint64_t n_reads,
                             int encoding, const int32_t *n_pos, const uint64_t *n_pos_offsets);
/* async != 0: returns when the kernels are queued (see pbk_keyx_partition_device_async) */
int  pbk_keyx_pull_partition_device(pbk_ctx *ctx, const void *d_bases, const void *d_read_offsets, uint64_t n_reads,
                                    uint64_t n_bases, int async);
int  pbk_keyx_pull_insert(pbk_ctx *ctx);
/* queues a copy of two u64 counters -- records staged since pbk_reset, keys waiting on the overflow list (a full segment puts
 * them there) -- into 16 bytes of device memory of the caller: summed over the ranks by the barrier's own collective, a
 * non-zero sum tells every rank that the record route has to run, without a collective and a host round trip of its own */
int  pbk_keyx_staged_count_device(pbk_ctx *ctx, void *d_two_u64);
/* frees the store and drops every mapping (to plan a larger layout; every peer must release before anybody partitions again) */
int  pbk_keyx_pull_release(pbk_ctx *ctx);

/* ---- several GPUs of one box behind one handle (SURVEY.md section 8b "Ownership": the library owns the devices' contexts
 * and the exchange between them) ------------------------------------------------------------------------------------------
 * For a host program that is ONE process, like the reference's `assemble` (iterate.cpp:244-264 spawns it): the group owns
 * one pbk_ctx per device, cuts every batch into one slice per device, drives the devices from one host thread each, and
 * exchanges keys inside the library -- k <= 32: the pull exchange above over peer-mapped HBM (NVLink), k > 32: (key, count)
 * records with cudaMemcpyPeer.  No collective library, no second process.  Results are identical to a one-GPU count.
 * cfg->device / n_shards / shard_rank are ignored; devices == NULL means ordinals 0..n_devices-1.                        */
typedef struct pbk_group pbk_group;
int  pbk_device_count(void);
int  pbk_group_create(pbk_group **out, const pbk_config *cfg, const int32_t *devices, uint32_t n_devices);
void pbk_group_destroy(pbk_group *g);
uint32_t pbk_group_size(const pbk_group *g);
pbk_ctx *pbk_group_member(pbk_group *g, uint32_t i);            /* e.g. for pbk_get_stats of one device */
const char *pbk_group_last_error(const pbk_group *g);
int  pbk_group_reset(pbk_group *g, uint32_t k);
/* as pbk_push_reads: returns when the batch has been counted into the devices' tables */
int  pbk_group_push_reads(pbk_group *g, const uint8_t *bases, const uint64_t *read_offsets, uint64_t n_reads,
                          int encoding, const int32_t *n_pos, const uint64_t *n_pos_offsets);
/* as pbk_finalize, over all devices: the shards' histograms add up (disjoint key sets) */
int  pbk_group_finalize(pbk_group *g, uint64_t *occ_hist, uint64_t *len_hist, uint64_t *n_distinct,
                        uint64_t *n_instances, uint64_t *max_occurrence);
/* as pbk_export, over all devices: sorted != 0 merges the shards' sorted pieces into one ascending list */
int  pbk_group_export(pbk_group *g, uint32_t min_count, int sorted, uint64_t *keys, uint16_t *counts,
                      uint64_t capacity, uint64_t *n_out);
/* as pbk_neighbor_flags, over all devices (OR of the shards' answers) */
int  pbk_group_neighbor_flags(pbk_group *g, uint32_t min_count, const uint64_t *keys, uint64_t n, uint8_t *flags_out);

/* ---- consumers of the table --------------------------------------
 * Occurrence of every k-mer window of a batch of sequences: ContigDivider::getOccurrenceArray
 * (kmer_divide.cpp:151-197, via Counter::findValue) and the table probe of
 * divideKmerUsedMakingPreviousContig (counter.h:828-861).  The table is only read.
 *   occ_out  u16 per BASE: occ_out[read_offsets[r] + i] = min(count, 65534) of the canonical k-mer of
 *            the window starting at base i of read r; 0 if the k-mer is not in the table, if the
 *            window contains an N, and for the last k - 1 bases of every read (no window starts there).
 * With n_shards > 1 only keys this shard owns are found: summing occ_out over the shards gives the
 * global answer.  Other arguments as pbk_push_reads; at most 2^31 bases per call; a sequence may be of any length
 * (contigs: the 500000-base limit of reads does not apply).                                         */
int  pbk_lookup(pbk_ctx *ctx, const uint8_t *bases, const uint64_t *read_offsets, uint64_t n_reads,
                int encoding, const int32_t *n_pos, const uint64_t *n_pos_offsets, uint16_t *occ_out);
/* same with inputs and output in device memory (d_occ_out: n_bases u16, 8-byte aligned)              */
int  pbk_lookup_device(pbk_ctx *ctx, const void *d_bases, const void *d_read_offsets, uint64_t n_reads,
                       uint64_t n_bases, void *d_occ_out);
/* The eight neighbour probes BruijnGraph::makeInitialBruijnGraph makes for every k-mer of sortedKeyFP before it can tell a
 * junction from a straight node (graph.h:337-375: 4 x findValue on the left, 4 on the right, each a random access into the
 * host DoubleHash) -- SURVEY.md section 8f row 4 -- answered for all kept k-mers in one device pass over the resident table:
 *   keys       n x ceil(k/32) words, e.g. what pbk_export(min_count, sorted = 1) returned (stored orientation = canonical)
 *   flags_out  n bytes: (leftFlags << 4) | rightFlags, bit b of leftFlags = "base b + first k-1 bases" is a k-mer with count
 *              >= min_count, bit b of rightFlags = "last k-1 bases + base b" is -- the value Junction::out gets (graph.h:398).
 * For the other orientation of a k-mer (graph.h walks both) mirror it: left'[b] = right[3 - b], right'[b] = left[3 - b].
 * With n_shards > 1 only neighbours this shard owns are found: OR the shards' outputs (pbk_group_neighbor_flags does).       */
int  pbk_neighbor_flags(pbk_ctx *ctx, uint32_t min_count, const uint64_t *keys, uint64_t n, uint8_t *flags_out);
/* Counter::pickupReadMatchedEdgeKmer (counter.h:870-910): matched_out[r] = 1 if a usable k-mer window
 * of read r (no N) is in the table, else 0 -- the reads the next assembly round keeps.               */
int  pbk_match_reads(pbk_ctx *ctx, const uint8_t *bases, const uint64_t *read_offsets, uint64_t n_reads,
                     int encoding, const int32_t *n_pos, const uint64_t *n_pos_offsets, uint8_t *matched_out);
/* Counter::makeKmerReadDistributionConsideringPreviousGraph (counter.h:663-750): the k-mers given here
 * (the table Assemble::saveAndRedoAssemble seeds from the previous round's contigs, assemble.cpp:393-404)
 * keep their value: read windows that hit them are not counted (divideKmerUsedMakingPreviousContig,
 * counter.h:828-861) and every other k-mer of the reads is counted as usual.  Call before pbk_finalize,
 * in any order with pbk_push_reads; entries with value 0 are ignored (counter.h:700); pbk_reset
 * forgets them.                                                                                     */
int  pbk_seed_entries(pbk_ctx *ctx, const uint64_t *keys, const uint16_t *counts, uint64_t n);
/* Counter::makeKmerReadDistributionFromContig (counter.h:511-593): every k-mer of the given sequences gets
 * max(its value so far, max(coverage of its sequence, min_occurrence)) -- a k-mer of several contigs keeps the
 * largest.  pbk_finalize then yields the distribution writeKmerDistribution would (counter.h:582-590).
 * ASCII sequences; windows containing N are skipped (the reference feeds them in as garbage; its contigs
 * carry no N here); sequences may be as long as the whole call (no 500000 limit: read_offsets semantics only). */
int  pbk_push_contigs(pbk_ctx *ctx, const uint8_t *bases, const uint64_t *seq_offsets, uint64_t n_seqs,
                      const uint16_t *coverage, uint64_t min_occurrence);
/* Add n (key, count) entries to the table (host arrays; keys n x ceil(k/32) words, word 0 first):
 * what Counter::readOccurrenceTableBinary leaves in memory, or a contig-seeded table
 * (makeKmerReadDistributionFromContig, counter.h:511-593).  Counts of equal keys add up, saturating. */
int  pbk_load_entries(pbk_ctx *ctx, const uint64_t *keys, const uint16_t *counts, uint64_t n);
/* Counter::readOccurrenceTableBinary (counter.h:967-993) + DoubleHash::readTable (doubleHash.h:280-293):
 * the entries of a PREFIX_kmer_occ.bin as malloc'ed arrays (release with pbk_free).                 */
int  pbk_read_kmer_occ_bin(const char *path, uint32_t *k_out, uint64_t *index_size_out,
                           uint64_t **keys_out, uint16_t **counts_out, uint64_t *n_out);
void pbk_free(void *p);

/* ---- host-side pieces of the path that stay on the CPU (negligible cost) --- */
/* Counter::getLeftLocalMinimalValue (counter.h:245-267) */
uint64_t pbk_left_local_min(const uint64_t *occ_hist, uint64_t max_occurrence, uint64_t window);
/* cutoff rule of Assemble::initialKmerAssemble (assemble.cpp:318-321); n_opt = -n, repeat = -repeat */
uint64_t pbk_coverage_cutoff(const uint64_t *occ_hist, uint64_t max_occurrence, int n_opt, int repeat);
/* Counter::calcDistributionAverage (counter.h:221-238); PBK_E_KMER_DIST when empty */
int  pbk_distribution_average(const uint64_t *dist, uint64_t n_bins, uint64_t start, uint64_t end, double *out);
/* doubleHashSize returned by makeKmerReadDistributionMT (counter.h:300-309) for -m `memory_bytes` */
uint64_t pbk_double_hash_size(uint64_t memory_bytes, uint32_t k);
/* Counter::outputOccurrenceDistribution (counter.h:1000-1007): PREFIX_<k>merFrq.tsv */
int  pbk_write_frq_tsv(const char *path, const uint64_t *occ_hist, uint64_t max_occurrence);
/* loadKmer + outputOccurrenceTableBinary (counter.h:600-640, 955-963; doubleHash.h:266-278):
 * PREFIX_kmer_occ.bin from n (key, count) entries, all of which are written.  Records are placed
 * by DoubleHash probing so that the reference's readTable/find_any (doubleHash.h:280-293, 170-184)
 * finds them.  *load_size_out (may be NULL) receives loadKmer's return value. */
int  pbk_write_kmer_occ_bin(const char *path, uint32_t k, const uint64_t *keys, const uint16_t *counts,
                            uint64_t n, uint64_t double_hash_size, uint64_t *load_size_out);

/* ---- measurement helper: random read-modify-write rate of this GPU's memory system -------------
 * Uniform-random 64-bit key probe + 32-bit atomic add on a table of `table_bytes` (SURVEY.md 8d:
 * R_atomic).  mode 0: red.add only; 1: ld key + red.add (steady-state hit path); 2: CAS-claim
 * inserts of distinct keys; 3-11 and >= 100: 64-bit atomics with return on 8-byte slots, skewed, streamed-key and
 * region-sweep variants (pbk_kernels.cu, scripts/gpu_probe.py name them).  Returns operations per second. */
int  pbk_microbench_atomics(int device, uint64_t table_bytes, uint64_t n_ops, int mode, double *ops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* PBK_H */
