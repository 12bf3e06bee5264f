// pbk_api.cu -- the C ABI of libpbk.so (include/pbk.h): context, streams, batching, table growth.
//
// One context drives one GPU.  Reads arrive in batches (pbk_push_reads); each batch is cut into chunks of
// CHUNK_BASES bases that flow  H2D copy (copy stream, ring of N_STAGE staging buffers)  ->  pack kernel  ->
// counting (compute stream).  Counting takes one of three routes:
//   * small batches (< PART_MIN_WINDOWS windows): count_kernel inserts straight into the table, the host reads a
//     48-byte counter block per chunk to decide whether the table has to grow;
//   * the first large batch of a context: Pass A (partition_kernel) per chunk, then Pass B (bucket_insert_*) in a
//     pilot launch over 1/16 of the buckets that measures the new-key ratio, then the rest (flush_buckets);
//   * later large batches with k <= 32: the table is sized up front from that ratio and Pass A, the device-built
//     tile map and Pass B are chained on the GPU without host round trips -- in up to four groups of chunks when
//     the input comes from the host, so that the passes hide behind the H2D copies (`Pipe`).
// Nothing but counters crosses PCIe back until pbk_finalize / pbk_export.
#include "../../include/pbk.h"
#include "pbk_kernels.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace pbk;

namespace {

constexpr u64 CHUNK_BASES = 1ull << 25;          // 32 Mi bases per pipeline stage (multiple of 32)
constexpr u64 MIN_SLOTS = 1ull << 16;
constexpr double MAX_LOAD = 0.5;                  // k > 32 (16+ byte slots, any capacity)
constexpr double MAX_LOAD_COMPACT = 0.6;          // k <= 32: capacity is a power of two, real load ends up 0.3-0.6
constexpr u64 U32_HEADROOM = (1ull << 32) - 65536 - 2;
constexpr u64 PART_MIN_WINDOWS = 1ull << 22;    // smaller batches go straight to the table
// Second form of Pass B for one-word keys (split_kernel + region_build_kernel: sub-regions of the table built in shared memory,
// no L2 atomics): opt-in, PBK_PASSB2=1 for a context's own bucket store, PBK_PASSB2_GATHER=1 on top for the key / pull exchange.
// Measured on a B200 (profiles/r2m_variants.jsonl): 3.50 ms against 3.72 ms for the first form on device-resident C1, but 13.2
// against 12.5 ms from host buffers (every chunk group reads and rewrites the whole table), for 4.4 GB more store -- not the
// default.  For the exchange, Pass B's time is the NVLink transfer of the keys, which the first form overlaps with its atomics
// tile by tile, while split_gather_kernel would only move them.
constexpr bool PASSB2_DEFAULT = false;
constexpr u64 MAX_PUSH_BASES = 1ull << 31;       // larger pushes are cut into internal batches (2 Gi bases: 16 GiB of bucket store at k <= 32)

enum LaunchClass { LC_PACK = 0, LC_COUNT = 1, LC_OTHER = 2, LC_PART = 3, LC_INSERT = 4, LC_N = 5 };

struct TimedSpan { cudaEvent_t a, b; int cls; };

}  // namespace

constexpr int N_STAGE = 4;                       // H2D staging buffers: the copy stream may run this far ahead of the kernels

struct pbk_ctx {
    int device = 0, sm_count = 148;
    u32 k = 0; int W = 0; u32 flags = 0;
    ShardInfo shard{1, 0};
    cudaStream_t s_compute = nullptr, s_copy = nullptr;
    // side streams (PBK_NO_OVERLAP=1 switches both uses off): s_clear zero-fills the tables at pbk_reset while the next batch's
    // pack and Pass A -- which never touch a table -- already run on s_compute (table_ready() orders the first table access behind
    // it); s_pack runs the pack kernel of chunk i + 1 (memory bound) next to Pass A of chunk i (issue bound) for device-resident input
    cudaStream_t s_clear = nullptr, s_pack = nullptr;
    cudaEvent_t ev_table_clear = nullptr, ev_pack[N_STAGE] = {}, ev_pack_may_start = nullptr;
    bool table_clear_pending = false, overlap_enabled = true;
    u64 budget = 0, used = 0;

    // batch buffers (grow-only)
    u64 *d_stream_raw = nullptr; u32 *d_nflag_raw = nullptr, *d_rflag_raw = nullptr; u64 stream_cap_words = 0;
    u64 *d_offsets = nullptr; u64 offsets_cap = 0;
    uint8_t *d_stage[N_STAGE] = {}; size_t stage_bytes = 0;
    cudaEvent_t ev_copy_done[N_STAGE] = {}, ev_stage_free[N_STAGE] = {};

    TableView table{nullptr, 0, 0}, remote{nullptr, 0, 0};
    u64 occupied = 0, occupied_remote = 0;
    u64 table_hint = 0;
    Counters *d_ctr = nullptr, *h_ctr = nullptr;
    Counters last{};                 // cumulative counters at the last read-back
    u64 *d_ovf = nullptr; u64 ovf_cap = 0;
    // partitioned counting: bucket store filled by Pass A, drained by Pass B
    PartitionPlan plan{};
    u64 *d_bkt_keys = nullptr; size_t bkt_bytes = 0;
    u64 *d_bkt_cursor = nullptr, *h_bkt_cursor = nullptr;
    void *d_passb = nullptr, *h_passb = nullptr;    // Pass B bucket descriptors
    bool partition_enabled = true, partition_forced = false;
    u32 pass = 0;                    // hash-range pass (pbk_config.n_passes >= 2): (n_passes << 16) | pass_index, else 0
    // second form of Pass B (k <= 32, unsharded; split_kernel + region_build_kernel): the sub-region segments and their fill counts
    bool passb2_enabled = false, passb2_gather = false, passb2_fresh = true;
    u64 *d_sub_keys = nullptr; size_t sub_bytes = 0;
    u64 *d_sub_cursor = nullptr; u64 sub_cursor_cap = 0;
    u64 store_windows_ub = 0;        // windows the current bucket store was planned for
    u64 n_passb2 = 0;                // Pass B launches that took the second form
    bool table_touched = true;       // something may have written the main table since it was last zero-filled
    // sub-batched counting of host input (k <= 32): Pass A + Pass B per group of chunks, chained on the GPU, so that
    // only the last group's Pass B is left to do when the last H2D copy lands
    bool pipeline_enabled = true, ratio_known = false;
    u64 n_pipelined = 0;
    u64 *d_len_hist = nullptr, *d_occ_hist = nullptr, *d_shard_counts = nullptr;
    std::vector<u64> h_occ_hist, h_shard_counts;
    u64 stage_gen = 0, shard_counts_gen = ~0ull;   // remote-staging table changes / generation the cached counts belong to
    bool remote_dirty = false;                      // packed but not yet cleared (cleared lazily: a reset usually follows)

    // key exchange (pbk_keyx_*): layout agreed between the ranks, and -- only during a pbk_keyx_partition* call -- the
    // caller's send buffer and cursors that Pass A fills instead of the context's own bucket store
    u64 *d_len_scratch = nullptr; Counters *d_ctr_scratch = nullptr;     // pbk_lookup: its reads must not enter the histograms
    std::vector<u64> seed_rec;                                           // pbk_seed_entries: (key words, value) records, applied by pbk_finalize
    PartitionPlan keyx_plan{}; u64 keyx_max_windows = 0;
    u64 keyx_last_total = 0;            // keys received by the last host-sized insert: sizes the table for the queued ones
    bool counters_pending = false;      // a queued key-exchange insert has not had its counters read back yet (settle())
    u64 pending_new = 0;                // upper bound on the new keys of the inserts queued since the last read-back
    u64 *keyx_send = nullptr, *keyx_cursors = nullptr;
    bool keyx_async = false;            // this partition call returns without reading the counters back
    // pull form of the key exchange (pbk_keyx_pull_*): this context's own owner-major bucket store, double-buffered, in ONE
    // allocation [keys parity 0][keys parity 1][cursors parity 0][cursors parity 1] that the peers map (same process: peer
    // access; other processes: CUDA IPC) and read in place during their Pass B
    char *pull_base = nullptr; size_t pull_bytes = 0, pull_keys_bytes = 0, pull_cur_bytes = 0;
    char *pull_peer[KEYX_MAX_SRC] = {};      // base of every rank's allocation as seen from this GPU ([rank] = own)
    bool pull_peer_ipc[KEYX_MAX_SRC] = {};
    int pull_parity = 1;                     // parity of the store the last partition call filled (first call: 0)
    cudaEvent_t ev_signal = nullptr, ev_wait = nullptr;     // pbk_stream_signal / pbk_stream_wait

    u64 *d_npos_abs = nullptr; u64 npos_abs_cap = 0;      // PBK_ENC_PACKED2: N positions of the current batch (grow-only)
    cudaEvent_t ev_stream_free = nullptr;                 // PBK_ENC_PACKED2: H2D copies go straight into the stream buffer
    void *d_scratch = nullptr; size_t scratch_bytes = 0;   // grow-only arena of pbk_export (entries, sort temporaries)

    bool finalized = false;
    u64 n_reads = 0, n_bases = 0, inst_since_clamp = 0, n_grow = 0;
    double new_ratio = 0.20;         // new keys per window, adapted from what the data shows
    u64 launches[LC_N] = {0, 0, 0, 0, 0};
    double ms[LC_N] = {0, 0, 0, 0, 0};
    std::vector<TimedSpan> spans;
    cudaEvent_t timer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    u64 h2d_bytes = 0, d2h_bytes = 0;
    std::string err;
};

namespace {

const bool g_debug = getenv("PBK_DEBUG") != nullptr;
#define DBG(...) do { if (g_debug) { fprintf(stderr, "[pbk] " __VA_ARGS__); fputc('\n', stderr); fflush(stderr); } } while (0)

double max_load(const pbk_ctx *c)
{
    static const double wide = getenv("PBK_MAX_LOAD_WIDE") ? std::min(0.9, std::max(0.2, atof(getenv("PBK_MAX_LOAD_WIDE")))) : MAX_LOAD;
    return c->W == 1 ? MAX_LOAD_COMPACT : wide;
}

// a hash-range pass (pbk_config.n_passes) keeps 1 / n_passes of a batch's windows: what tables and bucket stores are sized for
u64 pass_share(const pbk_ctx *c, u64 windows)
{
    if (!c->pass) return windows;
    const u64 n = c->pass >> 16;
    return windows / n + windows / (8 * n) + 4096;
}

int fail(pbk_ctx *c, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (c) c->err = buf;
    return code;
}

#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return fail(c, e__ == cudaErrorMemoryAllocation ? PBK_E_NOMEM : PBK_E_CUDA, "%s: %s (%s:%d)", #call, \
                        cudaGetErrorString(e__), __FILE__, __LINE__);                                \
    } while (0)

#define TRY(call) do { int rc__ = (call); if (rc__ != PBK_OK) return rc__; } while (0)

int dev_alloc(pbk_ctx *c, void **p, size_t bytes)
{
    if (bytes == 0) bytes = 8;
    if (c->used + bytes > c->budget)
        return fail(c, PBK_E_NOMEM, "HBM budget exceeded: %.1f MB in use + %.1f MB requested > %.1f MB",
                    c->used / 1e6, bytes / 1e6, c->budget / 1e6);
    CK(cudaMalloc(p, bytes));
    c->used += bytes;
    return PBK_OK;
}

void dev_free(pbk_ctx *c, void *p, size_t bytes)
{
    if (!p) return;
    cudaFree(p);
    if (bytes == 0) bytes = 8;
    c->used -= std::min<u64>(c->used, bytes);
}

struct Span {
    pbk_ctx *c; int cls; TimedSpan t{nullptr, nullptr, 0}; bool on;
    Span(pbk_ctx *c_, int cls_) : c(c_), cls(cls_), on((c_->flags & PBK_F_TIMING) != 0)
    {
        c->launches[cls] += 1;
        if (on) {
            cudaEventCreate(&t.a); cudaEventCreate(&t.b); t.cls = cls;
            cudaEventRecord(t.a, c->s_compute);
        }
    }
    ~Span() { if (on) { cudaEventRecord(t.b, c->s_compute); c->spans.push_back(t); } }
};

void resolve_spans(pbk_ctx *c)
{
    for (auto &s : c->spans) {
        float ms = 0;
        if (cudaEventSynchronize(s.b) == cudaSuccess && cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) c->ms[s.cls] += ms;
        cudaEventDestroy(s.a); cudaEventDestroy(s.b);
    }
    c->spans.clear();
}

// k <= 32: power of two >= 2^27 (compact slots need the count-field headroom); k > 32: any multiple of 1024
u64 round_slots(const pbk_ctx *c, u64 want)
{
    if (c->W == 1) {
        u64 s = 1ull << CT_MIN_QBITS;
        while (s < want) s <<= 1;
        return s;
    }
    return std::max<u64>(MIN_SLOTS, (want + 1023) & ~1023ull);
}

// the tables may still be being zero-filled on the side stream (pbk_reset): everything that touches a table calls this first
int table_ready(pbk_ctx *c)
{
    c->table_touched = true;                         // whoever asks is about to read or write it
    if (!c->table_clear_pending) return PBK_OK;
    c->table_clear_pending = false;
    CK(cudaStreamWaitEvent(c->s_compute, c->ev_table_clear, 0));
    return PBK_OK;
}

// the same ordering for the three Pass B call sites, which decide themselves (at launch time) whether the table may hold
// entries: the flag keeps its value here and is set once their launch is queued
int table_ordered(pbk_ctx *c)
{
    const bool touched = c->table_touched;
    const int rc = table_ready(c);
    c->table_touched = touched;
    return rc;
}

int table_alloc(pbk_ctx *c, TableView *t, u64 slots)
{
    TRY(table_ready(c));
    t->words = c->W; t->cap = slots; t->slots = nullptr;
    TRY(dev_alloc(c, &t->slots, t->bytes()));
    { Span sp(c, LC_OTHER); launch_table_init(*t, c->s_compute); }
    CK(cudaGetLastError());
    if (t == &c->table) c->table_touched = false;
    return PBK_OK;
}

int read_counters(pbk_ctx *c)
{
    CK(cudaMemcpyAsync(c->h_ctr, c->d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, c->s_compute));
    CK(cudaStreamSynchronize(c->s_compute));
    c->d2h_bytes += sizeof(Counters);
    const Counters &n = *c->h_ctr;
    c->occupied += n.new_keys - c->last.new_keys;
    c->occupied_remote += n.new_keys_remote - c->last.new_keys_remote;
    c->inst_since_clamp += n.instances - c->last.instances;
    c->last = n;
    if (n.error_flags & ERR_READ_TOO_LONG) return fail(c, PBK_E_READ_TOO_LONG, "a read has >= 500000 bases");
    if (n.error_flags & ERR_BAD_BASE) return fail(c, PBK_E_BAD_BASE, "input contains a character with no Char2Bin code (only ACGTN, any case)");
    if (n.error_flags & ERR_OVERFLOW_LOST) return fail(c, PBK_E_CUDA, "internal: overflow list exhausted");
    return PBK_OK;
}

int grow_table(pbk_ctx *c, TableView *t, u64 occupied, u64 want_slots)
{
    TRY(table_ready(c));
    const u64 ns = round_slots(c, std::max<u64>(t->cap + t->cap / 2, want_slots));
    DBG("grow table %llu -> %llu slots (occupied %llu)", t->cap, ns, occupied);
    TableView nt{nullptr, ns, c->W};
    TRY(table_alloc(c, &nt, ns));
    { Span sp(c, LC_OTHER); launch_table_rehash(*t, nt, c->d_ctr, c->s_compute); }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->s_compute));
    dev_free(c, t->slots, t->bytes());
    *t = nt;
    c->table_touched = true;
    c->n_grow += 1;
    (void)occupied;
    return PBK_OK;
}

// make room for `expect_new` more keys in the main (and, when sharded, the remote) table
int ensure_room(pbk_ctx *c, u64 expect_new)
{
    u64 local_new = expect_new, remote_new = 0;
    if (c->shard.n_shards > 1) {
        local_new = expect_new / c->shard.n_shards + 1;
        remote_new = expect_new - local_new + 1;
    }
    if ((double)(c->occupied + local_new) > max_load(c) * (double)c->table.capacity())
        TRY(grow_table(c, &c->table, c->occupied, (u64)((c->occupied + local_new) / max_load(c)) + 1));
    if (c->shard.n_shards > 1 && (double)(c->occupied_remote + remote_new) > max_load(c) * (double)c->remote.capacity())
        TRY(grow_table(c, &c->remote, c->occupied_remote, (u64)((c->occupied_remote + remote_new) / max_load(c)) + 1));
    return PBK_OK;
}

int settle(pbk_ctx *c);

int ensure_overflow(pbk_ctx *c, u64 records)
{
    if (records <= c->ovf_cap) return PBK_OK;
    if (c->counters_pending) TRY(settle(c));         // queued launches may still append to the list that is about to be replaced
    if (c->d_ovf) { CK(cudaStreamSynchronize(c->s_compute)); dev_free(c, c->d_ovf, c->ovf_cap * (c->W + 1) * 8); c->d_ovf = nullptr; c->ovf_cap = 0; }
    TRY(dev_alloc(c, (void **)&c->d_ovf, records * (c->W + 1) * 8));
    c->ovf_cap = records;
    return PBK_OK;
}

// Overflow list for a partitioned batch of `windows` windows.  Pass A spills what does not fit its bucket segment, and
// with heavily duplicated input (amplicons: millions of copies of a few k-mers) that can be most of the batch; every
// window spills at most once, so a list as long as the batch cannot be exhausted.  cudaMalloc backs every byte with HBM,
// so that guarantee is only given up to OVF_FULL_RECORDS windows (1 GiB of list at k <= 32); beyond, the list holds a
// quarter of the batch -- far more than any real read set spills (ERR_OVERFLOW_LOST is reported, never silent, if some
// synthetic input does exceed it).  If the HBM budget does not allow even that, the fixed floor stays.
int ensure_overflow_for_batch(pbk_ctx *c, u64 windows)
{
    const u64 floor_records = 1ull << 22, OVF_FULL_RECORDS = 1ull << 26;
    const u64 want = std::max<u64>(floor_records, windows <= OVF_FULL_RECORDS ? windows : std::max<u64>(OVF_FULL_RECORDS, windows / 4));
    if (want <= c->ovf_cap) return PBK_OK;
    const int rc = ensure_overflow(c, want);
    if (rc != PBK_E_NOMEM) return rc;
    c->err.clear();
    return ensure_overflow(c, floor_records);
}

// Spilled keys (no slot within the probe limit, or a full bucket segment in Pass A): insert them from
// a private copy of the list.  The table only grows if it is really loaded, or if a plain retry
// spilled again.
int ensure_tables(pbk_ctx *c, u64 first_batch_windows, bool need_remote);

int drain_overflow(pbk_ctx *c)
{
    if (c->h_ctr->overflow_n > 0) TRY(table_ready(c));
    if (c->h_ctr->overflow_n > 0) TRY(ensure_tables(c, 1024, c->shard.n_shards > 1));   // (key-exchange pushes create no table)
    for (int attempt = 0; c->h_ctr->overflow_n > 0; ++attempt) {
        const u64 n = std::min<u64>(c->h_ctr->overflow_n, c->ovf_cap);
        DBG("drain overflow: %llu records (cap %llu), attempt %d", c->h_ctr->overflow_n, c->ovf_cap, attempt);
        const size_t bytes = n * (c->W + 1) * 8;
        u64 *tmp = nullptr;
        TRY(dev_alloc(c, (void **)&tmp, bytes));
        CK(cudaMemcpyAsync(tmp, c->d_ovf, bytes, cudaMemcpyDeviceToDevice, c->s_compute));
        CK(cudaMemsetAsync(&c->d_ctr->overflow_n, 0, sizeof(u64), c->s_compute));
        c->last.overflow_n = 0;
        int rc = PBK_OK;
        const bool loaded = (double)(c->occupied + n) > max_load(c) * (double)c->table.capacity();
        const bool rloaded = c->shard.n_shards > 1 && (double)(c->occupied_remote + n) > max_load(c) * (double)c->remote.capacity();
        if (loaded || attempt > 0) rc = grow_table(c, &c->table, c->occupied, (u64)((c->occupied + n) / max_load(c)) * 2 + 1);
        if (rc == PBK_OK && (rloaded || (attempt > 0 && c->shard.n_shards > 1)))
            rc = grow_table(c, &c->remote, c->occupied_remote, (u64)((c->occupied_remote + n) / max_load(c)) * 2 + 1);
        if (rc == PBK_OK) {
            Span sp(c, LC_COUNT);
            launch_insert_records(tmp, n, true, c->table, c->remote, c->shard, c->d_ctr, c->d_ovf, c->ovf_cap, c->sm_count, c->s_compute);
            c->table_touched = true;                 // (the table may have been created after the table_ready() above)
        }
        if (rc == PBK_OK) rc = read_counters(c);
        else cudaStreamSynchronize(c->s_compute);
        dev_free(c, tmp, bytes);
        TRY(rc);
    }
    return PBK_OK;
}

// pbk_keyx_insert_device may queue its launches and return (no host round trip per received chunk); whoever needs the
// host's view of the table (occupancy, staged records, the overflow list) calls this first
int settle(pbk_ctx *c)
{
    TRY(table_ready(c));
    if (!c->counters_pending) return PBK_OK;
    c->counters_pending = false;
    c->pending_new = 0;
    TRY(read_counters(c));
    return drain_overflow(c);
}

int maybe_clamp(pbk_ctx *c, u64 upcoming)
{
    if (c->inst_since_clamp + upcoming < U32_HEADROOM) return PBK_OK;
    TRY(table_ready(c));
    { Span sp(c, LC_OTHER); launch_table_clamp(c->table, c->s_compute); }
    if (c->shard.n_shards > 1 && c->remote.slots) { Span sp(c, LC_OTHER); launch_table_clamp(c->remote, c->s_compute); }
    CK(cudaGetLastError());
    c->inst_since_clamp = 0;
    return PBK_OK;
}

int ensure_tables(pbk_ctx *c, u64 first_batch_windows, bool need_remote)
{
    u64 want = c->table_hint ? c->table_hint : (u64)(first_batch_windows * c->new_ratio / max_load(c)) + 1;
    if (need_remote && c->shard.n_shards > 1 && !c->remote.slots)
        TRY(table_alloc(c, &c->remote, round_slots(c, want)));     // remote-staging table holds (n-1)/n of the keys
    if (c->table.slots) return PBK_OK;
    if (c->shard.n_shards > 1) want = want / c->shard.n_shards + 1;
    TRY(table_alloc(c, &c->table, round_slots(c, want)));
    return PBK_OK;
}

int ensure_batch_buffers(pbk_ctx *c, u64 n_bases, u64 n_reads)
{
    const u64 words = (n_bases + 31) / 32;
    if (words > c->stream_cap_words) {
        CK(cudaStreamSynchronize(c->s_compute));
        const u64 old = c->stream_cap_words ? c->stream_cap_words + STREAM_PAD_WORDS : 0;
        dev_free(c, c->d_stream_raw, old * 8); dev_free(c, c->d_nflag_raw, old * 4); dev_free(c, c->d_rflag_raw, old * 4);
        c->d_stream_raw = nullptr; c->d_nflag_raw = c->d_rflag_raw = nullptr; c->stream_cap_words = 0;
        const u64 cap = words + words / 8 + 1024, tot = cap + STREAM_PAD_WORDS;
        TRY(dev_alloc(c, (void **)&c->d_stream_raw, tot * 8));
        TRY(dev_alloc(c, (void **)&c->d_nflag_raw, tot * 4));
        TRY(dev_alloc(c, (void **)&c->d_rflag_raw, tot * 4));
        CK(cudaMemsetAsync(c->d_stream_raw, 0, STREAM_PAD_WORDS * 8, c->s_compute));
        CK(cudaMemsetAsync(c->d_nflag_raw, 0, STREAM_PAD_WORDS * 4, c->s_compute));
        CK(cudaMemsetAsync(c->d_rflag_raw, 0, STREAM_PAD_WORDS * 4, c->s_compute));
        c->stream_cap_words = cap;
    }
    if (n_reads + 1 > c->offsets_cap) {
        CK(cudaStreamSynchronize(c->s_compute));
        dev_free(c, c->d_offsets, c->offsets_cap * 8);
        c->d_offsets = nullptr; c->offsets_cap = 0;
        const u64 cap = n_reads + n_reads / 8 + 1024;
        TRY(dev_alloc(c, (void **)&c->d_offsets, cap * 8));
        c->offsets_cap = cap;
    }
    return PBK_OK;
}

// count the windows ending in stream words [w0, w1) of the current batch; grows the table as needed
int count_range(pbk_ctx *c, u64 w0, u64 w1)
{
    if (w1 <= w0) return PBK_OK;
    TRY(table_ready(c));
    const u64 windows = (w1 - w0) * 32;
    TRY(maybe_clamp(c, windows));
    TRY(ensure_room(c, (u64)(pass_share(c, windows) * std::min(1.0, c->new_ratio * 1.25))));
    TRY(ensure_overflow(c, pass_share(c, windows)));   // (only keys that are counted can spill)
    const u64 before_new = c->occupied + c->occupied_remote, before_inst = c->last.instances;
    {
        Span sp(c, LC_COUNT);
        launch_count(c->d_stream_raw + STREAM_PAD_WORDS, c->d_nflag_raw + STREAM_PAD_WORDS, c->d_rflag_raw + STREAM_PAD_WORDS,
                     w0, w1, (int)c->k, c->table, c->remote, c->shard, c->d_ctr, c->d_ovf, c->ovf_cap, c->sm_count, c->s_compute, c->pass);
        c->table_touched = true;
    }
    CK(cudaGetLastError());
    TRY(read_counters(c));
    TRY(drain_overflow(c));
    const u64 inst = c->last.instances - before_inst, newk = c->occupied + c->occupied_remote - before_new;
    if (inst > 4096) c->new_ratio = std::max(0.01, std::min(1.0, (double)newk / (double)inst));
    return PBK_OK;
}

// ---- partitioned path ---------------------------------------------------------------------------

// cursors and Pass B descriptors, device + pinned host copies
int ensure_passb_buffers(pbk_ctx *c)
{
    if (c->d_bkt_cursor) return PBK_OK;
    TRY(dev_alloc(c, (void **)&c->d_bkt_cursor, PART_MAX_BUCKETS * 8));
    if (cudaMallocHost((void **)&c->h_bkt_cursor, PART_MAX_BUCKETS * 8) != cudaSuccess) return fail(c, PBK_E_NOMEM, "pinned host memory");
    TRY(dev_alloc(c, (void **)&c->d_passb, passb_desc_bytes(PART_MAX_BUCKETS)));
    if (cudaMallocHost((void **)&c->h_passb, passb_desc_bytes(PART_MAX_BUCKETS)) != cudaSuccess) return fail(c, PBK_E_NOMEM, "pinned host memory");
    return PBK_OK;
}

// bucket store for `windows_ub` windows (the whole batch, or one sub-batch); the table is sized for `windows_total`
int prepare_partition(pbk_ctx *c, u64 windows_ub, u64 windows_total)
{
    const u64 est_slots = round_slots(c, std::max<u64>(c->table.slots ? c->table.cap : 0,
                                                       (u64)((c->occupied + windows_total * c->new_ratio) / max_load(c))));
    TableView tv{nullptr, est_slots, c->W};
    c->plan = plan_partition(tv.bytes(), windows_ub, c->W);
    c->store_windows_ub = windows_ub;
    const size_t need = (size_t)c->plan.n_buckets * c->plan.seg_cap * c->W * 8;
    if (need > c->bkt_bytes) {
        CK(cudaStreamSynchronize(c->s_compute));
        dev_free(c, c->d_bkt_keys, c->bkt_bytes);
        c->d_bkt_keys = nullptr; c->bkt_bytes = 0;
        TRY(dev_alloc(c, (void **)&c->d_bkt_keys, need));
        c->bkt_bytes = need;
    }
    TRY(ensure_passb_buffers(c));
    CK(cudaMemsetAsync(c->d_bkt_cursor, 0, PART_MAX_BUCKETS * 8, c->s_compute));
    // a spilled key (bucket segment or bin full) is rare; the list is also used by Pass B
    TRY(ensure_overflow_for_batch(c, windows_total));
    return PBK_OK;
}

// Pass B, second form (PBK_PASSB2; k <= 32): the keys go through split_kernel into per-sub-region segments and
// region_build_kernel builds every sub-region of the table in shared memory.
//   srcs == nullptr  buckets [d0, d1) of the context's own bucket store (unsharded table)
//   srcs             key exchange: descriptors [d0, d1) of the n_src x n_regions (region-major) sources -- the all-to-all
//                    receive buffer, or the peers' stores read in place over NVLink (pull form); d0, d1 multiples of n_src
// If nothing has touched the table since it was zero-filled (c->table_touched) the sub-regions are not read.  Returns 1 if the
// launches were queued, 0 if this table / plan cannot take the route (the caller runs the first form), a negative status on error.
int passb2_run(pbk_ctx *c, const KeyxSources *srcs, u32 d0, u32 d1)
{
    const bool was_touched = c->table_touched;
    if (!c->passb2_enabled || c->W != 1 || d1 <= d0) return 0;
    if (!srcs && c->shard.n_shards > 1) return 0;       // (record exchange: foreign keys go to the remote-staging table)
    if (srcs && !c->passb2_gather) return 0;
    const u32 G = srcs ? c->shard.n_shards : 1u;
    const u32 n_desc = srcs ? c->keyx_plan.n_buckets : c->plan.n_buckets, n_regions = n_desc / G;
    const u64 seg_cap = srcs ? c->keyx_plan.seg_cap : c->plan.seg_cap;
    Passb2Geom geom;
    if (!passb2_geom(c->table, n_regions, &geom) || d0 % G || d1 % G) return 0;
    const u64 sub_cap = passb2_sub_cap(srcs ? c->keyx_max_windows : c->store_windows_ub, geom.n_sub);
    const size_t need = (size_t)geom.n_sub * sub_cap * 8;
    if (need > c->sub_bytes || geom.n_sub > c->sub_cursor_cap) {
        CK(cudaStreamSynchronize(c->s_compute));
        dev_free(c, c->d_sub_keys, c->sub_bytes); dev_free(c, c->d_sub_cursor, c->sub_cursor_cap * 8);
        c->d_sub_keys = c->d_sub_cursor = nullptr; c->sub_bytes = 0; c->sub_cursor_cap = 0;
        if (dev_alloc(c, (void **)&c->d_sub_keys, need) != PBK_OK || dev_alloc(c, (void **)&c->d_sub_cursor, geom.n_sub * 8) != PBK_OK) {
            DBG("pass B second form: no room for %.1f MB of sub-region segments, first form from here on", need / 1e6);
            dev_free(c, c->d_sub_keys, need);
            c->d_sub_keys = nullptr; c->err.clear(); c->passb2_enabled = false;
            return 0;
        }
        c->sub_bytes = need; c->sub_cursor_cap = geom.n_sub;
    }
    DBG("pass B second form: descriptors [%u,%u) of %u (%u sources), %u sub-regions per region, sub_cap %llu, %s", d0, d1, n_desc, G, geom.F,
        (unsigned long long)sub_cap, was_touched ? "table read" : "table known empty");
    CK(cudaMemsetAsync(c->d_sub_cursor, 0, geom.n_sub * 8, c->s_compute));
    {
        Span sp(c, LC_OTHER);
        if (srcs) launch_passb2_desc_gather(*srcs, seg_cap, G, n_regions, c->d_passb, c->s_compute);
        else launch_passb2_desc(c->d_bkt_cursor, seg_cap, n_desc, c->d_passb, c->s_compute);
    }
    CK(cudaGetLastError());
    {
        Span sp(c, LC_INSERT);
        launch_passb2_split(c->d_bkt_keys, srcs, G, seg_cap, c->d_passb, d0, d1, geom, c->d_sub_keys, sub_cap, c->d_sub_cursor, c->d_ctr,
                            c->d_ovf, c->ovf_cap, c->sm_count, c->s_compute);
    }
    CK(cudaGetLastError());
    {
        Span sp(c, LC_INSERT);
        launch_passb2_build(c->d_sub_keys, sub_cap, c->d_sub_cursor, d0 / G, d1 / G, geom, c->table, was_touched || !c->passb2_fresh, c->d_ctr,
                            c->d_ovf, c->ovf_cap, c->sm_count, c->s_compute);
    }
    CK(cudaGetLastError());
    c->n_passb2 += 1;
    return 1;
}

// ---- sub-batched path (k <= 32, host input): the chunks of a batch are counted in groups -- Pass A per chunk, then
// a device-built tile map and Pass B for the group, all queued on the compute stream without a host round trip --
// while the copy stream keeps feeding the staging buffers.  When the last copy lands only one group's Pass B is
// left.  (Running Pass B on a second stream next to Pass A was measured and is slower: both passes are bound by
// the SM's load/store path -- shared-memory atomics in A, 32-sector global atomics in B -- not by different units.)
struct Pipe {
    bool on = false;
    u32 sizes[8] = {0}, n_sb = 0, cur = 0, in_sb = 0;   // chunks per group, current group, chunks already partitioned into it
    u32 sb_chunks() const { return sizes[cur < n_sb ? cur : n_sb - 1]; }
};

int pipe_finish_subbatch(pbk_ctx *c, Pipe &p)
{
    if (p.in_sb == 0) return PBK_OK;
    TRY(table_ordered(c));                           // Pass B is the first thing of a batch that touches the table
    const int second = passb2_run(c, nullptr, 0, c->plan.n_buckets);
    if (second < 0) return second;
    c->table_touched = true;
    if (second == 0) {
        { Span sp(c, LC_OTHER); launch_passb_desc(c->d_bkt_cursor, c->plan.seg_cap, c->plan.n_buckets, c->table, c->remote, c->shard, c->d_passb, c->s_compute); }
        CK(cudaGetLastError());
        {
            Span sp(c, LC_INSERT);
            launch_bucket_insert_chained(c->d_bkt_keys, c->plan.seg_cap, c->d_passb, c->plan.n_buckets, c->table, c->remote, c->shard, c->d_ctr,
                                         c->d_ovf, c->ovf_cap, c->sm_count, 3, c->s_compute);
        }
        CK(cudaGetLastError());
    }
    CK(cudaMemsetAsync(c->d_bkt_cursor, 0, PART_MAX_BUCKETS * 8, c->s_compute));
    p.in_sb = 0;
    p.cur += 1;
    return PBK_OK;
}

int pipe_end(pbk_ctx *c, Pipe &p)
{
    TRY(pipe_finish_subbatch(c, p));
    const u64 occ_before = c->occupied + c->occupied_remote, inst_before = c->last.instances;
    TRY(read_counters(c));
    TRY(drain_overflow(c));
    const u64 inst = c->last.instances - inst_before;
    if (inst > 4096) c->new_ratio = std::max(0.01, std::min(1.0, (double)(c->occupied + c->occupied_remote - occ_before) / (double)inst));
    c->n_pipelined += 1;
    return PBK_OK;
}

// one Pass B launch over buckets [b0, b1)
int passb_launch(pbk_ctx *c, u32 b0, u32 b1)
{
    TRY(table_ordered(c));
    u64 total = 0;
    for (u32 b = b0; b < b1; ++b) total += c->h_bkt_cursor[b];
    DBG("pass B buckets [%u,%u) of %u: %llu keys, table %llu slots, occupied %llu", b0, b1, c->plan.n_buckets, total, c->table.cap, c->occupied);
    const int second = passb2_run(c, nullptr, b0, b1);
    if (second < 0) return second;
    c->table_touched = true;
    if (second == 0) {
        Span sp(c, LC_INSERT);
        launch_bucket_insert(c->d_bkt_keys, c->plan.seg_cap, c->h_bkt_cursor, c->h_passb, c->d_passb, b0, b1, c->plan.n_buckets,
                             c->table, c->remote, c->shard, c->d_ctr, c->d_ovf, c->ovf_cap, c->sm_count, c->s_compute);
    }
    CK(cudaGetLastError());
    TRY(read_counters(c));                          // also makes h_passb reusable
    TRY(drain_overflow(c));
    return PBK_OK;
}

// Pass B: drain the bucket store into the table, one L2-sized hash range at a time
int flush_buckets(pbk_ctx *c)
{
    const u32 P = c->plan.n_buckets;
    CK(cudaMemcpyAsync(c->h_bkt_cursor, c->d_bkt_cursor, P * 8, cudaMemcpyDeviceToHost, c->s_compute));
    TRY(read_counters(c));                          // also picks up instances counted by Pass A
    TRY(drain_overflow(c));
    c->d2h_bytes += P * 8;
    u64 total = 0;
    for (u32 b = 0; b < P; ++b) { c->h_bkt_cursor[b] = std::min<u64>(c->h_bkt_cursor[b], c->plan.seg_cap); total += c->h_bkt_cursor[b]; }
    DBG("pass A done: %llu keys in %u buckets (seg_cap %llu, bin_cap %u, threads %d), instances %llu", total, P, c->plan.seg_cap, c->plan.bin_cap, c->plan.threads, c->last.instances);
    if (total == 0) return PBK_OK;
    TRY(maybe_clamp(c, total));
    const u64 occ_before = c->occupied + c->occupied_remote;
    // pilot: buckets are statistically identical hash ranges, so what the first 1/16 of them adds
    // predicts the rest; size the table once, then run everything else in a single launch
    const u32 pilot = std::max<u32>(1, P / 16);
    u64 pilot_keys = 0;
    for (u32 b = 0; b < pilot; ++b) pilot_keys += c->h_bkt_cursor[b];
    // the pilot fills only its own hash ranges, so the table must already have the size the whole batch needs
    TRY(ensure_room(c, (u64)(total * std::min(1.0, c->new_ratio))));
    TRY(passb_launch(c, 0, pilot));
    double per_key = pilot_keys ? (double)(c->occupied + c->occupied_remote - occ_before) / (double)pilot_keys : c->new_ratio;
    if (pilot < P) {
        TRY(ensure_room(c, (u64)((double)(total - pilot_keys) * std::min(1.0, per_key * 1.05)) + 4096));
        TRY(passb_launch(c, pilot, P));
    }
    if (total > 4096) {
        c->new_ratio = std::max(0.01, std::min(1.0, (double)(c->occupied + c->occupied_remote - occ_before) / (double)total));
        c->ratio_known = true;
    }
    CK(cudaMemsetAsync(c->d_bkt_cursor, 0, PART_MAX_BUCKETS * 8, c->s_compute));
    return PBK_OK;
}

int push_common(pbk_ctx *c, const uint8_t *h_bases, const uint8_t *d_bases_in, const u64 *h_offsets,
                const u64 *d_offsets_in, u64 n_reads, u64 n_bases, int encoding, const int32_t *n_pos,
                const u64 *n_pos_offsets)
{
    if (c->finalized) return fail(c, PBK_E_STATE, "pbk_push_reads after pbk_finalize (call pbk_reset first)");
    if (n_reads == 0) return PBK_OK;
    CK(cudaSetDevice(c->device));
    if (!c->keyx_send && c->counters_pending) TRY(settle(c));   // (a key-exchange partition only runs Pass A; its own read-back at the end settles)
    // (no table_ready() here: read marks, pack and Pass A do not touch the tables, which may still be being cleared on s_clear)
    c->stage_gen += 1;
    if (c->remote_dirty) { TRY(table_ready(c)); Span sp(c, LC_OTHER); launch_table_init(c->remote, c->s_compute); c->remote_dirty = false; }
    TRY(ensure_batch_buffers(c, n_bases, n_reads));
    const u64 windows_ub = n_bases > (u64)n_reads * (c->k - 1) ? n_bases - (u64)n_reads * (c->k - 1) : 0;
    const bool keyx = c->keyx_send != nullptr;       // Pass A only, into the caller's all-to-all send buffer
    if (keyx && windows_ub > c->keyx_max_windows)
        return fail(c, PBK_E_ARG, "batch has up to %llu windows, the key-exchange layout was planned for %llu",
                    (unsigned long long)windows_ub, (unsigned long long)c->keyx_max_windows);
    if (!keyx) TRY(ensure_tables(c, std::max<u64>(pass_share(c, windows_ub), 1024), true));

    u64 *stream = c->d_stream_raw + STREAM_PAD_WORDS;
    u32 *nflag = c->d_nflag_raw + STREAM_PAD_WORDS, *rflag = c->d_rflag_raw + STREAM_PAD_WORDS;
    const u64 words = (n_bases + 31) / 32;

    // read starts + length histogram
    const u64 *d_off = d_offsets_in;
    if (!d_off) {
        CK(cudaMemcpyAsync(c->d_offsets, h_offsets, (n_reads + 1) * 8, cudaMemcpyHostToDevice, c->s_compute));
        c->h2d_bytes += (n_reads + 1) * 8;
        d_off = c->d_offsets;
    }
    CK(cudaMemsetAsync(rflag, 0, words * 4, c->s_compute));
    { Span sp(c, LC_OTHER); launch_read_marks(d_off, n_reads, c->d_len_hist, rflag, c->d_ctr, c->s_compute); }
    CK(cudaGetLastError());

    const bool deferred_count = (encoding == PBK_ENC_PLATANUS);     // N flags arrive after all packs
    const bool partitioned = keyx || (c->partition_enabled && (windows_ub >= PART_MIN_WINDOWS || (c->partition_forced && windows_ub > 0)));
    // Chained (new-key ratio known from an earlier batch of this context, so the table can be sized up front;
    // the first large batch takes the other path with its pilot launch): Pass A per chunk, then a device-built tile map
    // and Pass B, all queued without a host round trip.  Host input is counted in up to four such groups of chunks so
    // that only the last group's Pass B is left when the last H2D copy lands; device-resident input in one group.
    const u64 n_chunks_total = (n_bases + CHUNK_BASES - 1) / CHUNK_BASES;
    Pipe pipe;
    static const bool wide_pipe = !(getenv("PBK_WIDE_PIPE") && atoi(getenv("PBK_WIDE_PIPE")) == 0);
    pipe.on = partitioned && !keyx && c->pipeline_enabled && (c->W == 1 || wide_pipe) && c->ratio_known;
    if (pipe.on) {
        const u32 max_sb = getenv("PBK_N_SB") ? (u32)std::min(8, std::max(1, atoi(getenv("PBK_N_SB")))) : 4u;
        // (multi-word keys: every group's Pass B sweeps a table of 32-byte slots once more, so at most two groups)
        const u32 cap_sb = c->W == 1 ? max_sb : std::min<u32>(max_sb, 2u);
        const u32 n_sb = (h_bases != nullptr && n_chunks_total >= 4) ? (u32)std::max<u64>(1, std::min<u64>(cap_sb, n_chunks_total / 2)) : 1u;
        // host input: groups of equal size (a schedule that ends with a single-chunk group was measured: 2 % slower)
        pipe.n_sb = n_sb;
        for (u32 i = 0; i < n_sb; ++i) pipe.sizes[i] = (u32)((n_chunks_total + n_sb - 1) / n_sb);
        u32 largest = 1;
        for (u32 i = 0; i < n_sb; ++i) largest = std::max(largest, pipe.sizes[i]);
        const u64 sb_windows = std::min<u64>(windows_ub, (u64)largest * CHUNK_BASES);
        TRY(maybe_clamp(c, windows_ub));
        TRY(ensure_room(c, (u64)(pass_share(c, windows_ub) * std::min(1.0, c->new_ratio * 1.15)) + 65536));
        TRY(prepare_partition(c, pass_share(c, sb_windows), pass_share(c, windows_ub)));
    } else if (keyx) {
        c->plan = c->keyx_plan;
        CK(cudaMemsetAsync(c->keyx_cursors, 0, (size_t)c->plan.n_buckets * 8, c->s_compute));
        TRY(ensure_overflow_for_batch(c, windows_ub));
    } else if (partitioned) {
        TRY(prepare_partition(c, pass_share(c, windows_ub), pass_share(c, windows_ub)));
    }
    u64 *const bkt_keys = keyx ? c->keyx_send : c->d_bkt_keys, *const bkt_cursor = keyx ? c->keyx_cursors : c->d_bkt_cursor;
    // large batches: Pass A per chunk (no host sync), Pass B once at the end.  Small ones: straight to the table.
    auto count_words = [&](u64 w0, u64 w1) -> int {
        if (!partitioned) return count_range(c, w0, w1);
        {
            Span sp(c, LC_PART);
            launch_partition(stream, nflag, rflag, w0, w1, (int)c->k, c->W, c->plan, bkt_keys, bkt_cursor, c->d_ctr,
                             c->d_ovf, c->ovf_cap, c->sm_count, c->s_compute, keyx ? c->shard.n_shards : 1u, c->pass);
        }
        if (pipe.on && ++pipe.in_sb >= pipe.sb_chunks()) TRY(pipe_finish_subbatch(c, pipe));
        return PBK_OK;
    };
    if (d_bases_in) {
        // inputs already in HBM: pack chunk by chunk so the chunk's stream words are still in L2 when counted.  With Pass A behind
        // it (partitioned batches) the pack kernel of a chunk runs on its own stream, so that it overlaps Pass A of the chunk before:
        // one is bound by memory, the other by instruction issue.  (Not with per-launch timing: the spans bracket s_compute.)
        const bool side = c->overlap_enabled && partitioned && !(c->flags & PBK_F_TIMING);
        if (side) {
            CK(cudaEventRecord(c->ev_pack_may_start, c->s_compute));       // whatever still reads the previous batch's stream comes first
            CK(cudaStreamWaitEvent(c->s_pack, c->ev_pack_may_start, 0));
        }
        u64 ci = 0;
        for (u64 b0 = 0; b0 < n_bases; b0 += CHUNK_BASES, ++ci) {
            const u64 nb = std::min(CHUNK_BASES, n_bases - b0), w0 = b0 / 32, nw = (nb + 31) / 32;
            const int mode = encoding | ((c->flags & PBK_F_UNKNOWN_AS_N) ? 0x100 : 0);
            if (side) {
                c->launches[LC_PACK] += 1;
                launch_pack(d_bases_in + b0, nb, nw, mode, stream, nflag, w0, c->d_ctr, c->s_pack);
                CK(cudaEventRecord(c->ev_pack[ci % N_STAGE], c->s_pack));
                CK(cudaStreamWaitEvent(c->s_compute, c->ev_pack[ci % N_STAGE], 0));
            } else {
                Span sp(c, LC_PACK);
                launch_pack(d_bases_in + b0, nb, nw, mode, stream, nflag, w0, c->d_ctr, c->s_compute);
            }
            CK(cudaGetLastError());
            TRY(count_words(w0, w0 + nw));
        }
    } else if (encoding == PBK_ENC_PACKED2) {
        // the host hands over 2-bit words in the stream's own layout: chunks are copied straight into the stream buffer on the
        // copy stream (no staging ring, no pack kernel); n_pos carries the batch's absolute N positions, n_pos_offsets[0] their number
        const u64 *h_words = reinterpret_cast<const u64 *>(h_bases);
        const u64 *h_npos = reinterpret_cast<const u64 *>(n_pos);
        const u64 n_n = n_pos_offsets ? n_pos_offsets[0] : 0;
        for (int i = 0; i < N_STAGE; ++i)
            if (!c->ev_copy_done[i]) CK(cudaEventCreateWithFlags(&c->ev_copy_done[i], cudaEventDisableTiming));
        if (!c->ev_stream_free) CK(cudaEventCreateWithFlags(&c->ev_stream_free, cudaEventDisableTiming));
        if (n_n > c->npos_abs_cap) {
            CK(cudaStreamSynchronize(c->s_compute));
            dev_free(c, c->d_npos_abs, c->npos_abs_cap * 8);
            c->d_npos_abs = nullptr; c->npos_abs_cap = 0;
            TRY(dev_alloc(c, (void **)&c->d_npos_abs, (n_n + n_n / 4 + 1024) * 8));
            c->npos_abs_cap = n_n + n_n / 4 + 1024;
        }
        CK(cudaMemsetAsync(nflag, 0, words * 4, c->s_compute));
        if (n_n) { CK(cudaMemcpyAsync(c->d_npos_abs, h_npos, n_n * 8, cudaMemcpyHostToDevice, c->s_compute)); c->h2d_bytes += n_n * 8; }
        { Span sp(c, LC_OTHER); launch_npos_abs_scatter(c->d_npos_abs, n_n, n_bases, nflag, c->s_compute); }
        CK(cudaGetLastError());
        CK(cudaEventRecord(c->ev_stream_free, c->s_compute));          // whatever still read the previous batch's stream is queued before this
        CK(cudaStreamWaitEvent(c->s_copy, c->ev_stream_free, 0));
        const u64 n_chunks = (n_bases + CHUNK_BASES - 1) / CHUNK_BASES;
        auto enqueue_copy = [&](u64 ci) -> int {
            const u64 w0 = ci * (CHUNK_BASES / 32), nw = std::min<u64>(CHUNK_BASES / 32, words - w0);
            CK(cudaMemcpyAsync(stream + w0, h_words + w0, nw * 8, cudaMemcpyHostToDevice, c->s_copy));
            CK(cudaEventRecord(c->ev_copy_done[ci % N_STAGE], c->s_copy));
            c->h2d_bytes += nw * 8;
            return PBK_OK;
        };
        for (u64 ci = 0; ci < std::min<u64>(n_chunks, N_STAGE - 1); ++ci) TRY(enqueue_copy(ci));
        for (u64 ci = 0; ci < n_chunks; ++ci) {
            const u64 w0 = ci * (CHUNK_BASES / 32), nw = std::min<u64>(CHUNK_BASES / 32, words - w0);
            if (ci + N_STAGE - 1 < n_chunks) TRY(enqueue_copy(ci + N_STAGE - 1));
            CK(cudaStreamWaitEvent(c->s_compute, c->ev_copy_done[ci % N_STAGE], 0));
            TRY(count_words(w0, w0 + nw));
        }
    } else {
        if (!c->d_stage[0]) {
            for (int i = 0; i < N_STAGE; ++i) {
                TRY(dev_alloc(c, (void **)&c->d_stage[i], CHUNK_BASES));
                if (!c->ev_copy_done[i]) CK(cudaEventCreateWithFlags(&c->ev_copy_done[i], cudaEventDisableTiming));
                CK(cudaEventCreateWithFlags(&c->ev_stage_free[i], cudaEventDisableTiming));
            }
            c->stage_bytes = CHUNK_BASES;
        }
        const u64 n_chunks = (n_bases + CHUNK_BASES - 1) / CHUNK_BASES;
        auto enqueue_copy = [&](u64 ci) -> int {
            const int buf = (int)(ci % N_STAGE);
            const u64 b0 = ci * CHUNK_BASES, nb = std::min(CHUNK_BASES, n_bases - b0);
            CK(cudaStreamWaitEvent(c->s_copy, c->ev_stage_free[buf], 0));
            CK(cudaMemcpyAsync(c->d_stage[buf], h_bases + b0, nb, cudaMemcpyHostToDevice, c->s_copy));
            CK(cudaEventRecord(c->ev_copy_done[buf], c->s_copy));
            c->h2d_bytes += nb;
            return PBK_OK;
        };
        for (u64 ci = 0; ci < std::min<u64>(n_chunks, N_STAGE - 1); ++ci) TRY(enqueue_copy(ci));
        for (u64 ci = 0; ci < n_chunks; ++ci) {
            const int buf = (int)(ci % N_STAGE);
            const u64 b0 = ci * CHUNK_BASES, nb = std::min(CHUNK_BASES, n_bases - b0), w0 = b0 / 32, nw = (nb + 31) / 32;
            if (ci + N_STAGE - 1 < n_chunks) TRY(enqueue_copy(ci + N_STAGE - 1));   // copies run ahead of this chunk's kernels
            CK(cudaStreamWaitEvent(c->s_compute, c->ev_copy_done[buf], 0));
            { Span sp(c, LC_PACK); launch_pack(c->d_stage[buf], nb, nw, encoding | ((c->flags & PBK_F_UNKNOWN_AS_N) ? 0x100 : 0), stream, nflag, w0, c->d_ctr, c->s_compute); }
            CK(cudaGetLastError());
            CK(cudaEventRecord(c->ev_stage_free[buf], c->s_compute));
            if (!deferred_count) TRY(count_words(w0, w0 + nw));
        }
    }
    if (deferred_count) {
        if (!n_pos || !n_pos_offsets) return fail(c, PBK_E_ARG, "PBK_ENC_PLATANUS needs n_pos and n_pos_offsets");
        const u64 total_n = n_pos_offsets[n_reads];
        int32_t *d_np = nullptr; u64 *d_npo = nullptr;
        TRY(dev_alloc(c, (void **)&d_np, total_n * 4));
        int rc = dev_alloc(c, (void **)&d_npo, (n_reads + 1) * 8);
        if (rc == PBK_OK) {
            cudaMemcpyAsync(d_np, n_pos, total_n * 4, cudaMemcpyHostToDevice, c->s_compute);
            cudaMemcpyAsync(d_npo, n_pos_offsets, (n_reads + 1) * 8, cudaMemcpyHostToDevice, c->s_compute);
            c->h2d_bytes += total_n * 4 + (n_reads + 1) * 8;
            { Span sp(c, LC_OTHER); launch_npos_scatter(d_off, d_np, d_npo, n_reads, nflag, c->s_compute); }
            for (u64 w0 = 0; w0 < words && rc == PBK_OK; w0 += CHUNK_BASES / 32)
                rc = count_words(w0, std::min(words, w0 + CHUNK_BASES / 32));
        }
        cudaStreamSynchronize(c->s_compute);
        dev_free(c, d_np, total_n * 4); dev_free(c, d_npo, (n_reads + 1) * 8);
        TRY(rc);
    }
    if (keyx) {
        // the keys stay in the caller's buffer; what the host needs now: error flags, the instance count, and the few
        // keys that found their segment full (a heavily repeated k-mer), which take the record route (remote-staging
        // table -> pbk_shard_pack_device) to their owner
        CK(cudaGetLastError());
        if (c->keyx_async) {
            c->counters_pending = true;                 // settle() reads them (and drains spills) when the host next needs them
        } else {
            c->counters_pending = false;                // the read-back below also covers a queued pbk_keyx_insert_device
            TRY(read_counters(c));
            TRY(drain_overflow(c));
        }
    } else if (pipe.on) {
        CK(cudaGetLastError());
        TRY(pipe_end(c, pipe));
    } else if (partitioned) {
        CK(cudaGetLastError());
        TRY(flush_buckets(c));
    }
    c->n_reads += n_reads;
    c->n_bases += n_bases;
    return PBK_OK;
}

void release_all(pbk_ctx *c)
{
    cudaSetDevice(c->device);
    if (c->s_compute) cudaStreamSynchronize(c->s_compute);
    if (c->s_copy) cudaStreamSynchronize(c->s_copy);
    if (c->s_clear) { cudaStreamSynchronize(c->s_clear); cudaStreamDestroy(c->s_clear); }
    if (c->s_pack) { cudaStreamSynchronize(c->s_pack); cudaStreamDestroy(c->s_pack); }
    if (c->ev_table_clear) cudaEventDestroy(c->ev_table_clear);
    if (c->ev_pack_may_start) cudaEventDestroy(c->ev_pack_may_start);
    for (int i = 0; i < N_STAGE; ++i) if (c->ev_pack[i]) cudaEventDestroy(c->ev_pack[i]);
    resolve_spans(c);
    cudaFree(c->d_stream_raw); cudaFree(c->d_nflag_raw); cudaFree(c->d_rflag_raw); cudaFree(c->d_offsets);
    for (int i = 0; i < N_STAGE; ++i) {
        cudaFree(c->d_stage[i]);
        if (c->ev_copy_done[i]) cudaEventDestroy(c->ev_copy_done[i]);
        if (c->ev_stage_free[i]) cudaEventDestroy(c->ev_stage_free[i]);
    }
    cudaFree(c->d_bkt_keys); cudaFree(c->d_bkt_cursor); if (c->h_bkt_cursor) cudaFreeHost(c->h_bkt_cursor);
    cudaFree(c->d_sub_keys); cudaFree(c->d_sub_cursor);
    cudaFree(c->d_passb); if (c->h_passb) cudaFreeHost(c->h_passb);
    cudaFree(c->table.slots); cudaFree(c->remote.slots); cudaFree(c->d_ctr); cudaFree(c->d_ovf);
    cudaFree(c->d_len_hist); cudaFree(c->d_occ_hist); cudaFree(c->d_shard_counts);
    cudaFree(c->d_len_scratch); cudaFree(c->d_ctr_scratch); cudaFree(c->d_scratch); cudaFree(c->d_npos_abs);
    if (c->ev_stream_free) cudaEventDestroy(c->ev_stream_free);
    for (int i = 0; i < KEYX_MAX_SRC; ++i) if (c->pull_peer_ipc[i] && c->pull_peer[i]) cudaIpcCloseMemHandle(c->pull_peer[i]);
    cudaFree(c->pull_base);
    if (c->h_ctr) cudaFreeHost(c->h_ctr);
    for (auto &e : c->timer) if (e) cudaEventDestroy(e);
    if (c->ev_signal) cudaEventDestroy(c->ev_signal);
    if (c->ev_wait) cudaEventDestroy(c->ev_wait);
    if (c->s_compute) cudaStreamDestroy(c->s_compute);
    if (c->s_copy) cudaStreamDestroy(c->s_copy);
}

}  // namespace

// pbk_seed_entries: the seeded k-mers get their seeded value, whatever the reads added (see override_records_kernel)
int apply_seeds(pbk_ctx *c)
{
    TRY(table_ready(c));
    const u64 n = c->seed_rec.size() / (c->W + 1);
    TRY(ensure_tables(c, std::max<u64>(n, 1024), false));
    if ((double)(c->occupied + n) > max_load(c) * (double)c->table.capacity())
        TRY(grow_table(c, &c->table, c->occupied, (u64)((c->occupied + n) / max_load(c)) + 1));
    TRY(ensure_overflow(c, n));
    u64 *d_rec = nullptr;
    const size_t bytes = c->seed_rec.size() * 8;
    TRY(dev_alloc(c, (void **)&d_rec, bytes));
    int rc = PBK_OK;
    if (cudaMemcpyAsync(d_rec, c->seed_rec.data(), bytes, cudaMemcpyHostToDevice, c->s_compute) != cudaSuccess) rc = fail(c, PBK_E_CUDA, "upload of seeded entries failed");
    c->h2d_bytes += bytes;
    if (rc == PBK_OK) {
        { Span sp(c, LC_OTHER); launch_override_records(d_rec, n, c->table, c->d_ctr, c->d_ovf, c->ovf_cap, c->sm_count, c->s_compute); }
        c->table_touched = true;
        if (cudaGetLastError() != cudaSuccess) rc = fail(c, PBK_E_CUDA, "override launch failed");
    }
    if (rc == PBK_OK) rc = read_counters(c);
    else cudaStreamSynchronize(c->s_compute);
    dev_free(c, d_rec, bytes);
    TRY(rc);
    // a seeded key that found no slot within the probe limit: absent from the table, so the weighted insert of the
    // overflow route sets it as well
    const ShardInfo keep = c->shard;
    c->shard = ShardInfo{1, 0};
    rc = drain_overflow(c);
    c->shard = keep;
    return rc;
}

// =================================================================================================
// exported functions
// =================================================================================================

extern "C" {

int pbk_abi_version(void) { return PBK_ABI_VERSION; }

const char *pbk_strerror(int s)
{
    switch (s) {
    case PBK_OK: return "ok";
    case PBK_E_ARG: return "bad argument";
    case PBK_E_NO_DEVICE: return "no usable CUDA device (libpbk has no CPU fallback)";
    case PBK_E_CUDA: return "CUDA error";
    case PBK_E_NOMEM: return "out of memory (HBM budget or host)";
    case PBK_E_READ_TOO_LONG: return "read length >= 500000 (platanus::ReadError)";
    case PBK_E_BAD_BASE: return "character without a Char2Bin code in the input";
    case PBK_E_KMER_DIST: return "empty k-mer distribution (platanus::KmerDistError)";
    case PBK_E_STATE: return "call out of order";
    case PBK_E_IO: return "file error (platanus::FILEError)";
    case PBK_E_UNSUPPORTED_K: return "k must be in 1..256";
    default: return "unknown status";
    }
}

const char *pbk_last_error(const pbk_ctx *ctx) { return ctx ? ctx->err.c_str() : ""; }

int pbk_create(pbk_ctx **out, const pbk_config *cfg_in)
{
    if (!out || !cfg_in || cfg_in->struct_size < PBK_CONFIG_SIZE_V1) return PBK_E_ARG;
    *out = nullptr;
    pbk_config cfg_full;
    memset(&cfg_full, 0, sizeof cfg_full);
    memcpy(&cfg_full, cfg_in, std::min<size_t>(cfg_in->struct_size, sizeof cfg_full));     // (a shorter, earlier layout: the new fields are 0)
    const pbk_config *cfg = &cfg_full;
    if (cfg->n_passes > 1 && (cfg->pass_index >= cfg->n_passes || cfg->n_passes > 0xFFFFu || cfg->n_shards > 1)) return PBK_E_ARG;
    if (cfg->k == 0 || cfg->k > PBK_MAX_K) return PBK_E_UNSUPPORTED_K;
    if (cfg->n_shards > 1 && cfg->shard_rank >= cfg->n_shards) return PBK_E_ARG;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) { cudaGetLastError(); return PBK_E_NO_DEVICE; }
    pbk_ctx *c = new (std::nothrow) pbk_ctx();
    if (!c) return PBK_E_NOMEM;
    int dev = cfg->device;
    if (dev < 0 && cudaGetDevice(&dev) != cudaSuccess) { delete c; return PBK_E_NO_DEVICE; }
    if (dev >= n_dev || cudaSetDevice(dev) != cudaSuccess) { delete c; return PBK_E_NO_DEVICE; }
    c->device = dev;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { delete c; return PBK_E_NO_DEVICE; }
    c->sm_count = prop.multiProcessorCount;
    c->k = cfg->k; c->W = (int)((cfg->k + 31) / 32); c->flags = cfg->flags;
    c->shard.n_shards = cfg->n_shards > 1 ? cfg->n_shards : 1;
    c->shard.rank = cfg->n_shards > 1 ? cfg->shard_rank : 0;
    c->table_hint = cfg->table_slots_hint;
    c->pass = cfg->n_passes > 1 ? ((cfg->n_passes << 16) | cfg->pass_index) : 0u;
    c->partition_enabled = !(cfg->flags & PBK_F_NO_PARTITION) && getenv("PBK_NO_PARTITION") == nullptr;
    c->partition_forced = (cfg->flags & PBK_F_FORCE_PARTITION) != 0;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    c->budget = cfg->hbm_budget_bytes ? cfg->hbm_budget_bytes : (u64)(free_b * 0.85);
    auto bail = [&](int code) { release_all(c); delete c; return code; };
    if (cudaStreamCreateWithFlags(&c->s_compute, cudaStreamNonBlocking) != cudaSuccess) return bail(PBK_E_CUDA);
    if (cudaStreamCreateWithFlags(&c->s_copy, cudaStreamNonBlocking) != cudaSuccess) return bail(PBK_E_CUDA);
    if (cudaStreamCreateWithFlags(&c->s_clear, cudaStreamNonBlocking) != cudaSuccess) return bail(PBK_E_CUDA);
    if (cudaStreamCreateWithFlags(&c->s_pack, cudaStreamNonBlocking) != cudaSuccess) return bail(PBK_E_CUDA);
    if (cudaEventCreateWithFlags(&c->ev_table_clear, cudaEventDisableTiming) != cudaSuccess) return bail(PBK_E_CUDA);
    if (cudaEventCreateWithFlags(&c->ev_pack_may_start, cudaEventDisableTiming) != cudaSuccess) return bail(PBK_E_CUDA);
    for (int i = 0; i < N_STAGE; ++i)
        if (cudaEventCreateWithFlags(&c->ev_pack[i], cudaEventDisableTiming) != cudaSuccess) return bail(PBK_E_CUDA);
    c->overlap_enabled = getenv("PBK_NO_OVERLAP") == nullptr;
    c->pipeline_enabled = !(cfg->flags & PBK_F_NO_PIPELINE) && getenv("PBK_NO_PIPELINE") == nullptr;
    c->passb2_enabled = getenv("PBK_PASSB2") ? atoi(getenv("PBK_PASSB2")) != 0 : PASSB2_DEFAULT;
    c->passb2_gather = getenv("PBK_PASSB2_GATHER") && atoi(getenv("PBK_PASSB2_GATHER")) != 0;
    c->passb2_fresh = !(getenv("PBK_PASSB2_FRESH") && atoi(getenv("PBK_PASSB2_FRESH")) == 0);
    if (getenv("PBK_UNKNOWN_AS_N") && atoi(getenv("PBK_UNKNOWN_AS_N")) != 0) c->flags |= PBK_F_UNKNOWN_AS_N;
    if (cudaMallocHost((void **)&c->h_ctr, sizeof(Counters)) != cudaSuccess) return bail(PBK_E_NOMEM);
    memset(c->h_ctr, 0, sizeof(Counters));
    if (dev_alloc(c, (void **)&c->d_ctr, sizeof(Counters)) || dev_alloc(c, (void **)&c->d_len_hist, PBK_LEN_BINS * 8) ||
        dev_alloc(c, (void **)&c->d_occ_hist, PBK_OCC_BINS * 8) || dev_alloc(c, (void **)&c->d_shard_counts, (c->shard.n_shards + 1) * 8))
        return bail(PBK_E_NOMEM);
    cudaMemsetAsync(c->d_ctr, 0, sizeof(Counters), c->s_compute);
    cudaMemsetAsync(c->d_len_hist, 0, PBK_LEN_BINS * 8, c->s_compute);
    if (cudaStreamSynchronize(c->s_compute) != cudaSuccess) return bail(PBK_E_CUDA);
    *out = c;
    return PBK_OK;
}

void pbk_destroy(pbk_ctx *c)
{
    if (!c) return;
    release_all(c);
    delete c;
}

int pbk_host_alloc(void **out, size_t bytes)
{
    if (!out) return PBK_E_ARG;
    cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
    if (e != cudaSuccess) { cudaGetLastError(); return e == cudaErrorMemoryAllocation ? PBK_E_NOMEM : PBK_E_NO_DEVICE; }
    return PBK_OK;
}

void pbk_host_free(void *p) { if (p) cudaFreeHost(p); }

int pbk_push_reads(pbk_ctx *c, const uint8_t *bases, const uint64_t *read_offsets, uint64_t n_reads,
                   int encoding, const int32_t *n_pos, const uint64_t *n_pos_offsets)
{
    if (!c) return PBK_E_ARG;
    if (n_reads == 0) return PBK_OK;
    if (!read_offsets || (encoding != PBK_ENC_ASCII && encoding != PBK_ENC_PLATANUS)) return fail(c, PBK_E_ARG, "bad arguments");
    if (read_offsets[0] != 0) return fail(c, PBK_E_ARG, "read_offsets[0] must be 0");
    const u64 n_bases = read_offsets[n_reads];
    if (n_bases && !bases) return fail(c, PBK_E_ARG, "bases is NULL");
    // A batch keeps its 2-bit stream and its bucket store (8W bytes per window) in HBM, so very large pushes are cut
    // into internal batches of at most MAX_PUSH_BASES bases at read boundaries.
    static const u64 max_push = getenv("PBK_MAX_PUSH_BASES") ? strtoull(getenv("PBK_MAX_PUSH_BASES"), nullptr, 10) : MAX_PUSH_BASES;
    if (n_bases <= max_push)
        return push_common(c, bases, nullptr, (const u64 *)read_offsets, nullptr, n_reads, n_bases, encoding, n_pos, (const u64 *)n_pos_offsets);
    std::vector<u64> off, npo;
    for (u64 r0 = 0; r0 < n_reads;) {
        u64 r1 = r0 + 1;                                               // at least one read per batch
        while (r1 < n_reads && read_offsets[r1 + 1] - read_offsets[r0] <= max_push) ++r1;
        off.resize(r1 - r0 + 1);
        for (u64 i = r0; i <= r1; ++i) off[i - r0] = read_offsets[i] - read_offsets[r0];
        const int32_t *np = nullptr;
        if (encoding == PBK_ENC_PLATANUS && n_pos && n_pos_offsets) {
            npo.resize(r1 - r0 + 1);
            for (u64 i = r0; i <= r1; ++i) npo[i - r0] = n_pos_offsets[i] - n_pos_offsets[r0];
            np = n_pos + n_pos_offsets[r0];
        }
        TRY(push_common(c, bases + read_offsets[r0], nullptr, off.data(), nullptr, r1 - r0, off.back(), encoding, np,
                        np ? npo.data() : nullptr));
        r0 = r1;
    }
    return PBK_OK;
}

int pbk_push_reads_packed(pbk_ctx *c, const uint64_t *words, const uint64_t *read_offsets, uint64_t n_reads,
                          const uint64_t *n_positions, uint64_t n_n)
{
    if (!c) return PBK_E_ARG;
    if (n_reads == 0) return PBK_OK;
    if (!read_offsets || read_offsets[0] != 0) return fail(c, PBK_E_ARG, "bad read_offsets");
    const u64 n_bases = read_offsets[n_reads];
    if (n_bases && !words) return fail(c, PBK_E_ARG, "words is NULL");
    if (n_n && !n_positions) return fail(c, PBK_E_ARG, "n_positions is NULL");
    if (n_bases > MAX_PUSH_BASES) return fail(c, PBK_E_ARG, "pbk_push_reads_packed takes at most %llu bases per call", (unsigned long long)MAX_PUSH_BASES);
    // (push_common's n_pos / n_pos_offsets slots carry the absolute N positions and their number for this encoding)
    const u64 count = n_n;
    return push_common(c, reinterpret_cast<const uint8_t *>(words), nullptr, (const u64 *)read_offsets, nullptr, n_reads, n_bases,
                       PBK_ENC_PACKED2, reinterpret_cast<const int32_t *>(n_positions), &count);
}

int pbk_push_reads_device(pbk_ctx *c, const void *d_bases, const void *d_read_offsets, uint64_t n_reads, uint64_t n_bases)
{
    if (!c) return PBK_E_ARG;
    if (n_reads == 0) return PBK_OK;
    if (!d_read_offsets || (n_bases && !d_bases)) return fail(c, PBK_E_ARG, "NULL device pointer");
    return push_common(c, nullptr, (const uint8_t *)d_bases, nullptr, (const u64 *)d_read_offsets, n_reads, n_bases,
                       PBK_ENC_ASCII, nullptr, nullptr);
}

int pbk_finalize(pbk_ctx *c, uint64_t *occ_hist, uint64_t *len_hist, uint64_t *n_distinct,
                 uint64_t *n_instances, uint64_t *max_occurrence)
{
    if (!c) return PBK_E_ARG;
    CK(cudaSetDevice(c->device));
    TRY(settle(c));
    if (c->shard.n_shards > 1 && c->occupied_remote > 0)
        return fail(c, PBK_E_STATE, "%llu staged records have not been exchanged (pbk_shard_pack_device)", (unsigned long long)c->occupied_remote);
    if (!c->seed_rec.empty()) TRY(apply_seeds(c));
    c->h_occ_hist.assign(PBK_OCC_BINS, 0);
    if (c->table.slots) {
        CK(cudaMemsetAsync(c->d_occ_hist, 0, PBK_OCC_BINS * 8, c->s_compute));
        { Span sp(c, LC_OTHER); launch_table_histogram(c->table, c->d_occ_hist, c->s_compute); }
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(c->h_occ_hist.data(), c->d_occ_hist, PBK_OCC_BINS * 8, cudaMemcpyDeviceToHost, c->s_compute));
        c->d2h_bytes += PBK_OCC_BINS * 8;
    }
    if (len_hist) {
        CK(cudaMemcpyAsync(len_hist, c->d_len_hist, PBK_LEN_BINS * 8, cudaMemcpyDeviceToHost, c->s_compute));
        c->d2h_bytes += PBK_LEN_BINS * 8;
    }
    CK(cudaStreamSynchronize(c->s_compute));
    c->inst_since_clamp = 0;
    u64 nd = 0, mx = 0;
    for (u64 i = 1; i < PBK_OCC_BINS; ++i) { nd += c->h_occ_hist[i]; if (c->h_occ_hist[i]) mx = i; }
    if (nd != c->occupied)
        return fail(c, PBK_E_CUDA, "internal: histogram holds %llu keys, table tracked %llu", (unsigned long long)nd, (unsigned long long)c->occupied);
    if (occ_hist) memcpy(occ_hist, c->h_occ_hist.data(), PBK_OCC_BINS * 8);
    if (n_distinct) *n_distinct = nd;
    if (n_instances) *n_instances = c->last.instances;
    if (max_occurrence) *max_occurrence = mx;
    c->finalized = true;
    return PBK_OK;
}

int pbk_export(pbk_ctx *c, uint32_t min_count, int sorted, uint64_t *keys, uint16_t *counts, uint64_t capacity, uint64_t *n_out)
{
    if (!c) return PBK_E_ARG;
    if (!c->finalized) return fail(c, PBK_E_STATE, "pbk_export before pbk_finalize");
    CK(cudaSetDevice(c->device));
    TRY(table_ready(c));
    u64 n = 0;
    for (u64 i = std::max<u64>(min_count, 1); i < PBK_OCC_BINS; ++i) n += c->h_occ_hist[i];
    if (n_out) *n_out = n;
    if (capacity == 0) return PBK_OK;
    if (capacity < n || !keys || !counts) return fail(c, PBK_E_ARG, "export needs room for %llu entries", (unsigned long long)n);
    if (n == 0) return PBK_OK;
    // entries, their counter and the sort's temporaries live in one grow-only block of the context (no cudaMalloc /
    // cudaFree -- both synchronise the device -- on the repeated-export path)
    const size_t kb = n * 8 * c->W, cb = n * 2;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t sort_bytes = sorted ? sort_export_scratch_bytes(n, c->W) : 0;
    const size_t need = al(kb) + al(cb) + 256 + sort_bytes;
    if (need > c->scratch_bytes) {
        CK(cudaStreamSynchronize(c->s_compute));
        dev_free(c, c->d_scratch, c->scratch_bytes);
        c->d_scratch = nullptr; c->scratch_bytes = 0;
        TRY(dev_alloc(c, &c->d_scratch, need + need / 8));
        c->scratch_bytes = need + need / 8;
    }
    char *base = (char *)c->d_scratch;
    u64 *d_keys = (u64 *)base; uint16_t *d_counts = (uint16_t *)(base + al(kb)); u64 *d_n = (u64 *)(base + al(kb) + al(cb));
    void *d_sort = base + al(kb) + al(cb) + 256;
    int rc = PBK_OK;
    cudaMemsetAsync(d_n, 0, 8, c->s_compute);
    { Span sp(c, LC_OTHER); launch_table_export(c->table, std::max<u32>(min_count, 1), d_keys, d_counts, n, d_n, c->s_compute); }
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && sorted) { c->launches[LC_OTHER] += 1; e = sort_export(d_keys, d_counts, n, c->W, (int)c->k, d_sort, sort_bytes, c->s_compute); }
    u64 got = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&got, d_n, 8, cudaMemcpyDeviceToHost, c->s_compute);
    if (e == cudaSuccess) e = cudaMemcpyAsync(keys, d_keys, kb, cudaMemcpyDeviceToHost, c->s_compute);
    if (e == cudaSuccess) e = cudaMemcpyAsync(counts, d_counts, cb, cudaMemcpyDeviceToHost, c->s_compute);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->s_compute);
    c->d2h_bytes += kb + cb + 8;
    if (e != cudaSuccess) { cudaStreamSynchronize(c->s_compute); rc = fail(c, PBK_E_CUDA, "export: %s", cudaGetErrorString(e)); }
    else if (got != n) rc = fail(c, PBK_E_CUDA, "internal: export found %llu entries, histogram says %llu", (unsigned long long)got, (unsigned long long)n);
    return rc;
}

int pbk_get_stats(const pbk_ctx *cc, pbk_stats *out)
{
    if (!cc || !out) return PBK_E_ARG;
    pbk_ctx *c = const_cast<pbk_ctx *>(cc);
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->s_compute);
    if (settle(c) != PBK_OK) return PBK_E_CUDA;
    resolve_spans(c);
    memset(out, 0, sizeof(*out));
    out->n_reads = c->n_reads; out->n_bases = c->n_bases; out->n_instances = c->last.instances;
    out->n_distinct = c->occupied; out->table_slots = c->table.slots ? c->table.capacity() : 0;
    out->table_bytes = c->table.slots ? c->table.bytes() : 0; out->n_grow = c->n_grow;
    out->launches_pack = c->launches[LC_PACK]; out->launches_other = c->launches[LC_OTHER];
    out->launches_partition = c->launches[LC_PART]; out->launches_insert = c->launches[LC_INSERT];
    out->launches_count = c->launches[LC_COUNT] + c->launches[LC_PART] + c->launches[LC_INSERT];
    out->ms_pack = c->ms[LC_PACK]; out->ms_other = c->ms[LC_OTHER];
    out->ms_partition = c->ms[LC_PART]; out->ms_insert = c->ms[LC_INSERT];
    out->ms_count = c->ms[LC_COUNT] + c->ms[LC_PART] + c->ms[LC_INSERT];
    out->h2d_bytes = c->h2d_bytes; out->d2h_bytes = c->d2h_bytes;
    out->ms_count_elapsed = out->ms_count;         // all counting launches are serial on one stream
    out->n_pipelined_batches = c->n_pipelined;
    out->n_split_build = c->n_passb2;
    return PBK_OK;
}

int pbk_set_timing(pbk_ctx *c, int on)
{
    if (!c) return PBK_E_ARG;
    if (on) c->flags |= PBK_F_TIMING; else c->flags &= ~(u32)PBK_F_TIMING;
    return PBK_OK;
}

int pbk_timer_mark(pbk_ctx *c, int slot)
{
    if (!c || slot < 0 || slot >= 8) return PBK_E_ARG;
    CK(cudaSetDevice(c->device));
    if (!c->timer[slot]) CK(cudaEventCreate(&c->timer[slot]));
    CK(cudaEventRecord(c->timer[slot], c->s_compute));
    return PBK_OK;
}

int pbk_timer_elapsed_ms(pbk_ctx *c, int start, int stop, double *ms)
{
    if (!c || !ms || start < 0 || start >= 8 || stop < 0 || stop >= 8 || !c->timer[start] || !c->timer[stop]) return PBK_E_ARG;
    float f = 0;
    CK(cudaEventSynchronize(c->timer[stop]));
    CK(cudaEventElapsedTime(&f, c->timer[start], c->timer[stop]));
    *ms = f;
    return PBK_OK;
}

int pbk_reset(pbk_ctx *c, uint32_t k)
{
    if (!c) return PBK_E_ARG;
    if (k == 0) k = c->k;
    if (k > PBK_MAX_K) return PBK_E_UNSUPPORTED_K;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->s_compute));
    c->counters_pending = false;                       // whatever was queued is about to be forgotten
    c->pending_new = 0;
    const int W = (int)((k + 31) / 32);
    if (c->table_clear_pending) { CK(cudaStreamSynchronize(c->s_clear)); c->table_clear_pending = false; }
    if (W != c->W) {                                   // slot size changes: tables are re-created lazily
        if (c->table.slots) dev_free(c, c->table.slots, c->table.bytes());
        if (c->remote.slots) dev_free(c, c->remote.slots, c->remote.bytes());
        if (c->d_ovf) dev_free(c, c->d_ovf, c->ovf_cap * (c->W + 1) * 8);
        c->table = TableView{nullptr, 0, 0}; c->remote = TableView{nullptr, 0, 0}; c->d_ovf = nullptr; c->ovf_cap = 0;
    } else if (c->overlap_enabled && !(c->flags & PBK_F_TIMING)) {
        // zero-fill on the side stream (s_compute is idle: synchronised above); the first table access waits for it (table_ready)
        if (c->table.slots) { c->launches[LC_OTHER] += 1; launch_table_init(c->table, c->s_clear); }
        if (c->remote.slots) { c->launches[LC_OTHER] += 1; launch_table_init(c->remote, c->s_clear); }
        CK(cudaEventRecord(c->ev_table_clear, c->s_clear));
        c->table_clear_pending = true;
    } else {
        TRY(table_ready(c));
        if (c->table.slots) { Span sp(c, LC_OTHER); launch_table_init(c->table, c->s_compute); }
        if (c->remote.slots) { Span sp(c, LC_OTHER); launch_table_init(c->remote, c->s_compute); }
    }
    c->table_touched = false;                          // (zero-filled, or gone and zero-filled again when it is re-created)
    c->remote_dirty = false; c->stage_gen += 1;
    c->seed_rec.clear();
    c->k = k; c->W = W;
    CK(cudaMemsetAsync(c->d_ctr, 0, sizeof(Counters), c->s_compute));
    CK(cudaMemsetAsync(c->d_len_hist, 0, PBK_LEN_BINS * 8, c->s_compute));
    // (no synchronize here: everything that follows is ordered on s_compute, and the next push's H2D copies on the
    //  copy stream only touch the staging buffers, so they may run while the table is still being cleared)
    c->last = Counters{}; c->occupied = c->occupied_remote = 0; c->inst_since_clamp = 0;
    c->n_reads = c->n_bases = 0; c->finalized = false; c->h_occ_hist.clear();
    return PBK_OK;
}

// ---- sharding ---------------------------------------------------------------------------------

uint32_t pbk_shard_record_bytes(const pbk_ctx *c) { return c ? (uint32_t)((c->W + 1) * 8) : 0; }

int pbk_shard_send_counts(pbk_ctx *c, uint64_t *counts)
{
    if (!c || !counts) return PBK_E_ARG;
    TRY(settle(c));
    const u32 n = c->shard.n_shards;
    if (c->shard_counts_gen == c->stage_gen && c->h_shard_counts.size() == n) {    // nothing staged since the last scan
        memcpy(counts, c->h_shard_counts.data(), n * 8);
        return PBK_OK;
    }
    c->h_shard_counts.assign(n, 0);
    c->shard_counts_gen = c->stage_gen;
    if (n > 1 && c->remote.slots && c->occupied_remote) {
        CK(cudaSetDevice(c->device));
        CK(cudaMemsetAsync(c->d_shard_counts, 0, n * 8, c->s_compute));
        { Span sp(c, LC_OTHER); launch_shard_count(c->remote, n, c->d_shard_counts, c->s_compute); }
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(c->h_shard_counts.data(), c->d_shard_counts, n * 8, cudaMemcpyDeviceToHost, c->s_compute));
        CK(cudaStreamSynchronize(c->s_compute));
        c->d2h_bytes += n * 8;
    }
    memcpy(counts, c->h_shard_counts.data(), n * 8);
    return PBK_OK;
}

int pbk_shard_pack_device(pbk_ctx *c, void *d_records, uint64_t capacity_records)
{
    if (!c) return PBK_E_ARG;
    const u32 n = c->shard.n_shards;
    if (n <= 1 || !c->remote.slots || c->occupied_remote == 0) return PBK_OK;
    std::vector<uint64_t> cnt(n);
    TRY(pbk_shard_send_counts(c, cnt.data()));
    std::vector<u64> cur(n);
    u64 total = 0;
    for (u32 i = 0; i < n; ++i) { cur[i] = total; total += cnt[i]; }
    if (total != c->occupied_remote) return fail(c, PBK_E_CUDA, "internal: staged %llu records, tracked %llu", (unsigned long long)total, (unsigned long long)c->occupied_remote);
    if (total > capacity_records || !d_records) return fail(c, PBK_E_ARG, "exchange buffer too small: need %llu records", (unsigned long long)total);
    CK(cudaMemcpyAsync(c->d_shard_counts, cur.data(), n * 8, cudaMemcpyHostToDevice, c->s_compute));
    { Span sp(c, LC_OTHER); launch_shard_pack(c->remote, n, c->d_shard_counts, (u64 *)d_records, c->s_compute); }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->s_compute));
    c->occupied_remote = 0;
    c->remote_dirty = true;                         // cleared before the next push touches it
    c->stage_gen += 1;
    return PBK_OK;
}

static int insert_records_device(pbk_ctx *c, const void *d_records, uint64_t n_records);

int pbk_shard_insert_device(pbk_ctx *c, const void *d_records, uint64_t n_records)
{
    if (!c) return PBK_E_ARG;
    if (n_records == 0) return PBK_OK;
    if (!d_records) return PBK_E_ARG;
    if (c->finalized) return fail(c, PBK_E_STATE, "insert after finalize");
    return insert_records_device(c, d_records, n_records);
}

// weighted inserts of (key words, count) records that this context owns
static int insert_records_device(pbk_ctx *c, const void *d_records, uint64_t n_records)
{
    CK(cudaSetDevice(c->device));
    TRY(settle(c));
    TRY(ensure_tables(c, n_records, false));
    const ShardInfo local{1, 0};                      // received records are owned by this shard
    for (u64 at = 0; at < n_records; at += CHUNK_BASES) {
        const u64 n = std::min<u64>(CHUNK_BASES, n_records - at);
        TRY(maybe_clamp(c, (u64)c->shard.n_shards * COUNT_SAT));
        if ((double)(c->occupied + n) > max_load(c) * (double)c->table.capacity())
            TRY(grow_table(c, &c->table, c->occupied, (u64)((c->occupied + n) / max_load(c)) + 1));
        TRY(ensure_overflow(c, n));
        {
            Span sp(c, LC_COUNT);
            launch_insert_records((const u64 *)d_records + at * (c->W + 1), n, true, c->table, c->remote, local, c->d_ctr,
                                  c->d_ovf, c->ovf_cap, c->sm_count, c->s_compute);
            c->table_touched = true;
        }
        CK(cudaGetLastError());
        TRY(read_counters(c));
        TRY(drain_overflow(c));
    }
    return PBK_OK;
}

// ---- consumers of the table: bulk load and occurrence lookup ---------------------------------------

int pbk_load_entries(pbk_ctx *c, const uint64_t *keys, const uint16_t *counts, uint64_t n)
{
    if (!c) return PBK_E_ARG;
    if (n == 0) return PBK_OK;
    if (!keys || !counts) return fail(c, PBK_E_ARG, "NULL entries");
    if (c->finalized) return fail(c, PBK_E_STATE, "pbk_load_entries after pbk_finalize (call pbk_reset first)");
    CK(cudaSetDevice(c->device));
    const int W = c->W;
    const u64 PIECE = 1ull << 22;                                    // records per upload
    std::vector<u64> rec;
    u64 *d_rec = nullptr;
    const size_t bytes = (size_t)std::min<u64>(n, PIECE) * (W + 1) * 8;
    TRY(dev_alloc(c, (void **)&d_rec, bytes));
    int rc = PBK_OK;
    for (u64 at = 0; at < n && rc == PBK_OK; at += PIECE) {
        const u64 m = std::min<u64>(PIECE, n - at);
        rec.clear();
        rec.reserve((size_t)m * (W + 1));
        for (u64 i = at; i < at + m; ++i) {
            if (counts[i] == 0) continue;                            // an empty slot of the reference table
            if (c->shard.n_shards > 1 && pbk_shard_of_key(keys + i * W, c->k, c->shard.n_shards) != c->shard.rank) continue;
            if (c->pass && pbk_shard_of_key(keys + i * W, c->k, c->pass >> 16) != (c->pass & 0xFFFFu)) continue;   // another pass's key
            for (int j = 0; j < W; ++j) rec.push_back(keys[i * W + j]);
            rec.push_back(counts[i]);
        }
        const u64 kept = rec.size() / (W + 1);
        if (kept == 0) continue;
        if (cudaMemcpyAsync(d_rec, rec.data(), rec.size() * 8, cudaMemcpyHostToDevice, c->s_compute) != cudaSuccess ||
            cudaStreamSynchronize(c->s_compute) != cudaSuccess) { rc = fail(c, PBK_E_CUDA, "upload of table entries failed"); break; }
        c->h2d_bytes += rec.size() * 8;
        rc = insert_records_device(c, d_rec, kept);
    }
    cudaStreamSynchronize(c->s_compute);
    dev_free(c, d_rec, bytes);
    return rc;
}

// h_out / d_out: occurrence per base (pbk_lookup); h_match: one byte per read instead (pbk_match_reads)
static int lookup_common(pbk_ctx *c, const uint8_t *h_bases, const uint8_t *d_bases_in, const u64 *h_offsets,
                         const u64 *d_offsets_in, u64 n_reads, u64 n_bases, int encoding, const int32_t *n_pos,
                         const u64 *n_pos_offsets, uint16_t *h_out, uint16_t *d_out, uint8_t *h_match = nullptr)
{
    CK(cudaSetDevice(c->device));
    TRY(settle(c));
    if (n_bases > MAX_PUSH_BASES) return fail(c, PBK_E_ARG, "pbk_lookup takes at most %llu bases per call", (unsigned long long)MAX_PUSH_BASES);
    if (h_match) memset(h_match, 0, n_reads);
    else if (h_out) memset(h_out, 0, n_bases * 2);
    else CK(cudaMemsetAsync(d_out, 0, n_bases * 2, c->s_compute));
    if (!c->table.slots || n_bases < c->k) { CK(cudaStreamSynchronize(c->s_compute)); return PBK_OK; }
    TRY(ensure_batch_buffers(c, n_bases, n_reads));
    if (!c->d_len_scratch) {
        TRY(dev_alloc(c, (void **)&c->d_len_scratch, PBK_LEN_BINS * 8));
        TRY(dev_alloc(c, (void **)&c->d_ctr_scratch, sizeof(Counters)));
    }
    CK(cudaMemsetAsync(c->d_ctr_scratch, 0, sizeof(Counters), c->s_compute));
    u64 *stream = c->d_stream_raw + STREAM_PAD_WORDS;
    u32 *nflag = c->d_nflag_raw + STREAM_PAD_WORDS, *rflag = c->d_rflag_raw + STREAM_PAD_WORDS;
    const u64 words = (n_bases + 31) / 32;
    const u64 *d_off = d_offsets_in;
    if (!d_off) {
        CK(cudaMemcpyAsync(c->d_offsets, h_offsets, (n_reads + 1) * 8, cudaMemcpyHostToDevice, c->s_compute));
        c->h2d_bytes += (n_reads + 1) * 8;
        d_off = c->d_offsets;
    }
    CK(cudaMemsetAsync(rflag, 0, words * 4, c->s_compute));
    { Span sp(c, LC_OTHER); launch_read_marks(d_off, n_reads, c->d_len_scratch, rflag, c->d_ctr_scratch, c->s_compute); }
    CK(cudaGetLastError());
    uint8_t *d_tmp_stage = nullptr;
    if (!d_bases_in) TRY(dev_alloc(c, (void **)&d_tmp_stage, CHUNK_BASES));
    uint16_t *d_occ = nullptr;
    uint8_t *d_match = nullptr;
    int32_t *d_np = nullptr; u64 *d_npo = nullptr;
    const u64 total_n = (encoding == PBK_ENC_PLATANUS && n_pos_offsets) ? n_pos_offsets[n_reads] : 0;
    int rc = dev_alloc(c, (void **)&d_occ, words * 64);
    for (u64 b0 = 0; b0 < n_bases && rc == PBK_OK; b0 += CHUNK_BASES) {
        const u64 nb = std::min(CHUNK_BASES, n_bases - b0), w0 = b0 / 32, nw = (nb + 31) / 32;
        const uint8_t *src = d_bases_in ? d_bases_in + b0 : d_tmp_stage;
        if (!d_bases_in) {
            if (cudaMemcpyAsync(d_tmp_stage, h_bases + b0, nb, cudaMemcpyHostToDevice, c->s_compute) != cudaSuccess) rc = fail(c, PBK_E_CUDA, "H2D copy failed");
            c->h2d_bytes += nb;
        }
        { Span sp(c, LC_PACK); launch_pack(src, nb, nw, encoding | ((c->flags & PBK_F_UNKNOWN_AS_N) ? 0x100 : 0), stream, nflag, w0, c->d_ctr_scratch, c->s_compute); }
    }
    if (rc == PBK_OK && encoding == PBK_ENC_PLATANUS) {
        if (!n_pos || !n_pos_offsets) rc = fail(c, PBK_E_ARG, "PBK_ENC_PLATANUS needs n_pos and n_pos_offsets");
        if (rc == PBK_OK) rc = dev_alloc(c, (void **)&d_np, total_n * 4);
        if (rc == PBK_OK) rc = dev_alloc(c, (void **)&d_npo, (n_reads + 1) * 8);
        if (rc == PBK_OK) {
            cudaMemcpyAsync(d_np, n_pos, total_n * 4, cudaMemcpyHostToDevice, c->s_compute);
            cudaMemcpyAsync(d_npo, n_pos_offsets, (n_reads + 1) * 8, cudaMemcpyHostToDevice, c->s_compute);
            c->h2d_bytes += total_n * 4 + (n_reads + 1) * 8;
            Span sp(c, LC_OTHER);
            launch_npos_scatter(d_off, d_np, d_npo, n_reads, nflag, c->s_compute);
        }
    }
    if (rc == PBK_OK) {
        { Span sp(c, LC_OTHER); launch_lookup(stream, nflag, rflag, 0, words, (int)c->k, c->table, d_occ, c->sm_count, c->s_compute); }
        // the kernel indexes by the window's END; the caller gets the window's START: shift by k - 1
        const u64 n_starts = n_bases - (c->k - 1);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess && h_match) {
            rc = dev_alloc(c, (void **)&d_match, n_reads);
            if (rc == PBK_OK) {
                { Span sp(c, LC_OTHER); launch_read_match(d_off, n_reads, d_occ, (int)c->k, d_match, c->sm_count, c->s_compute); }
                e = cudaGetLastError();
                if (e == cudaSuccess) e = cudaMemcpyAsync(h_match, d_match, n_reads, cudaMemcpyDeviceToHost, c->s_compute);
                c->d2h_bytes += n_reads;
            }
        } else if (e == cudaSuccess)
            e = h_out ? cudaMemcpyAsync(h_out, d_occ + (c->k - 1), n_starts * 2, cudaMemcpyDeviceToHost, c->s_compute)
                      : cudaMemcpyAsync(d_out, d_occ + (c->k - 1), n_starts * 2, cudaMemcpyDeviceToDevice, c->s_compute);
        if (e == cudaSuccess) e = cudaMemcpyAsync(c->h_ctr, c->d_ctr_scratch, sizeof(Counters), cudaMemcpyDeviceToHost, c->s_compute);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->s_compute);
        if (h_out && !h_match) c->d2h_bytes += n_starts * 2;
        if (rc != PBK_OK) {}
        else if (e != cudaSuccess) rc = fail(c, PBK_E_CUDA, "lookup: %s", cudaGetErrorString(e));
        else {
            const u32 flags = c->h_ctr->error_flags;
            *c->h_ctr = c->last;                                      // h_ctr mirrors the counting counters between calls
            // (no length limit here: ContigDivider::getOccurrenceArray walks contigs of any length, kmer_divide.cpp:151-197; the
            //  500000-base limit belongs to SEQ::convertFromString on the counting path)
            if (flags & ERR_BAD_BASE) rc = fail(c, PBK_E_BAD_BASE, "input contains a character with no Char2Bin code (only ACGTN, any case)");
        }
    }
    cudaStreamSynchronize(c->s_compute);
    dev_free(c, d_occ, words * 64); dev_free(c, d_tmp_stage, CHUNK_BASES); dev_free(c, d_match, n_reads);
    dev_free(c, d_np, total_n * 4); dev_free(c, d_npo, (n_reads + 1) * 8);
    return rc;
}

int pbk_lookup(pbk_ctx *c, const uint8_t *bases, const uint64_t *read_offsets, uint64_t n_reads, int encoding,
               const int32_t *n_pos, const uint64_t *n_pos_offsets, uint16_t *occ_out)
{
    if (!c) return PBK_E_ARG;
    if (n_reads == 0) return PBK_OK;
    if (!read_offsets || !occ_out || (encoding != PBK_ENC_ASCII && encoding != PBK_ENC_PLATANUS)) return fail(c, PBK_E_ARG, "bad arguments");
    if (read_offsets[0] != 0) return fail(c, PBK_E_ARG, "read_offsets[0] must be 0");
    const u64 n_bases = read_offsets[n_reads];
    if (n_bases && !bases) return fail(c, PBK_E_ARG, "bases is NULL");
    if (n_bases == 0) return PBK_OK;
    return lookup_common(c, bases, nullptr, (const u64 *)read_offsets, nullptr, n_reads, n_bases, encoding, n_pos,
                         (const u64 *)n_pos_offsets, occ_out, nullptr);
}

int pbk_match_reads(pbk_ctx *c, const uint8_t *bases, const uint64_t *read_offsets, uint64_t n_reads, int encoding,
                    const int32_t *n_pos, const uint64_t *n_pos_offsets, uint8_t *matched_out)
{
    if (!c) return PBK_E_ARG;
    if (n_reads == 0) return PBK_OK;
    if (!read_offsets || !matched_out || (encoding != PBK_ENC_ASCII && encoding != PBK_ENC_PLATANUS)) return fail(c, PBK_E_ARG, "bad arguments");
    if (read_offsets[0] != 0) return fail(c, PBK_E_ARG, "read_offsets[0] must be 0");
    const u64 n_bases = read_offsets[n_reads];
    if (n_bases && !bases) return fail(c, PBK_E_ARG, "bases is NULL");
    if (n_bases == 0) { memset(matched_out, 0, n_reads); return PBK_OK; }
    return lookup_common(c, bases, nullptr, (const u64 *)read_offsets, nullptr, n_reads, n_bases, encoding, n_pos,
                         (const u64 *)n_pos_offsets, nullptr, nullptr, matched_out);
}

int pbk_push_contigs(pbk_ctx *c, const uint8_t *bases, const uint64_t *seq_offsets, uint64_t n_seqs, const uint16_t *coverage,
                     uint64_t min_occurrence)
{
    if (!c) return PBK_E_ARG;
    if (n_seqs == 0) return PBK_OK;
    if (!seq_offsets || !coverage || seq_offsets[0] != 0) return fail(c, PBK_E_ARG, "bad arguments");
    if (c->finalized) return fail(c, PBK_E_STATE, "pbk_push_contigs after pbk_finalize (call pbk_reset first)");
    if (c->pass) return fail(c, PBK_E_STATE, "pbk_push_contigs is not available in a hash-range pass (pbk_config.n_passes)");
    const u64 n_bases = seq_offsets[n_seqs];
    if (n_bases == 0) return PBK_OK;
    if (!bases) return fail(c, PBK_E_ARG, "bases is NULL");
    if (n_bases > MAX_PUSH_BASES) return fail(c, PBK_E_ARG, "pbk_push_contigs takes at most %llu bases per call", (unsigned long long)MAX_PUSH_BASES);
    CK(cudaSetDevice(c->device));
    TRY(settle(c));
    // every window may be a new key: make room for all of them, so that no probe sequence can run out (see contig_max_kernel)
    const u64 windows = n_bases;
    TRY(ensure_tables(c, (u64)(windows / std::max(0.05, c->new_ratio)) + 1024, false));
    TRY(maybe_clamp(c, 0));
    if ((double)(c->occupied + windows) > max_load(c) * (double)c->table.capacity())
        TRY(grow_table(c, &c->table, c->occupied, (u64)((c->occupied + windows) / max_load(c)) + 1));
    TRY(ensure_batch_buffers(c, n_bases, n_seqs));
    if (!c->d_len_scratch) {
        TRY(dev_alloc(c, (void **)&c->d_len_scratch, PBK_LEN_BINS * 8));
        TRY(dev_alloc(c, (void **)&c->d_ctr_scratch, sizeof(Counters)));
    }
    CK(cudaMemsetAsync(c->d_ctr_scratch, 0, sizeof(Counters), c->s_compute));
    u64 *stream = c->d_stream_raw + STREAM_PAD_WORDS;
    u32 *nflag = c->d_nflag_raw + STREAM_PAD_WORDS, *rflag = c->d_rflag_raw + STREAM_PAD_WORDS;
    const u64 words = (n_bases + 31) / 32;
    // value of every stream position: max(coverage of its contig, minOccurrence) (counter.h:573-574)
    std::vector<uint16_t> val(words * 32, 0);
    for (u64 r = 0; r < n_seqs; ++r) {
        const u64 v = std::min<u64>(std::max<u64>(coverage[r], min_occurrence), COUNT_SAT);
        std::fill(val.begin() + seq_offsets[r], val.begin() + seq_offsets[r + 1], (uint16_t)v);
    }
    CK(cudaMemcpyAsync(c->d_offsets, seq_offsets, (n_seqs + 1) * 8, cudaMemcpyHostToDevice, c->s_compute));
    CK(cudaMemsetAsync(rflag, 0, words * 4, c->s_compute));
    { Span sp(c, LC_OTHER); launch_read_marks(c->d_offsets, n_seqs, c->d_len_scratch, rflag, c->d_ctr_scratch, c->s_compute); }
    uint8_t *d_stage = nullptr; uint16_t *d_val = nullptr;
    TRY(dev_alloc(c, (void **)&d_stage, CHUNK_BASES));
    int rc = dev_alloc(c, (void **)&d_val, words * 64);
    if (rc == PBK_OK && cudaMemcpyAsync(d_val, val.data(), words * 64, cudaMemcpyHostToDevice, c->s_compute) != cudaSuccess) rc = fail(c, PBK_E_CUDA, "H2D copy failed");
    c->h2d_bytes += words * 64 + n_bases + (n_seqs + 1) * 8;
    for (u64 b0 = 0; b0 < n_bases && rc == PBK_OK; b0 += CHUNK_BASES) {
        const u64 nb = std::min(CHUNK_BASES, n_bases - b0), w0 = b0 / 32, nw = (nb + 31) / 32;
        if (cudaMemcpyAsync(d_stage, bases + b0, nb, cudaMemcpyHostToDevice, c->s_compute) != cudaSuccess) rc = fail(c, PBK_E_CUDA, "H2D copy failed");
        { Span sp(c, LC_PACK); launch_pack(d_stage, nb, nw, PBK_ENC_ASCII, stream, nflag, w0, c->d_ctr_scratch, c->s_compute); }
    }
    if (rc == PBK_OK) {
        { Span sp(c, LC_OTHER); launch_contig_max(stream, nflag, rflag, 0, words, (int)c->k, c->table, d_val, c->d_ctr, c->sm_count, c->s_compute); }
        c->table_touched = true;
        if (cudaGetLastError() != cudaSuccess) rc = fail(c, PBK_E_CUDA, "contig_max launch failed");
    }
    if (rc == PBK_OK) rc = read_counters(c);
    else cudaStreamSynchronize(c->s_compute);
    if (rc == PBK_OK && c->h_ctr->overflow_n > 0) {
        cudaMemsetAsync(&c->d_ctr->overflow_n, 0, sizeof(u64), c->s_compute);
        c->last.overflow_n = 0; c->h_ctr->overflow_n = 0;
        rc = fail(c, PBK_E_CUDA, "internal: a contig k-mer found no slot in a table sized for every window");
    }
    if (rc == PBK_OK) {                                   // characters without a Char2Bin code, over-long sequences
        Counters sc;
        if (cudaMemcpyAsync(&sc, c->d_ctr_scratch, sizeof sc, cudaMemcpyDeviceToHost, c->s_compute) != cudaSuccess ||
            cudaStreamSynchronize(c->s_compute) != cudaSuccess) rc = fail(c, PBK_E_CUDA, "read-back failed");
        else if (sc.error_flags & ERR_BAD_BASE) rc = fail(c, PBK_E_BAD_BASE, "input contains a character with no Char2Bin code (only ACGTN, any case)");
    }
    cudaStreamSynchronize(c->s_compute);
    dev_free(c, d_stage, CHUNK_BASES); dev_free(c, d_val, words * 64);
    return rc;
}

int pbk_seed_entries(pbk_ctx *c, const uint64_t *keys, const uint16_t *counts, uint64_t n)
{
    if (!c) return PBK_E_ARG;
    if (n == 0) return PBK_OK;
    if (!keys || !counts) return fail(c, PBK_E_ARG, "NULL entries");
    if (c->finalized) return fail(c, PBK_E_STATE, "pbk_seed_entries after pbk_finalize (call pbk_reset first)");
    const int W = c->W;
    for (u64 i = 0; i < n; ++i) {
        if (counts[i] == 0) continue;                                // counter.h:700: only entries with a value are written
        if (c->shard.n_shards > 1 && pbk_shard_of_key(keys + i * W, c->k, c->shard.n_shards) != c->shard.rank) continue;
        if (c->pass && pbk_shard_of_key(keys + i * W, c->k, c->pass >> 16) != (c->pass & 0xFFFFu)) continue;       // another pass's key
        for (int j = 0; j < W; ++j) c->seed_rec.push_back(keys[i * W + j]);
        c->seed_rec.push_back(counts[i]);
    }
    return PBK_OK;
}

int pbk_lookup_device(pbk_ctx *c, const void *d_bases, const void *d_read_offsets, uint64_t n_reads, uint64_t n_bases, void *d_occ_out)
{
    if (!c) return PBK_E_ARG;
    if (n_reads == 0 || n_bases == 0) return PBK_OK;
    if (!d_bases || !d_read_offsets || !d_occ_out) return fail(c, PBK_E_ARG, "NULL device pointer");
    return lookup_common(c, nullptr, (const uint8_t *)d_bases, nullptr, (const u64 *)d_read_offsets, n_reads, n_bases,
                         PBK_ENC_ASCII, nullptr, nullptr, nullptr, (uint16_t *)d_occ_out);
}

// the graph builder's eight findValue probes per kept k-mer, for all of them at once
int pbk_neighbor_flags(pbk_ctx *c, uint32_t min_count, const uint64_t *keys, uint64_t n, uint8_t *flags_out)
{
    if (!c) return PBK_E_ARG;
    if (n == 0) return PBK_OK;
    if (!keys || !flags_out) return fail(c, PBK_E_ARG, "NULL argument");
    CK(cudaSetDevice(c->device));
    TRY(settle(c));
    memset(flags_out, 0, n);
    if (!c->table.slots) return PBK_OK;
    const size_t kb = (size_t)n * c->W * 8;
    u64 *d_keys = nullptr; uint8_t *d_out = nullptr;
    TRY(dev_alloc(c, (void **)&d_keys, kb));
    int rc = dev_alloc(c, (void **)&d_out, n);
    if (rc == PBK_OK) {
        cudaError_t e = cudaMemcpyAsync(d_keys, keys, kb, cudaMemcpyHostToDevice, c->s_compute);
        c->h2d_bytes += kb;
        if (e == cudaSuccess) {
            { Span sp(c, LC_OTHER); launch_neighbor_flags(d_keys, n, (int)c->k, c->table, min_count, d_out, c->sm_count, c->s_compute); }
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(flags_out, d_out, n, cudaMemcpyDeviceToHost, c->s_compute);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->s_compute);
        c->d2h_bytes += n;
        if (e != cudaSuccess) rc = fail(c, PBK_E_CUDA, "neighbor flags: %s", cudaGetErrorString(e));
    }
    cudaStreamSynchronize(c->s_compute);
    dev_free(c, d_keys, kb); dev_free(c, d_out, n);
    return rc;
}

// ---- key exchange (k <= 32) ----------------------------------------------------------------------

int pbk_keyx_plan(pbk_ctx *c, uint64_t max_windows_any_rank, pbk_keyx_layout *out)
{
    if (!c || !out) return PBK_E_ARG;
    if (c->W != 1) return fail(c, PBK_E_UNSUPPORTED_K, "the key exchange is implemented for k <= 32 (k = %u: use the record exchange)", c->k);
    if (c->shard.n_shards < 2 || c->shard.n_shards > 64) return fail(c, PBK_E_ARG, "key exchange needs 2..64 shards");
    c->keyx_plan = plan_partition_keyx(c->shard.n_shards, std::max<u64>(max_windows_any_rank, 1), c->W);
    c->keyx_max_windows = max_windows_any_rank;
    out->n_dest = c->shard.n_shards;
    out->n_regions = c->keyx_plan.n_buckets / c->shard.n_shards;
    out->seg_cap = c->keyx_plan.seg_cap;
    out->entry_bytes = 8ull * c->W;
    out->bytes_per_dest = (u64)out->n_regions * out->seg_cap * out->entry_bytes;
    out->cursors_per_dest = out->n_regions;
    return PBK_OK;
}

static int keyx_bind(pbk_ctx *c, void *d_send, void *d_cursors)
{
    if (!d_send || !d_cursors) return fail(c, PBK_E_ARG, "NULL key-exchange buffer");
    if (c->keyx_plan.n_buckets == 0) return fail(c, PBK_E_STATE, "pbk_keyx_partition before pbk_keyx_plan");
    if (c->W != 1) return fail(c, PBK_E_UNSUPPORTED_K, "the key exchange is implemented for k <= 32");
    c->keyx_send = (u64 *)d_send; c->keyx_cursors = (u64 *)d_cursors;
    return PBK_OK;
}

int pbk_keyx_partition(pbk_ctx *c, const uint8_t *bases, const uint64_t *read_offsets, uint64_t n_reads, int encoding,
                       const int32_t *n_pos, const uint64_t *n_pos_offsets, void *d_send, void *d_cursors)
{
    if (!c) return PBK_E_ARG;
    if (!read_offsets || (encoding != PBK_ENC_ASCII && encoding != PBK_ENC_PLATANUS)) return fail(c, PBK_E_ARG, "bad arguments");
    if (n_reads && read_offsets[0] != 0) return fail(c, PBK_E_ARG, "read_offsets[0] must be 0");
    const u64 n_bases = n_reads ? read_offsets[n_reads] : 0;
    if (n_bases && !bases) return fail(c, PBK_E_ARG, "bases is NULL");
    TRY(keyx_bind(c, d_send, d_cursors));
    int rc;
    if (n_reads == 0) {                                  // nothing to send, but the cursors must say so
        rc = cudaMemsetAsync(d_cursors, 0, (size_t)c->keyx_plan.n_buckets * 8, c->s_compute) == cudaSuccess &&
             cudaStreamSynchronize(c->s_compute) == cudaSuccess ? PBK_OK : fail(c, PBK_E_CUDA, "clearing the cursors failed");
    } else {
        rc = push_common(c, bases, nullptr, (const u64 *)read_offsets, nullptr, n_reads, n_bases, encoding, n_pos, (const u64 *)n_pos_offsets);
    }
    c->keyx_send = c->keyx_cursors = nullptr;
    return rc;
}

int pbk_keyx_partition_device(pbk_ctx *c, const void *d_bases, const void *d_read_offsets, uint64_t n_reads,
                              uint64_t n_bases, void *d_send, void *d_cursors)
{
    if (!c) return PBK_E_ARG;
    if (n_reads && (!d_read_offsets || (n_bases && !d_bases))) return fail(c, PBK_E_ARG, "NULL device pointer");
    TRY(keyx_bind(c, d_send, d_cursors));
    int rc;
    if (n_reads == 0) {
        rc = cudaMemsetAsync(d_cursors, 0, (size_t)c->keyx_plan.n_buckets * 8, c->s_compute) == cudaSuccess &&
             cudaStreamSynchronize(c->s_compute) == cudaSuccess ? PBK_OK : fail(c, PBK_E_CUDA, "clearing the cursors failed");
    } else {
        rc = push_common(c, nullptr, (const uint8_t *)d_bases, nullptr, (const u64 *)d_read_offsets, n_reads, n_bases, PBK_ENC_ASCII, nullptr, nullptr);
    }
    c->keyx_send = c->keyx_cursors = nullptr;
    return rc;
}

int pbk_keyx_partition_device_async(pbk_ctx *c, const void *d_bases, const void *d_read_offsets, uint64_t n_reads,
                                    uint64_t n_bases, void *d_send, void *d_cursors)
{
    if (!c) return PBK_E_ARG;
    // buffers that must grow are reallocated behind a stream synchronisation: fine, just not asynchronous that time
    c->keyx_async = n_reads != 0;
    const int rc = pbk_keyx_partition_device(c, d_bases, d_read_offsets, n_reads, n_bases, d_send, d_cursors);
    c->keyx_async = false;
    return rc;
}

int pbk_stream_signal(pbk_ctx *c, void *stream)
{
    if (!c) return PBK_E_ARG;
    CK(cudaSetDevice(c->device));
    if (!c->ev_signal) CK(cudaEventCreateWithFlags(&c->ev_signal, cudaEventDisableTiming));
    CK(cudaEventRecord(c->ev_signal, c->s_compute));
    CK(cudaStreamWaitEvent((cudaStream_t)stream, c->ev_signal, 0));
    return PBK_OK;
}

int pbk_stream_wait(pbk_ctx *c, void *stream)
{
    if (!c) return PBK_E_ARG;
    CK(cudaSetDevice(c->device));
    if (!c->ev_wait) CK(cudaEventCreateWithFlags(&c->ev_wait, cudaEventDisableTiming));
    CK(cudaEventRecord(c->ev_wait, (cudaStream_t)stream));
    CK(cudaStreamWaitEvent(c->s_compute, c->ev_wait, 0));
    return PBK_OK;
}

// Pass B of the key exchange over `srcs` (per source rank: its keys and fill counts for this shard)
static int keyx_insert_common(pbk_ctx *c, const KeyxSources &srcs)
{
    if (c->finalized) return fail(c, PBK_E_STATE, "insert after finalize");
    CK(cudaSetDevice(c->device));
    TRY(table_ordered(c));
    const u32 G = c->shard.n_shards, n_desc = c->keyx_plan.n_buckets, R = n_desc / G;
    const u64 seg_cap = c->keyx_plan.seg_cap;
    TRY(ensure_passb_buffers(c));
    // Steady state (an earlier insert of this context measured how many of the received keys are new and how many arrive):
    // size the table from that, build the tile map on the device, queue Pass B and return -- no host round trip per
    // received chunk.  A table that turns out too small only costs the overflow route (settle()).  PBK_KEYX_SYNC=1 keeps
    // the host-sized path below.
    static const bool always_sync = getenv("PBK_KEYX_SYNC") != nullptr;
    if (c->ratio_known && c->keyx_last_total && c->table.slots && c->pipeline_enabled && !always_sync) {
        const u64 expect = (u64)(c->keyx_last_total * std::min(1.0, c->new_ratio * 1.15)) + 65536;
        // The host's view of the table (c->occupied) is stale while launches are queued (a Pass A that returned without its
        // read-back, earlier inserts of this step).  Reading the counters back costs a stream synchronisation in the middle
        // of the step, so it only happens if the table might NOT have room even for everything queued so far.
        if (c->counters_pending && (double)(c->occupied + c->pending_new + expect) > max_load(c) * (double)c->table.capacity())
            TRY(settle(c));
        if (!c->counters_pending) c->pending_new = 0;
        TRY(maybe_clamp(c, c->keyx_last_total + c->keyx_last_total / 8));
        if ((double)(c->occupied + c->pending_new + expect) > max_load(c) * (double)c->table.capacity())
            TRY(grow_table(c, &c->table, c->occupied, (u64)((c->occupied + expect) / max_load(c)) + 1));
        c->pending_new += expect;
        TRY(ensure_overflow_for_batch(c, c->keyx_last_total + c->keyx_last_total / 8));
        const int second = passb2_run(c, &srcs, 0, n_desc);
        if (second < 0) return second;
        c->table_touched = true;
        if (second == 0) {
            Span sp(c, LC_INSERT);
            launch_bucket_insert_gathered_chained(srcs, seg_cap, c->d_passb, G, R, c->table, c->d_ctr, c->d_ovf, c->ovf_cap,
                                                  c->sm_count, c->s_compute);
        }
        CK(cudaGetLastError());
        c->counters_pending = true;
        c->n_pipelined += 1;
        return PBK_OK;
    }
    // fill counts, [source][region] -> descriptor order [region][source]
    for (u32 sr = 0; sr < G; ++sr)
        CK(cudaMemcpyAsync(c->h_bkt_cursor + (size_t)sr * R, srcs.cursors[sr], (size_t)R * 8, cudaMemcpyDefault, c->s_compute));
    CK(cudaStreamSynchronize(c->s_compute));
    c->d2h_bytes += (u64)n_desc * 8;
    std::vector<u64> cnt(n_desc);
    u64 total = 0;
    for (u32 j = 0; j < R; ++j)
        for (u32 sr = 0; sr < G; ++sr) { cnt[j * G + sr] = std::min<u64>(c->h_bkt_cursor[sr * R + j], seg_cap); total += cnt[j * G + sr]; }
    DBG("keyx insert: %llu keys from %u sources in %u regions (seg_cap %llu)", (unsigned long long)total, G, R, (unsigned long long)seg_cap);
    if (total == 0) return PBK_OK;
    if (!c->table.slots) {
        const u64 want = c->table_hint ? c->table_hint : (u64)(total * c->new_ratio / max_load(c)) + 1;
        TRY(table_alloc(c, &c->table, round_slots(c, want)));
    }
    TRY(maybe_clamp(c, total));
    TRY(ensure_overflow_for_batch(c, total));          // a first batch of mostly new keys can overrun the pilot's table regions
    auto room_for = [&](u64 expect_new) -> int {
        if ((double)(c->occupied + expect_new) > max_load(c) * (double)c->table.capacity())
            return grow_table(c, &c->table, c->occupied, (u64)((c->occupied + expect_new) / max_load(c)) + 1);
        return PBK_OK;
    };
    auto launch = [&](u32 d0, u32 d1) -> int {
        const int second = passb2_run(c, &srcs, d0, d1);
        if (second < 0) return second;
        c->table_touched = true;
        if (second == 0) {
            Span sp(c, LC_INSERT);
            launch_bucket_insert_gathered(srcs, seg_cap, cnt.data(), c->h_passb, c->d_passb, d0, d1, G, R, c->table,
                                          c->d_ctr, c->d_ovf, c->ovf_cap, c->sm_count, c->s_compute);
        }
        CK(cudaGetLastError());
        TRY(read_counters(c));                          // also makes h_passb reusable
        TRY(drain_overflow(c));
        return PBK_OK;
    };
    const u64 occ_before = c->occupied;
    TRY(room_for((u64)(total * std::min(1.0, c->new_ratio))));
    if (!c->ratio_known) {
        // first batch: the first 1/16 of the regions (statistically identical hash ranges) tell how many of the keys
        // are new, then the table is sized once for the rest -- like flush_buckets
        const u32 pilot = std::max<u32>(1, R / 16) * G;
        u64 pilot_keys = 0;
        for (u32 i = 0; i < pilot; ++i) pilot_keys += cnt[i];
        TRY(launch(0, pilot));
        const double per_key = pilot_keys ? (double)(c->occupied - occ_before) / (double)pilot_keys : c->new_ratio;
        if (pilot < n_desc) {
            TRY(room_for((u64)((double)(total - pilot_keys) * std::min(1.0, per_key * 1.05)) + 4096));
            TRY(launch(pilot, n_desc));
        }
    } else {
        TRY(launch(0, n_desc));
    }
    if (total > 4096) {
        c->new_ratio = std::max(0.01, std::min(1.0, (double)(c->occupied - occ_before) / (double)total));
        c->ratio_known = true;
        c->keyx_last_total = total;
    }
    return PBK_OK;
}

int pbk_keyx_insert_device(pbk_ctx *c, const void *d_recv, const void *d_recv_cursors)
{
    if (!c || !d_recv || !d_recv_cursors) return PBK_E_ARG;
    if (c->keyx_plan.n_buckets == 0 || c->W != 1) return fail(c, PBK_E_STATE, "pbk_keyx_insert_device before pbk_keyx_plan");
    const u32 G = c->shard.n_shards, R = c->keyx_plan.n_buckets / G;
    if (G > (u32)KEYX_MAX_SRC) return fail(c, PBK_E_ARG, "key exchange supports up to %d shards", KEYX_MAX_SRC);
    KeyxSources srcs{};
    for (u32 sr = 0; sr < G; ++sr) {
        srcs.keys[sr] = (const u64 *)d_recv + (u64)sr * R * c->keyx_plan.seg_cap;
        srcs.cursors[sr] = (const u64 *)d_recv_cursors + (u64)sr * R;
    }
    return keyx_insert_common(c, srcs);
}

// ---- key exchange, pull form: nobody sends anything ---------------------------------------------------------------------
// Pass A fills this context's own owner-major bucket store; every shard's Pass B then reads the segments addressed to it in
// place, out of its peers' HBM over NVLink (P2P loads through peer-mapped pointers), so the transfer is Pass B's own streamed
// loads and overlaps its atomics tile by tile.  What remains between the two passes is a barrier, not a data movement.

int pbk_keyx_pull_setup(pbk_ctx *c, uint64_t max_windows_any_rank, pbk_keyx_layout *out)
{
    if (!c || !out) return PBK_E_ARG;
    TRY(pbk_keyx_plan(c, max_windows_any_rank, out));
    if (c->shard.n_shards > (u32)KEYX_MAX_SRC) return fail(c, PBK_E_ARG, "the pull exchange supports up to %d shards", KEYX_MAX_SRC);
    CK(cudaSetDevice(c->device));
    const size_t keys_bytes = (size_t)c->keyx_plan.n_buckets * c->keyx_plan.seg_cap * 8;
    const size_t cur_bytes = ((size_t)c->keyx_plan.n_buckets * 8 + 255) & ~(size_t)255;
    const size_t need = 2 * keys_bytes + 2 * cur_bytes;
    if (c->pull_base && (keys_bytes != c->pull_keys_bytes || cur_bytes != c->pull_cur_bytes))
        return fail(c, PBK_E_STATE, "the pull store is already shared with another layout (destroy the context to change it)");
    if (!c->pull_base) {
        void *p = nullptr;
        TRY(dev_alloc(c, &p, need));                    // plain cudaMalloc: exportable with cudaIpcGetMemHandle
        c->pull_base = (char *)p; c->pull_bytes = need; c->pull_keys_bytes = keys_bytes; c->pull_cur_bytes = cur_bytes;
        CK(cudaMemsetAsync(c->pull_base + 2 * keys_bytes, 0, 2 * cur_bytes, c->s_compute));
        CK(cudaStreamSynchronize(c->s_compute));
        c->pull_peer[c->shard.rank] = c->pull_base;
        c->pull_parity = 1;
    }
    return PBK_OK;
}

int pbk_keyx_pull_handle(pbk_ctx *c, void *handle_out)
{
    if (!c || !handle_out) return PBK_E_ARG;
    if (!c->pull_base) return fail(c, PBK_E_STATE, "pbk_keyx_pull_handle before pbk_keyx_pull_setup");
    CK(cudaSetDevice(c->device));
    static_assert(sizeof(cudaIpcMemHandle_t) <= PBK_KEYX_HANDLE_BYTES, "handle size");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, c->pull_base));
    memset(handle_out, 0, PBK_KEYX_HANDLE_BYTES);
    memcpy(handle_out, &h, sizeof h);
    return PBK_OK;
}

int pbk_keyx_pull_connect_ipc(pbk_ctx *c, uint32_t src_rank, const void *handle)
{
    if (!c || !handle) return PBK_E_ARG;
    if (!c->pull_base) return fail(c, PBK_E_STATE, "pbk_keyx_pull_connect_ipc before pbk_keyx_pull_setup");
    if (src_rank >= c->shard.n_shards || src_rank == c->shard.rank) return fail(c, PBK_E_ARG, "bad source rank %u", src_rank);
    CK(cudaSetDevice(c->device));
    if (c->pull_peer[src_rank] && c->pull_peer_ipc[src_rank]) { cudaIpcCloseMemHandle(c->pull_peer[src_rank]); c->pull_peer[src_rank] = nullptr; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    void *p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->pull_peer[src_rank] = (char *)p; c->pull_peer_ipc[src_rank] = true;
    return PBK_OK;
}

int pbk_keyx_pull_connect_local(pbk_ctx *c, uint32_t src_rank, pbk_ctx *peer)
{
    if (!c || !peer) return PBK_E_ARG;
    if (!c->pull_base || !peer->pull_base) return fail(c, PBK_E_STATE, "pbk_keyx_pull_connect_local before pbk_keyx_pull_setup on both contexts");
    if (src_rank >= c->shard.n_shards || src_rank == c->shard.rank || peer->shard.rank != src_rank || peer->shard.n_shards != c->shard.n_shards)
        return fail(c, PBK_E_ARG, "bad source rank %u", src_rank);
    if (peer->pull_keys_bytes != c->pull_keys_bytes) return fail(c, PBK_E_ARG, "the two contexts planned different layouts");
    CK(cudaSetDevice(c->device));
    if (peer->device != c->device) {
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, c->device, peer->device));
        if (!can) return fail(c, PBK_E_CUDA, "GPU %d cannot map the memory of GPU %d (no peer access)", c->device, peer->device);
        const cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
        cudaGetLastError();
    }
    c->pull_peer[src_rank] = peer->pull_base; c->pull_peer_ipc[src_rank] = false;
    return PBK_OK;
}

// forget the store and every mapping (a larger layout is wanted: pbk_group_push_reads; the peers must release theirs too
// before anybody partitions again)
int pbk_keyx_pull_release(pbk_ctx *c)
{
    if (!c) return PBK_E_ARG;
    if (!c->pull_base) return PBK_OK;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->s_compute));
    for (int i = 0; i < KEYX_MAX_SRC; ++i) {
        if (c->pull_peer_ipc[i] && c->pull_peer[i]) cudaIpcCloseMemHandle(c->pull_peer[i]);
        c->pull_peer[i] = nullptr; c->pull_peer_ipc[i] = false;
    }
    dev_free(c, c->pull_base, c->pull_bytes);
    c->pull_base = nullptr; c->pull_bytes = c->pull_keys_bytes = c->pull_cur_bytes = 0;
    c->pull_parity = 1;
    return PBK_OK;
}

static int pull_bind(pbk_ctx *c)
{
    if (!c->pull_base) return fail(c, PBK_E_STATE, "pbk_keyx_pull_partition before pbk_keyx_pull_setup");
    c->pull_parity ^= 1;
    return PBK_OK;
}
static void *pull_keys(pbk_ctx *c) { return c->pull_base + (size_t)c->pull_parity * c->pull_keys_bytes; }
static void *pull_cursors(pbk_ctx *c) { return c->pull_base + 2 * c->pull_keys_bytes + (size_t)c->pull_parity * c->pull_cur_bytes; }

int pbk_keyx_pull_partition(pbk_ctx *c, const uint8_t *bases, const uint64_t *read_offsets, uint64_t n_reads, int encoding,
                            const int32_t *n_pos, const uint64_t *n_pos_offsets)
{
    if (!c) return PBK_E_ARG;
    TRY(pull_bind(c));
    return pbk_keyx_partition(c, bases, read_offsets, n_reads, encoding, n_pos, n_pos_offsets, pull_keys(c), pull_cursors(c));
}

int pbk_keyx_pull_partition_device(pbk_ctx *c, const void *d_bases, const void *d_read_offsets, uint64_t n_reads, uint64_t n_bases,
                                   int async)
{
    if (!c) return PBK_E_ARG;
    TRY(pull_bind(c));
    return (async ? pbk_keyx_partition_device_async : pbk_keyx_partition_device)(c, d_bases, d_read_offsets, n_reads, n_bases,
                                                                                 pull_keys(c), pull_cursors(c));
}

// Queue (on the context's stream) a copy of the number of records this context has staged since pbk_reset -- keys that found
// their segment full in Pass A -- into an 8-byte device word of the caller.  Summed over the ranks by the collective that also
// serves as the barrier between Pass A and Pass B, it tells every rank whether the record route has to run at all, without a
// collective and a host round trip of its own.
int pbk_keyx_staged_count_device(pbk_ctx *c, void *d_count_u64)
{
    if (!c || !d_count_u64) return PBK_E_ARG;
    CK(cudaSetDevice(c->device));
    // two consecutive counters: records already in the staging table, and keys still on the overflow list (a full segment in
    // Pass A puts them there; the host moves them to the staging table when it next reads the counters back)
    static_assert(offsetof(Counters, overflow_n) == offsetof(Counters, new_keys_remote) + 8, "counter layout");
    CK(cudaMemcpyAsync(d_count_u64, &c->d_ctr->new_keys_remote, 16, cudaMemcpyDeviceToDevice, c->s_compute));
    return PBK_OK;
}

int pbk_keyx_pull_insert(pbk_ctx *c)
{
    if (!c) return PBK_E_ARG;
    if (!c->pull_base || c->W != 1) return fail(c, PBK_E_STATE, "pbk_keyx_pull_insert before pbk_keyx_pull_setup");
    const u32 G = c->shard.n_shards, R = c->keyx_plan.n_buckets / G, me = c->shard.rank;
    KeyxSources srcs{};
    for (u32 sr = 0; sr < G; ++sr) {
        const char *base = c->pull_peer[sr];
        if (!base) return fail(c, PBK_E_STATE, "source rank %u is not connected (pbk_keyx_pull_connect_*)", sr);
        srcs.keys[sr] = (const u64 *)(base + (size_t)c->pull_parity * c->pull_keys_bytes) + (u64)me * R * c->keyx_plan.seg_cap;
        srcs.cursors[sr] = (const u64 *)(base + 2 * c->pull_keys_bytes + (size_t)c->pull_parity * c->pull_cur_bytes) + (u64)me * R;
    }
    return keyx_insert_common(c, srcs);
}

uint32_t pbk_shard_of_key(const uint64_t *key_words, uint32_t k, uint32_t n_shards)
{
    if (!key_words || n_shards <= 1) return 0;
    u64 h = 0;
    switch ((k + 31) / 32) {
    case 1: h = hash_key<1>((const u64 *)key_words); break;
    case 2: h = hash_key<2>((const u64 *)key_words); break;
    case 3: h = hash_key<3>((const u64 *)key_words); break;
    case 4: h = hash_key<4>((const u64 *)key_words); break;
    case 5: h = hash_key<5>((const u64 *)key_words); break;
    case 6: h = hash_key<6>((const u64 *)key_words); break;
    case 7: h = hash_key<7>((const u64 *)key_words); break;
    case 8: h = hash_key<8>((const u64 *)key_words); break;
    default: return 0;
    }
    return shard_of_hash(h, n_shards);
}

// ---- microbenchmark ---------------------------------------------------------------------------

int pbk_microbench_atomics(int device, uint64_t table_bytes, uint64_t n_ops, int mode, double *ops_per_s)
{
    if (!ops_per_s || table_bytes < 4096 || n_ops == 0 || mode < 0 || (mode > 11 && mode < 100) || mode > 200) return PBK_E_ARG;
#ifdef PBK_CPU_EMUL
    return PBK_E_NO_DEVICE;                          // (host emulation of the ABI for the CPU tests: nothing to measure)
#endif
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) { cudaGetLastError(); return PBK_E_NO_DEVICE; }
    if (device >= 0 && cudaSetDevice(device) != cudaSuccess) return PBK_E_NO_DEVICE;
    cudaDeviceProp prop; int dev = 0;
    cudaGetDevice(&dev); cudaGetDeviceProperties(&prop, dev);
    int log2slots = 0;
    while ((32ull << log2slots) <= table_bytes) ++log2slots;     // 16-byte slots, largest power of two that fits
    TableView t{nullptr, 2ull << log2slots, 1};                 // bytes = 16 << log2slots
    if (getenv("PBK_L2_FETCH32")) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
    const size_t extra = (mode >= 6 && (mode < 100 || (mode & 3) >= 2)) ? (size_t)1 << 30 : 0;       // modes 6-8 read keys from a 1 GiB stream behind the table
    if (cudaMalloc(&t.slots, t.bytes() + extra) != cudaSuccess) { cudaGetLastError(); return PBK_E_NOMEM; }
    if (extra) cudaMemset((char *)t.slots + t.bytes(), 0x5A, extra);
    cudaStream_t st; cudaStreamCreate(&st);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {                            // rep 0 warms up
        if (mode == 2 || rep == 0) launch_table_init(t, st);
        cudaEventRecord(a, st);
        launch_microbench(t.slots, log2slots, n_ops, mode, 0x9E3779B97F4A7C15ull * (rep + 1), prop.multiProcessorCount, st);
        cudaEventRecord(b, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) { cudaFree(t.slots); return PBK_E_CUDA; }
        float ms = 0; cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best) best = ms;
    }
    *ops_per_s = (double)n_ops / (best * 1e-3);
    cudaEventDestroy(a); cudaEventDestroy(b); cudaStreamDestroy(st); cudaFree(t.slots);
    return PBK_OK;
}

}  // extern "C"
