// pbk_group.cu -- several GPUs of one box behind ONE handle of the C ABI (include/pbk.h, pbk_group_*): what SURVEY.md section 8e
// calls hash-range sharding, for a host program that is a single process (the reference's `assemble` is one; `iterate` spawns it).
//
// The group owns one pbk_ctx per device.  Keys are owned by hash range (pbk_shard_of_key).  A batch of reads is cut into one
// contiguous slice per device; one host thread per device drives its context, so H2D copies and kernels of all devices run
// concurrently.  k <= 32: the pull exchange -- every device runs Pass A into its own owner-major store, then every device's Pass B
// reads the segments it owns straight out of its peers' HBM over NVLink (peer access between the contexts of this process,
// pbk_keyx_pull_connect_local); no collective library is involved.  k > 32: every device counts its slice (own keys into its table,
// foreign keys pre-aggregated in its staging table), then the (key, count) records move with cudaMemcpyPeer.
// Results (histograms, exported entries) are merged on the host and are identical to a single-GPU count.
#include "../../include/pbk.h"
#include "pbk_kernels.cuh"

#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

struct pbk_group {
    std::vector<pbk_ctx *> ctx;
    std::vector<int> device;
    uint32_t k = 0; int W = 0;
    uint64_t pull_max_windows = 0;            // what the pull stores were planned for (0 = not set up)
    std::vector<void *> d_rec;                // k > 32 and spilled keys: packed staged records per device
    std::vector<uint64_t> d_rec_cap;
    bool finalized = false;
    std::vector<uint64_t> occ_hist;
    std::string err;
};

namespace {

int gfail(pbk_group *g, int code, const std::string &msg) { if (g) g->err = msg; return code; }

// run f(i) for every member on its own host thread; first failure wins
template <typename F> int for_all(pbk_group *g, F f)
{
    const size_t n = g->ctx.size();
    std::vector<int> rc(n, PBK_OK);
#ifdef PBK_CPU_EMUL
    const bool sequential = true;             // tests/cpu_emul: kernels run on the calling thread with static "shared memory"
#else
    const bool sequential = n == 1;
#endif
    if (sequential) { for (size_t i = 0; i < n; ++i) rc[i] = f(i); }
    else {
        std::vector<std::thread> th;
        for (size_t i = 0; i < n; ++i) th.emplace_back([&, i]() { rc[i] = f(i); });
        for (auto &t : th) t.join();
    }
    for (size_t i = 0; i < n; ++i)
        if (rc[i] != PBK_OK) return gfail(g, rc[i], std::string("device ") + std::to_string(g->device[i]) + ": " + pbk_last_error(g->ctx[i]));
    return PBK_OK;
}

// the staged (key, count) records of every member go to their owners: pack on the source, cudaMemcpyPeer, weighted insert
int exchange_records(pbk_group *g)
{
    const uint32_t n = (uint32_t)g->ctx.size();
    const size_t rec_bytes = (size_t)(g->W + 1) * 8;
    std::vector<std::vector<uint64_t>> cnt(n, std::vector<uint64_t>(n, 0));
    int rc = for_all(g, [&](size_t i) { return pbk_shard_send_counts(g->ctx[i], cnt[i].data()); });
    if (rc != PBK_OK) return rc;
    uint64_t total = 0;
    for (auto &c : cnt) for (uint64_t v : c) total += v;
    if (total == 0) return PBK_OK;
    rc = for_all(g, [&](size_t i) -> int {
        uint64_t mine = 0;
        for (uint64_t v : cnt[i]) mine += v;
        if (mine == 0) return PBK_OK;
        if (cudaSetDevice(g->device[i]) != cudaSuccess) return PBK_E_CUDA;
        if (mine + 1 > g->d_rec_cap[i]) {
            cudaFree(g->d_rec[i]); g->d_rec[i] = nullptr; g->d_rec_cap[i] = 0;
            const uint64_t cap = mine + mine / 4 + 1024;
            if (cudaMalloc(&g->d_rec[i], cap * rec_bytes) != cudaSuccess) { cudaGetLastError(); return PBK_E_NOMEM; }
            g->d_rec_cap[i] = cap;
        }
        return pbk_shard_pack_device(g->ctx[i], g->d_rec[i], g->d_rec_cap[i]);
    });
    if (rc != PBK_OK) return rc;
    return for_all(g, [&](size_t dst) -> int {
        if (cudaSetDevice(g->device[dst]) != cudaSuccess) return PBK_E_CUDA;
        for (uint32_t src = 0; src < n; ++src) {
            const uint64_t m = cnt[src][dst];
            if (src == dst || m == 0) continue;
            uint64_t before = 0;
            for (uint32_t d = 0; d < dst; ++d) before += cnt[src][d];
            void *tmp = nullptr;
            if (cudaMalloc(&tmp, m * rec_bytes) != cudaSuccess) { cudaGetLastError(); return PBK_E_NOMEM; }
            // (cudaMemcpyPeer itself runs on the legacy default stream, which the contexts' non-blocking streams do not wait
            //  for: the copy gets its own stream and the host waits for it before the insert is queued)
            cudaStream_t cs = nullptr;
            cudaError_t e = cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaMemcpyPeerAsync(tmp, g->device[dst], (const char *)g->d_rec[src] + before * rec_bytes, g->device[src], m * rec_bytes, cs);
            if (e == cudaSuccess) e = cudaStreamSynchronize(cs);
            if (cs) cudaStreamDestroy(cs);
            int r = e == cudaSuccess ? pbk_shard_insert_device(g->ctx[dst], tmp, m) : PBK_E_CUDA;
            cudaFree(tmp);
            if (r != PBK_OK) return r;
        }
        return PBK_OK;
    });
}

}  // namespace

extern "C" {

int pbk_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int pbk_group_create(pbk_group **out, const pbk_config *cfg, const int32_t *devices, uint32_t n_devices)
{
    if (!out || !cfg || cfg->struct_size < PBK_CONFIG_SIZE_V1 || n_devices == 0 || n_devices > 16) return PBK_E_ARG;
    pbk_config cfg_full;                                            // (a shorter, earlier layout: the new fields are 0)
    memset(&cfg_full, 0, sizeof cfg_full);
    memcpy(&cfg_full, cfg, std::min<size_t>(cfg->struct_size, sizeof cfg_full));
    cfg_full.n_passes = cfg_full.pass_index = 0;                    // hash-range passes are a one-GPU mode
    cfg = &cfg_full;
    *out = nullptr;
    pbk_group *g = new (std::nothrow) pbk_group();
    if (!g) return PBK_E_NOMEM;
    g->k = cfg->k; g->W = (int)((cfg->k + 31) / 32);
    g->ctx.assign(n_devices, nullptr); g->device.assign(n_devices, 0);
    g->d_rec.assign(n_devices, nullptr); g->d_rec_cap.assign(n_devices, 0);
    for (uint32_t i = 0; i < n_devices; ++i) g->device[i] = devices ? devices[i] : (int)i;
    // contexts are created concurrently: the CUDA context of a device costs about a second
    const int rc = for_all(g, [&](size_t i) -> int {
        pbk_config c = *cfg;
        c.struct_size = sizeof c; c.device = g->device[i];
        c.n_shards = n_devices > 1 ? n_devices : 0; c.shard_rank = (uint32_t)i;
        return pbk_create(&g->ctx[i], &c);
    });
    if (rc != PBK_OK) {
        // (for_all read pbk_last_error of a context that may not exist: report the status only)
        for (auto c : g->ctx) if (c) pbk_destroy(c);
        delete g;
        return rc;
    }
    *out = g;
    return PBK_OK;
}

void pbk_group_destroy(pbk_group *g)
{
    if (!g) return;
    for (size_t i = 0; i < g->ctx.size(); ++i) {
        if (g->d_rec[i]) { cudaSetDevice(g->device[i]); cudaFree(g->d_rec[i]); }
        if (g->ctx[i]) pbk_destroy(g->ctx[i]);
    }
    delete g;
}

uint32_t pbk_group_size(const pbk_group *g) { return g ? (uint32_t)g->ctx.size() : 0; }
pbk_ctx *pbk_group_member(pbk_group *g, uint32_t i) { return g && i < g->ctx.size() ? g->ctx[i] : nullptr; }
const char *pbk_group_last_error(const pbk_group *g) { return g ? g->err.c_str() : ""; }

int pbk_group_reset(pbk_group *g, uint32_t k)
{
    if (!g) return PBK_E_ARG;
    if (k == 0) k = g->k;
    const int W = (int)((k + 31) / 32);
    const int rc = for_all(g, [&](size_t i) { return pbk_reset(g->ctx[i], k); });
    if (rc != PBK_OK) return rc;
    g->k = k; g->W = W; g->finalized = false; g->occ_hist.clear();
    return PBK_OK;
}

int pbk_group_push_reads(pbk_group *g, const uint8_t *bases, const uint64_t *read_offsets, uint64_t n_reads, int encoding,
                         const int32_t *n_pos, const uint64_t *n_pos_offsets)
{
    if (!g) return PBK_E_ARG;
    if (n_reads == 0) return PBK_OK;
    if (!read_offsets || read_offsets[0] != 0) return gfail(g, PBK_E_ARG, "bad read_offsets");
    if (g->finalized) return gfail(g, PBK_E_STATE, "pbk_group_push_reads after pbk_group_finalize (call pbk_group_reset first)");
    const uint32_t n = (uint32_t)g->ctx.size();
    if (n == 1) {
        const int rc = pbk_push_reads(g->ctx[0], bases, read_offsets, n_reads, encoding, n_pos, n_pos_offsets);
        return rc == PBK_OK ? rc : gfail(g, rc, pbk_last_error(g->ctx[0]));
    }
    // one contiguous slice of the batch per device, cut at read boundaries, about equal in bases
    const uint64_t total = read_offsets[n_reads];
    std::vector<uint64_t> cut(n + 1, n_reads);
    cut[0] = 0;
    for (uint32_t i = 1; i < n; ++i)
        cut[i] = (uint64_t)(std::lower_bound(read_offsets, read_offsets + n_reads + 1, total / n * i) - read_offsets);
    for (uint32_t i = 1; i <= n; ++i) cut[i] = std::max(cut[i], cut[i - 1]);
    std::vector<std::vector<uint64_t>> off(n), npo(n);
    uint64_t max_windows = 1;
    for (uint32_t i = 0; i < n; ++i) {
        const uint64_t r0 = cut[i], r1 = cut[i + 1];
        off[i].resize(r1 - r0 + 1);
        for (uint64_t r = r0; r <= r1; ++r) off[i][r - r0] = read_offsets[r] - read_offsets[r0];
        if (encoding == PBK_ENC_PLATANUS && n_pos_offsets) {
            npo[i].resize(r1 - r0 + 1);
            for (uint64_t r = r0; r <= r1; ++r) npo[i][r - r0] = n_pos_offsets[r] - n_pos_offsets[r0];
        }
        const uint64_t nb = off[i].back(), nr = r1 - r0;
        if (nb > nr * (g->k - 1)) max_windows = std::max(max_windows, nb - nr * (g->k - 1));
    }
    auto slice_bases = [&](uint32_t i) { return bases + read_offsets[cut[i]]; };
    auto slice_npos = [&](uint32_t i) { return (encoding == PBK_ENC_PLATANUS && n_pos && n_pos_offsets) ? n_pos + n_pos_offsets[cut[i]] : nullptr; };

    if (g->W == 1) {
        // ---- pull exchange -------------------------------------------------------------------------------------------
        if (max_windows > g->pull_max_windows) {                  // (first batch, or a larger one than planned for)
            const uint64_t plan = max_windows + max_windows / 4;
            int rc = for_all(g, [&](size_t i) -> int {
                pbk_keyx_layout lay;
                const int r = pbk_keyx_pull_release(g->ctx[i]);
                return r != PBK_OK ? r : pbk_keyx_pull_setup(g->ctx[i], plan, &lay);
            });
            if (rc == PBK_OK)
                rc = for_all(g, [&](size_t i) -> int {
                    for (uint32_t s = 0; s < n; ++s)
                        if (s != i) { const int r = pbk_keyx_pull_connect_local(g->ctx[i], s, g->ctx[s]); if (r != PBK_OK) return r; }
                    return PBK_OK;
                });
            if (rc != PBK_OK) return rc;
            g->pull_max_windows = plan;
        }
        // Pass A everywhere (each call returns when its store is complete: the join is the barrier) ...
        int rc = for_all(g, [&](size_t i) {
            return pbk_keyx_pull_partition(g->ctx[i], slice_bases((uint32_t)i), off[i].data(), off[i].size() - 1, encoding,
                                           slice_npos((uint32_t)i), npo[i].empty() ? nullptr : npo[i].data());
        });
        if (rc != PBK_OK) return rc;
        // ... then every device's Pass B over what all devices hold for it
        rc = for_all(g, [&](size_t i) { return pbk_keyx_pull_insert(g->ctx[i]); });
        if (rc != PBK_OK) return rc;
        return exchange_records(g);                               // keys that found their segment full (rare)
    }
    // ---- k > 32: count locally, exchange pre-aggregated records ------------------------------------------------------
    int rc = for_all(g, [&](size_t i) {
        return pbk_push_reads(g->ctx[i], slice_bases((uint32_t)i), off[i].data(), off[i].size() - 1, encoding, slice_npos((uint32_t)i),
                              npo[i].empty() ? nullptr : npo[i].data());
    });
    if (rc != PBK_OK) return rc;
    return exchange_records(g);
}

int pbk_group_finalize(pbk_group *g, uint64_t *occ_hist, uint64_t *len_hist, uint64_t *n_distinct, uint64_t *n_instances,
                       uint64_t *max_occurrence)
{
    if (!g) return PBK_E_ARG;
    const size_t n = g->ctx.size();
    std::vector<std::vector<uint64_t>> occ(n, std::vector<uint64_t>(PBK_OCC_BINS, 0)), len(n);
    std::vector<uint64_t> nd(n, 0), ni(n, 0), mx(n, 0);
    if (len_hist) for (auto &l : len) l.assign(PBK_LEN_BINS, 0);
    const int rc = for_all(g, [&](size_t i) {
        return pbk_finalize(g->ctx[i], occ[i].data(), len_hist ? len[i].data() : nullptr, &nd[i], &ni[i], &mx[i]);
    });
    if (rc != PBK_OK) return rc;
    g->occ_hist.assign(PBK_OCC_BINS, 0);
    uint64_t d = 0, inst = 0, m = 0;
    for (size_t i = 0; i < n; ++i) {                              // shards own disjoint key sets: the histograms simply add up
        for (uint32_t b = 0; b < PBK_OCC_BINS; ++b) g->occ_hist[b] += occ[i][b];
        d += nd[i]; inst += ni[i];
    }
    for (uint32_t b = 1; b < PBK_OCC_BINS; ++b) if (g->occ_hist[b]) m = b;
    if (occ_hist) memcpy(occ_hist, g->occ_hist.data(), PBK_OCC_BINS * 8);
    if (len_hist) {
        memset(len_hist, 0, PBK_LEN_BINS * 8);
        for (size_t i = 0; i < n; ++i) for (uint32_t b = 0; b < PBK_LEN_BINS; ++b) len_hist[b] += len[i][b];
    }
    if (n_distinct) *n_distinct = d;
    if (n_instances) *n_instances = inst;
    if (max_occurrence) *max_occurrence = m;
    g->finalized = true;
    return PBK_OK;
}

int pbk_group_export(pbk_group *g, uint32_t min_count, int sorted, uint64_t *keys, uint16_t *counts, uint64_t capacity, uint64_t *n_out)
{
    if (!g) return PBK_E_ARG;
    if (!g->finalized) return gfail(g, PBK_E_STATE, "pbk_group_export before pbk_group_finalize");
    const size_t n = g->ctx.size();
    const size_t W = (size_t)g->W;
    std::vector<uint64_t> cnt(n, 0);
    int rc = for_all(g, [&](size_t i) { return pbk_export(g->ctx[i], min_count, 0, nullptr, nullptr, 0, &cnt[i]); });
    if (rc != PBK_OK) return rc;
    uint64_t total = 0;
    for (uint64_t c : cnt) total += c;
    if (n_out) *n_out = total;
    if (capacity == 0) return PBK_OK;
    if (capacity < total || !keys || !counts) return gfail(g, PBK_E_ARG, "export needs room for " + std::to_string(total) + " entries");
    if (total == 0) return PBK_OK;
    if (n == 1) return pbk_export(g->ctx[0], min_count, sorted, keys, counts, capacity, n_out);
    // every shard exports (sorted on its GPU) into its own piece of a scratch array; shards are hash ranges, not key ranges, so the
    // ascending list is a merge of the pieces
    std::vector<uint64_t> kbuf(sorted ? total * W : 0);
    std::vector<uint16_t> cbuf(sorted ? total : 0);
    std::vector<uint64_t> start(n + 1, 0);
    for (size_t i = 0; i < n; ++i) start[i + 1] = start[i] + cnt[i];
    uint64_t *kdst = sorted ? kbuf.data() : keys;
    uint16_t *cdst = sorted ? cbuf.data() : counts;
    rc = for_all(g, [&](size_t i) -> int {
        if (cnt[i] == 0) return PBK_OK;
        uint64_t got = 0;
        return pbk_export(g->ctx[i], min_count, sorted, kdst + start[i] * W, cdst + start[i], cnt[i], &got);
    });
    if (rc != PBK_OK || !sorted) return rc;
    auto less = [&](const uint64_t *a, const uint64_t *b) {       // reference order: top word first (binstr.h:460-466)
        for (size_t j = W; j-- > 0;) if (a[j] != b[j]) return a[j] < b[j];
        return false;
    };
    std::vector<uint64_t> at(start.begin(), start.end() - 1);
    for (uint64_t o = 0; o < total; ++o) {                        // n-way merge (n <= 16: a linear scan of the heads)
        size_t best = n;
        for (size_t i = 0; i < n; ++i)
            if (at[i] < start[i + 1] && (best == n || less(&kbuf[at[i] * W], &kbuf[at[best] * W]))) best = i;
        memcpy(keys + o * W, &kbuf[at[best] * W], W * 8);
        counts[o] = cbuf[at[best]];
        ++at[best];
    }
    return PBK_OK;
}

// the shards answer for the neighbours they own: the flags of a key are the OR over the devices
int pbk_group_neighbor_flags(pbk_group *g, uint32_t min_count, const uint64_t *keys, uint64_t n, uint8_t *flags_out)
{
    if (!g) return PBK_E_ARG;
    if (n == 0) return PBK_OK;
    if (!keys || !flags_out) return gfail(g, PBK_E_ARG, "NULL argument");
    const size_t m = g->ctx.size();
    std::vector<std::vector<uint8_t>> part(m, std::vector<uint8_t>(n, 0));
    const int rc = for_all(g, [&](size_t i) { return pbk_neighbor_flags(g->ctx[i], min_count, keys, n, part[i].data()); });
    if (rc != PBK_OK) return rc;
    memset(flags_out, 0, n);
    for (size_t i = 0; i < m; ++i) for (uint64_t j = 0; j < n; ++j) flags_out[j] |= part[i][j];
    return PBK_OK;
}

}  // extern "C"
