// pbk_host.cpp -- the parts of the counting path that stay on the host:
// histogram statistics, the coverage cutoff rule, PREFIX_<k>merFrq.tsv and PREFIX_kmer_occ.bin.
// They are O(65535) or O(#kept k-mers) and serial in the reference as well; written from the
// behaviour of the cited reference lines, independently of oracle/.
#include "../../include/pbk.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

namespace {

typedef unsigned long long u64;

// sizeof(KEY) of the type Assemble::exec instantiates for k (assemble.cpp:174-185):
// u64 | Binstr63/95/127/159 = {vptr, value*, len, entity[2..5]} | binstr_t = {vptr, value*, len}
u64 key_raw_size(uint32_t k)
{
    if (k <= 32) return 8;
    if (k <= 64) return 24 + 16;
    if (k <= 96) return 24 + 24;
    if (k <= 128) return 24 + 32;
    if (k <= 160) return 24 + 40;
    return 24;
}

// DoubleHash::calcLength (doubleHash.h:107-115)
u64 calc_length(u64 len)
{
    for (u64 i = 1; i < 64; ++i)
        if ((len >> i) == 0) return i;
    return 64;
}

// LSD radix sort of (slot, index) pairs by slot
void sort_by_slot(std::vector<u64> &slot, std::vector<u64> &idx, u64 max_slot)
{
    const size_t n = slot.size();
    std::vector<u64> s2(n), i2(n);
    int bits = 1;
    while (bits < 64 && (max_slot >> bits)) ++bits;
    for (int shift = 0; shift < bits; shift += 11) {
        size_t hist[2049] = {0};
        for (size_t i = 0; i < n; ++i) ++hist[((slot[i] >> shift) & 2047) + 1];
        for (int b = 0; b < 2048; ++b) hist[b + 1] += hist[b];
        for (size_t i = 0; i < n; ++i) {
            size_t at = hist[(slot[i] >> shift) & 2047]++;
            s2[at] = slot[i]; i2[at] = idx[i];
        }
        slot.swap(s2); idx.swap(i2);
    }
}

}  // namespace

extern "C" {

// Counter::getLeftLocalMinimalValue (counter.h:245-267)
uint64_t pbk_left_local_min(const uint64_t *occ, uint64_t max_occ, uint64_t w)
{
    if (max_occ <= w) return 0;
    const u64 n = max_occ - w + 2;
    u64 prev = 0, i;
    for (i = 0; i < w; ++i) prev += occ[1 + i];
    for (i = 2; i < n; ++i) {
        const u64 cur = prev - occ[i - 1] + occ[i + w - 1];
        if (cur >= prev) break;
        prev = cur;
    }
    return i <= max_occ ? i - 1 + w / 2 : 1 + w / 2;
}

// assemble.cpp:318-321 with SMOOTHING_WINDOW = 1 (assemble.cpp:42)
uint64_t pbk_coverage_cutoff(const uint64_t *occ, uint64_t max_occ, int n_opt, int repeat)
{
    if (n_opt != 0) return (uint64_t)(long long)n_opt;
    u64 v = pbk_left_local_min(occ, max_occ, 1);
    if (!repeat) v /= 2;
    return std::max<u64>(v, 2);
}

// Counter::calcDistributionAverage (counter.h:221-238)
int pbk_distribution_average(const uint64_t *dist, uint64_t n_bins, uint64_t start, uint64_t end, double *out)
{
    if (!dist || !out) return PBK_E_ARG;
    if (end > n_bins || start > end) return PBK_E_KMER_DIST;
    u64 sum = 0, num = 0;
    for (u64 i = start; i <= end; ++i) { sum += i * dist[i]; num += dist[i]; }
    if (num == 0) return PBK_E_KMER_DIST;
    *out = static_cast<double>(sum) / num;
    return PBK_OK;
}

// counter.h:300-309: 2^floor(log2(memory / sizeof(pair<KEY, unsigned short>))) through double log/pow
uint64_t pbk_double_hash_size(uint64_t memory_bytes, uint32_t k)
{
    long base = (long)(key_raw_size(k) + 8);
    u64 tmp = memory_bytes / (u64)base;
    base = (long)(std::log((double)tmp) / std::log(2.0));
    tmp = (u64)std::pow(2.0, (double)base);
    while (tmp > memory_bytes) tmp >>= 1;
    return tmp;
}

// Counter::outputOccurrenceDistribution (counter.h:1000-1007)
int pbk_write_frq_tsv(const char *path, const uint64_t *occ, uint64_t max_occ)
{
    if (!path || !occ) return PBK_E_ARG;
    FILE *fp = fopen(path, "w");
    if (!fp) return PBK_E_IO;
    for (u64 i = 1; i <= max_occ; ++i) fprintf(fp, "%llu\t%llu\n", i, (u64)occ[i]);
    return fclose(fp) == 0 ? PBK_OK : PBK_E_IO;
}

// loadKmer (counter.h:600-640) + outputOccurrenceTableBinary (counter.h:955-963) + writeTable
// (doubleHash.h:266-278).  The reference materialises a zero-filled table of max(size, doubleHashSize)
// pair<KEY,u16> slots (8.6 GB at -m 16) only to learn each record's slot; an occupancy bitmap gives
// the same placement.
int pbk_write_kmer_occ_bin(const char *path, uint32_t k, const uint64_t *keys, const uint16_t *counts,
                           uint64_t n, uint64_t double_hash_size, uint64_t *load_size_out)
{
    if (!path || k == 0 || (n && (!keys || !counts))) return PBK_E_ARG;
    const unsigned words = (k + 31) / 32;
    // counter.h:621-622
    u64 size = (u64)(std::log((double)n / 0.9) / std::log(2.0));
    size = (u64)std::pow(2.0, (double)(size + 1));
    if (load_size_out) *load_size_out = size;
    const u64 slots = std::max<u64>(size, double_hash_size);            // counter.h:627
    if (slots == 0 || (slots & (slots - 1))) return PBK_E_ARG;           // DoubleHashError (doubleHash.h:48-56)
    const u64 index_size = slots - 1, index_length = calc_length(slots);
    const u64 shifter = index_length >= 32 ? 0 : 2 * index_length;       // doubleHash.h:233-235

    std::vector<u64> bitmap((slots + 63) / 64, 0), slot(n), idx(n);
    // the home slots are effectively random positions in a bitmap far larger than the caches: compute them a few
    // keys ahead and prefetch, so the placement loop does not pay one DRAM miss per key
    auto home_and_step = [&](u64 i, u64 *step) -> u64 {
        const uint64_t *key = keys + i * words;
        u64 h = 0, s = 0;                                                // makeHashKey / reHashKey
        for (unsigned j = 0; j < words; ++j) {
            h += key[j] + (key[j] >> index_length) + (key[j] >> shifter);
            s += ~key[j] ^ (key[j] >> index_length) ^ (key[j] >> shifter);
        }
        *step = s | 1;
        return h & index_size;
    };
    const u64 AHEAD = 32;
    u64 ring_v[AHEAD], ring_s[AHEAD];
    for (u64 i = 0; i < std::min<u64>(AHEAD, n); ++i) {
        ring_v[i] = home_and_step(i, &ring_s[i]);
        __builtin_prefetch(&bitmap[ring_v[i] >> 6], 1);
    }
    for (u64 i = 0; i < n; ++i) {
        u64 v = ring_v[i % AHEAD];
        const u64 step = ring_s[i % AHEAD];
        if (i + AHEAD < n) {
            ring_v[i % AHEAD] = home_and_step(i + AHEAD, &ring_s[i % AHEAD]);
            __builtin_prefetch(&bitmap[ring_v[i % AHEAD] >> 6], 1);
        }
        while (bitmap[v >> 6] >> (v & 63) & 1) v = (v + step) & index_size;   // find_any, keys are distinct
        bitmap[v >> 6] |= 1ull << (v & 63);
        slot[i] = v; idx[i] = i;
    }
    std::vector<u64>().swap(bitmap);
    sort_by_slot(slot, idx, index_size);

    FILE *fp = fopen(path, "wb");
    if (!fp) return PBK_E_IO;
    const u64 k64 = k, raw = key_raw_size(k);
    const size_t rec = 8 + raw + (k > 160 ? 8 * words : 0) + 2;
    std::vector<unsigned char> buf;
    buf.reserve(rec * 65536 + 16);
    buf.resize(16);
    memcpy(&buf[0], &k64, 8);                                            // counter.h:960
    memcpy(&buf[8], &index_size, 8);                                     // doubleHash.h:268
    bool ok = true;
    for (u64 j = 0; j < n && ok; ++j) {
        const size_t at = buf.size();
        buf.resize(at + rec, 0);
        unsigned char *p = &buf[at];
        const uint64_t *key = keys + idx[j] * words;
        memcpy(p, &slot[j], 8);                                          // doubleHash.h:272
        if (k <= 32) memcpy(p + 8, key, 8);
        else if (k <= 160) { memcpy(p + 8 + 16, &k64, 8); memcpy(p + 8 + 24, key, 8 * words); }   // {vptr, value*, len, entity}
        else { memcpy(p + 8 + 16, &k64, 8); memcpy(p + 8 + 24, key, 8 * words); }                 // {vptr, value*, len} + words (doubleHash.h:77-80)
        memcpy(p + rec - 2, &counts[idx[j]], 2);                         // doubleHash.h:275
        if (buf.size() >= rec * 65536) { ok = fwrite(buf.data(), 1, buf.size(), fp) == buf.size(); buf.clear(); }
    }
    if (ok && !buf.empty()) ok = fwrite(buf.data(), 1, buf.size(), fp) == buf.size();
    ok = (fclose(fp) == 0) && ok;
    return ok ? PBK_OK : PBK_E_IO;
}

}  // extern "C"
