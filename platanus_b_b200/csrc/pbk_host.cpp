// pbk_host.cpp -- the parts of the counting path that stay on the host:
// histogram statistics, the coverage cutoff rule, PREFIX_<k>merFrq.tsv and PREFIX_kmer_occ.bin.
// They are O(65535) or O(#kept k-mers) and serial in the reference as well; written from the
// behaviour of the cited reference lines, independently of oracle/.
#include "../../include/pbk.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <utility>
#include <vector>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

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

// ASCII -> the library's 2-bit stream words + N position list (pbk_push_reads_packed).  Same character rules as the device
// pack_kernel (Char2Bin, common.h:256, looks at the low nibble only): 1/A 3/C 7/G 4/T, 14/N, 15 -> A; anything else has no code.
int pbk_pack_reads(const uint8_t *bases, uint64_t n_bases, uint64_t *words_out, uint64_t *n_positions_out, uint64_t n_cap,
                   uint64_t *n_n_out)
{
    if ((n_bases && (!bases || !words_out)) || !n_n_out) return PBK_E_ARG;
    static const struct Lut { uint8_t code[256]; Lut() {
        for (int c = 0; c < 256; ++c) {
            const int nib = c & 15;
            code[c] = nib == 1 ? 0 : nib == 3 ? 1 : nib == 7 ? 2 : nib == 4 ? 3 : nib == 15 ? 0 : nib == 14 ? 4 : 255;
        }
    } } lut;
    u64 n_n = 0;
    bool bad = false, full = false;
    const u64 n_words = (n_bases + 31) / 32;
    for (u64 w = 0; w < n_words; ++w) {
        const u64 b0 = w * 32, m = std::min<u64>(32, n_bases - b0);
        u64 word = 0;
        for (u64 i = 0; i < m; ++i) {
            const uint8_t c = lut.code[bases[b0 + i]];
            if (c <= 3) { word |= (u64)c << (62 - 2 * i); continue; }
            if (c == 255) { bad = true; continue; }
            if (n_positions_out && n_n < n_cap) n_positions_out[n_n] = b0 + i; else full = true;
            ++n_n;
        }
        words_out[w] = word;
    }
    *n_n_out = n_n;
    if (bad) return PBK_E_BAD_BASE;
    return (full && n_positions_out) ? PBK_E_ARG : PBK_OK;
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

    std::vector<u64> bitmap((slots + 63) / 64, 0), slot(n);
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
        slot[i] = v;
    }
    std::vector<u64>().swap(bitmap);

    // Records go out in ascending slot order (writeTable scans the table, doubleHash.h:270).  The slot space is cut
    // into T ranges; every thread collects the entries of its range, sorts them, formats the records and writes
    // them at its own file offset (the reference does all of this in one serial scan of up to 2^29+ slots).
    const u64 k64 = k, raw = key_raw_size(k);
    const size_t rec = 8 + raw + (k > 160 ? 8 * words : 0) + 2;
    const int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) return PBK_E_IO;
    unsigned char header[16];
    memcpy(header, &k64, 8);                                             // counter.h:960
    memcpy(header + 8, &index_size, 8);                                  // doubleHash.h:268
    bool ok = pwrite(fd, header, 16, 0) == 16;
    unsigned T = n < 200000 ? 1u : std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
    std::vector<u64> range_count(T, 0), range_first(T + 1, 0);
    auto range_of = [&](u64 v) -> unsigned { return (unsigned)(((unsigned __int128)v * T) / slots); };
    for (u64 i = 0; i < n; ++i) range_count[range_of(slot[i])] += 1;
    for (unsigned t = 0; t < T; ++t) range_first[t + 1] = range_first[t] + range_count[t];
    std::vector<int> thread_ok(T, 1);
    auto work = [&](unsigned t) {
        std::vector<std::pair<u64, u64> > mine;
        mine.reserve(range_count[t]);
        for (u64 i = 0; i < n; ++i)
            if (range_of(slot[i]) == t) mine.push_back(std::make_pair(slot[i], i));
        std::sort(mine.begin(), mine.end());
        const size_t CH = 65536;
        std::vector<unsigned char> buf(rec * CH);
        for (size_t j0 = 0; j0 < mine.size(); j0 += CH) {
            const size_t m = std::min(CH, mine.size() - j0);
            memset(buf.data(), 0, rec * m);
            for (size_t j = 0; j < m; ++j) {
                unsigned char *p = &buf[rec * j];
                const u64 i = mine[j0 + j].second;
                const uint64_t *key = keys + i * words;
                memcpy(p, &mine[j0 + j].first, 8);                       // doubleHash.h:272
                if (k <= 32) memcpy(p + 8, key, 8);
                else { memcpy(p + 8 + 16, &k64, 8); memcpy(p + 8 + 24, key, 8 * words); }   // {vptr, value*, len, entity | words} (doubleHash.h:73-80)
                memcpy(p + rec - 2, &counts[i], 2);                      // doubleHash.h:275
            }
            const off_t at = (off_t)(16 + rec * (range_first[t] + j0));
            if (pwrite(fd, buf.data(), rec * m, at) != (ssize_t)(rec * m)) { thread_ok[t] = 0; return; }
        }
    };
    if (T == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < T; ++t) th.push_back(std::thread(work, t));
        for (unsigned t = 0; t < T; ++t) th[t].join();
    }
    for (unsigned t = 0; t < T; ++t) ok = ok && thread_ok[t];
    ok = (close(fd) == 0) && ok;
    return ok ? PBK_OK : PBK_E_IO;
}

// Counter::readOccurrenceTableBinary (counter.h:967-993) + DoubleHash::readTable / readKey (doubleHash.h:280-293,
// 83-91): the (key, count) entries of a PREFIX_kmer_occ.bin.  The reference drops every record into the slot the file
// names; here the entries are handed back as arrays (keys: n x ceil(k/32) words, word 0 first) and the caller's own
// table decides where they live (pbk_load_entries).  The arrays are malloc'ed; release them with pbk_free.
int pbk_read_kmer_occ_bin(const char *path, uint32_t *k_out, uint64_t *index_size_out, uint64_t **keys_out,
                          uint16_t **counts_out, uint64_t *n_out)
{
    if (!path || !k_out || !keys_out || !counts_out || !n_out) return PBK_E_ARG;
    *keys_out = NULL; *counts_out = NULL; *n_out = 0;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return PBK_E_IO;
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 16) { close(fd); return PBK_E_IO; }
    unsigned char header[16];
    if (pread(fd, header, 16, 0) != 16) { close(fd); return PBK_E_IO; }
    u64 k64, index_size;
    memcpy(&k64, header, 8);                                             // counter.h:972
    memcpy(&index_size, header + 8, 8);                                  // doubleHash.h:283
    if (k64 == 0 || k64 > PBK_MAX_K) { close(fd); return PBK_E_UNSUPPORTED_K; }
    const uint32_t k = (uint32_t)k64;
    const unsigned words = (k + 31) / 32;
    const u64 raw = key_raw_size(k);
    const size_t rec = 8 + raw + (k > 160 ? 8 * words : 0) + 2, key_at = k <= 32 ? 8 : 8 + 24;
    const u64 body = (u64)st.st_size - 16;
    if (body % rec != 0) { close(fd); return PBK_E_IO; }                 // truncated file
    const u64 n = body / rec;
    uint64_t *keys = (uint64_t *)malloc(std::max<size_t>(8, (size_t)n * words * 8));
    uint16_t *counts = (uint16_t *)malloc(std::max<size_t>(8, (size_t)n * 2));
    if (!keys || !counts) { free(keys); free(counts); close(fd); return PBK_E_NOMEM; }
    const size_t CH = 65536;
    std::vector<unsigned char> buf(rec * CH);
    bool ok = true;
    for (u64 i0 = 0; i0 < n && ok; i0 += CH) {
        const size_t m = (size_t)std::min<u64>(CH, n - i0);
        ok = pread(fd, buf.data(), rec * m, (off_t)(16 + rec * i0)) == (ssize_t)(rec * m);
        for (size_t j = 0; j < m && ok; ++j) {
            const unsigned char *p = &buf[rec * j];
            u64 slot;
            memcpy(&slot, p, 8);
            if (slot > index_size) ok = false;                           // not a slot of this table
            memcpy(keys + (i0 + j) * words, p + key_at, 8 * words);
            memcpy(counts + (i0 + j), p + rec - 2, 2);
        }
    }
    close(fd);
    if (!ok) { free(keys); free(counts); return PBK_E_IO; }
    *k_out = k;
    if (index_size_out) *index_size_out = index_size;
    *keys_out = keys; *counts_out = counts; *n_out = n;
    return PBK_OK;
}

void pbk_free(void *p) { free(p); }

}  // extern "C"
