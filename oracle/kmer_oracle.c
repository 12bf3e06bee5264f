/*
 * kmer_oracle.c -- TEST INFRASTRUCTURE ONLY (see kmer_oracle.h).
 *
 * Plain-C restatement of Platanus_B's k-mer occurrence counting path, written from the
 * behaviour of the reference (citations are file:line under /root/reference).  It keeps the
 * reference's *sequential* formulation on purpose (rolling forward/reverse words, N list,
 * saturating count) so that it is an independent check of the CUDA path, which is organised
 * completely differently.  Pinned against the reference binary by tests/test_oracle_golden.py.
 */
#define _GNU_SOURCE
#include "kmer_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/types.h>

/* ------------------------------------------------------------------------------------------ */
/* reads container                                                                            */
/* ------------------------------------------------------------------------------------------ */

pbo_reads *pbo_reads_new(void)
{
    pbo_reads *r = (pbo_reads *)calloc(1, sizeof(*r));
    if (!r) return NULL;
    r->cap_reads = 1024;
    r->cap_bases = 1 << 16;
    r->offsets = (uint64_t *)malloc((r->cap_reads + 1) * sizeof(uint64_t));
    r->bases = (char *)malloc(r->cap_bases);
    if (!r->offsets || !r->bases) { pbo_reads_free(r); return NULL; }
    r->offsets[0] = 0;
    return r;
}

void pbo_reads_free(pbo_reads *r)
{
    if (!r) return;
    free(r->bases);
    free(r->offsets);
    free(r);
}

int pbo_reads_add(pbo_reads *r, const char *seq, uint64_t len)
{
    uint64_t end = r->offsets[r->n_reads] + len;
    if (end > r->cap_bases) {
        uint64_t cap = r->cap_bases;
        while (cap < end) cap *= 2;
        char *p = (char *)realloc(r->bases, cap);
        if (!p) return PBO_E_NOMEM;
        r->bases = p; r->cap_bases = cap;
    }
    if (r->n_reads + 1 > r->cap_reads) {
        uint64_t cap = r->cap_reads * 2;
        uint64_t *p = (uint64_t *)realloc(r->offsets, (cap + 1) * sizeof(uint64_t));
        if (!p) return PBO_E_NOMEM;
        r->offsets = p; r->cap_reads = cap;
    }
    if (len) memcpy(r->bases + r->offsets[r->n_reads], seq, len);
    r->n_reads += 1;
    r->offsets[r->n_reads] = end;
    return PBO_OK;
}

/* platanus::getlineFILE (common.cpp:160-175): POSIX getline, strip one trailing '\n'; the result
 * is then assigned to a std::string through a C string, i.e. it is cut at the first NUL. */
static ssize_t oracle_getline(char **buf, size_t *cap, FILE *fp)
{
    ssize_t n = getline(buf, cap, fp);
    if (n <= 0) return -1;
    if ((*buf)[n - 1] == '\n') { (*buf)[n - 1] = '\0'; n -= 1; }
    return (ssize_t)strlen(*buf) <= n ? (ssize_t)strlen(*buf) : n;
}

/* BaseCommand::checkFileFormat (baseCommand.cpp:29-50) */
int pbo_check_file_format(const char *path)
{
    FILE *fp = fopen(path, "r");
    if (!fp) return PBO_E_IO;
    char *line[4] = {NULL, NULL, NULL, NULL};
    size_t cap[4] = {0, 0, 0, 0};
    ssize_t len[4];
    for (int i = 0; i < 4; ++i) {
        len[i] = oracle_getline(&line[i], &cap[i], fp);
        if (len[i] < 0) {            /* the reference leaves the std::string empty */
            free(line[i]); line[i] = (char *)calloc(1, 1); len[i] = 0;
        }
    }
    fclose(fp);
    int type = 0;
    int acgtn = (strspn(line[1], "ACGTN") == (size_t)len[1]);
    if (line[0][0] == '>' && acgtn) type = 1;
    else if (line[0][0] == '@' && acgtn && line[2][0] == '+') type = 2;
    for (int i = 0; i < 4; ++i) free(line[i]);
    return type;
}

typedef struct { char *p; size_t len, cap; } strbuf;
static int sb_append(strbuf *s, const char *t, size_t n)
{
    if (s->len + n + 1 > s->cap) {
        size_t cap = s->cap ? s->cap : 256;
        while (cap < s->len + n + 1) cap *= 2;
        char *p = (char *)realloc(s->p, cap);
        if (!p) return PBO_E_NOMEM;
        s->p = p; s->cap = cap;
    }
    memcpy(s->p + s->len, t, n);
    s->len += n;
    return PBO_OK;
}

/* std::getline(ifstream&, string&) keeps embedded NULs and strips only '\n'. */
static ssize_t stream_getline(char **buf, size_t *cap, FILE *fp)
{
    ssize_t n = getline(buf, cap, fp);
    if (n < 0) return -1;
    if (n > 0 && (*buf)[n - 1] == '\n') n -= 1;
    return n;
}

static int flush_read(pbo_reads *r, strbuf *read)
{
    /* SEQ::convertFromString throws ReadError at length >= MAX_READ_LEN (common.h:465) */
    if (read->len >= PBO_MAX_READ_LEN) return PBO_E_READ_TOO_LONG;
    int rc = pbo_reads_add(r, read->p, read->len);
    read->len = 0;
    return rc;
}

/* Assemble::readFastaUncompressed (assemble.cpp:816-848) */
static int read_fasta(pbo_reads *r, FILE *fp)
{
    char *line = NULL; size_t cap = 0; ssize_t n;
    strbuf read = {NULL, 0, 0};
    int rc = PBO_OK;
    while ((n = stream_getline(&line, &cap, fp)) >= 0)
        if (n > 0 && line[0] == '>') break;
    while ((n = stream_getline(&line, &cap, fp)) >= 0) {
        if (!(n > 0 && line[0] == '>')) {
            if ((rc = sb_append(&read, line, (size_t)n))) goto out;
        } else if (read.len != 0) {
            if ((rc = flush_read(r, &read))) goto out;
        }
    }
    rc = flush_read(r, &read);      /* unconditional final record, even if empty (:844-845) */
out:
    free(line); free(read.p);
    return rc;
}

/* Assemble::readFastqUncompressed (assemble.cpp:902-942) */
static int read_fastq(pbo_reads *r, FILE *fp)
{
    char *line = NULL; size_t cap = 0; ssize_t n;
    strbuf read = {NULL, 0, 0};
    int rc = PBO_OK, flag = 1;
    while ((n = stream_getline(&line, &cap, fp)) >= 0)
        if (n > 0 && line[0] == '@') break;
    while ((n = stream_getline(&line, &cap, fp)) >= 0) {
        if (n == 0) continue;
        if (line[0] != '@') {
            if (flag && line[0] != '+') {
                if ((rc = sb_append(&read, line, (size_t)n))) goto out;
            } else {
                flag = 0;
            }
        } else {
            if (read.len != 0)
                if ((rc = flush_read(r, &read))) goto out;
            flag = 1;
        }
    }
    rc = flush_read(r, &read);      /* :937-938 */
out:
    free(line); free(read.p);
    return rc;
}

int pbo_reads_add_file(pbo_reads *r, const char *path)
{
    int type = pbo_check_file_format(path);
    if (type < 0) return type;
    if (type == 0) return PBO_E_FORMAT;
    FILE *fp = fopen(path, "r");
    if (!fp) return PBO_E_IO;
    int rc = (type == 1) ? read_fasta(r, fp) : read_fastq(r, fp);
    fclose(fp);
    return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* base codes and multi-word keys                                                             */
/* ------------------------------------------------------------------------------------------ */

/* platanus::Char2Bin (common.h:256): index by the low nibble into a 16-byte literal */
unsigned char pbo_char2bin(char c)
{
    static const unsigned char table[17] = ".\x0.\x1\x3..\x2......\x4";
    return table[c & 0xF];
}

#define MAXW 8192                     /* k < 500000 would need 15625; tests stay far below */

/* BinstrBase::set (binstr.h:321-324) / Kmer31::setForward (kmer.h:129-132): clears the 2-bit
 * field and ORs the full 8-bit value shifted into place. */
static void key_set(uint64_t *v, unsigned pos, unsigned char val)
{
    unsigned w = pos / 32, sh = (pos % 32) * 2;
    v[w] = (v[w] & ~(0x3ull << sh)) | ((uint64_t)val << sh);
}

/* forward <<= 2 with the top-word mask (binstr.h:415-436); for one word, `<<= 2; &= mask`
 * (counter.h:418-419 with mask from counter.h:397). */
static void key_shl2(uint64_t *v, unsigned words, unsigned k)
{
    for (unsigned i = words - 1; i > 0; --i) v[i] = (v[i] << 2) | (v[i - 1] >> 62);
    v[0] <<= 2;
    if (words == 1) {
        uint64_t mask = k >= 32 ? ~0ull : ~(~0ull << (2 * k));
        v[0] &= mask;
    } else if (k % 32 > 0) {
        v[words - 1] &= ((0x1ull << (2 * (k % 32))) - 0x1ull);
    }
}

/* reverse >>= 2 (binstr.h:375-405): zero fill from the top, no mask */
static void key_shr2(uint64_t *v, unsigned words)
{
    for (unsigned i = 0; i + 1 < words; ++i) v[i] = (v[i] >> 2) | (v[i + 1] << 62);
    v[words - 1] >>= 2;
}

/* binstr.h:460-466 (top word first); plain `<` on u64 for one word */
int pbo_key_cmp(const uint64_t *a, const uint64_t *b, unsigned words)
{
    for (unsigned i = words; i-- > 0;) {
        if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
    }
    return 0;
}

static unsigned g_sort_words;
static int sort_cmp(const void *a, const void *b)
{
    return pbo_key_cmp((const uint64_t *)a, (const uint64_t *)b, g_sort_words);
}

/* ------------------------------------------------------------------------------------------ */
/* counting                                                                                   */
/* ------------------------------------------------------------------------------------------ */

void pbo_result_release(pbo_result *res)
{
    if (!res) return;
    free(res->keys); free(res->counts); free(res->len_hist);
    res->keys = NULL; res->counts = NULL; res->len_hist = NULL;
}

static uint16_t table_find(const uint64_t *keys, const uint16_t *counts, uint64_t n, unsigned words, const uint64_t *key);

/* seeds == NULL: Counter::makeKmerReadDistributionMT.  With seeds (sorted (key, value) dump of the table the counter
 * holds when it is called): Counter::makeKmerReadDistributionConsideringPreviousGraph (counter.h:663-750) -- a window
 * whose k-mer has a non-zero value in the table is not counted (divideKmerUsedMakingPreviousContig, counter.h:828-861:
 * `find_any(key)->second == 0` is the test for "not in the contigs"), the table's own entries are written with their
 * values (counter.h:695-705), everything else is counted as in the first pass (countKmerPerThreadSecond). */
static int count_impl(const pbo_reads *r, unsigned k, const uint64_t *seed_keys, const uint16_t *seed_counts,
                      uint64_t n_seed, pbo_result *res);

int pbo_count(const pbo_reads *r, unsigned k, pbo_result *res)
{
    return count_impl(r, k, NULL, NULL, 0, res);
}

int pbo_count_seeded(const pbo_reads *r, unsigned k, const uint64_t *seed_keys, const uint16_t *seed_counts,
                     uint64_t n_seed, pbo_result *res)
{
    return count_impl(r, k, seed_keys, seed_counts, n_seed, res);
}

static int count_impl(const pbo_reads *r, unsigned k, const uint64_t *seed_keys, const uint16_t *seed_counts,
                      uint64_t n_seed, pbo_result *res)
{
    if (k == 0) return PBO_E_ARG;
    unsigned words = (k + 31) / 32;
    if (words > MAXW) return PBO_E_ARG;
    memset(res, 0, sizeof(*res));
    res->k = k; res->words = words;
    res->len_hist = (uint64_t *)calloc(PBO_MAX_READ_LEN + 1, sizeof(uint64_t));
    if (!res->len_hist) return PBO_E_NOMEM;

    /* upper bound on instances */
    uint64_t max_inst = 0;
    for (uint64_t i = 0; i < r->n_reads; ++i) {
        uint64_t len = r->offsets[i + 1] - r->offsets[i];
        if (len >= k) max_inst += len - k + 1;
    }
    uint64_t *inst = (uint64_t *)malloc((max_inst ? max_inst : 1) * words * sizeof(uint64_t));
    if (!inst) return PBO_E_NOMEM;
    uint64_t n_inst = 0;

    uint64_t *fwd = (uint64_t *)calloc(words, sizeof(uint64_t));
    uint64_t *rev = (uint64_t *)calloc(words, sizeof(uint64_t));
    /* SEQ::base of the single parser-side SEQ object (assemble.cpp:819/905): N positions keep the
     * byte of the previous read (common.h:468-476); std::string::resize zero-fills growth. */
    unsigned char *base = (unsigned char *)calloc(PBO_MAX_READ_LEN + 1, 1);
    uint64_t base_len = 0;
    int64_t *npos = (int64_t *)malloc((PBO_MAX_READ_LEN + 2) * sizeof(int64_t));
    if (!fwd || !rev || !base || !npos) { free(inst); free(fwd); free(rev); free(base); free(npos); return PBO_E_NOMEM; }

    for (uint64_t ri = 0; ri < r->n_reads; ++ri) {
        const char *s = r->bases + r->offsets[ri];
        uint64_t len = r->offsets[ri + 1] - r->offsets[ri];
        /* convertFromString (common.h:460-477) */
        if (len < base_len) base_len = len;                    /* resize shrink */
        if (len > base_len) { memset(base + base_len, 0, len - base_len); base_len = len; }
        uint64_t nn = 0;
        for (uint64_t i = 0; i < len; ++i) {
            unsigned char c = pbo_char2bin(s[i]);
            if (c == 4) npos[nn++] = (int64_t)i; else base[i] = c;
        }
        /* countKmerPerThreadFirst (counter.h:405-432) */
        res->len_hist[len] += 1;
        if (len < k) continue;
        npos[nn] = PBO_MAX_READ_LEN + 1;                       /* sentinel, counter.h:411 */
        uint64_t cur = 0;
        /* NB: the reference constructs KMER once per thread and never clears it between reads
         * (counter.h:401); the k-1 priming sets plus k-1 shifts overwrite every position that can
         * reach a counted window, except garbage above bit 2k of `reverse` for k<32 left by codes
         * > 3.  We keep one state across reads for the same reason. */
        for (unsigned i = 0; i + 1 < k; ++i) {
            key_set(fwd, k - i - 2, base[i]);
            key_set(rev, i + 1, (unsigned char)(0x3 ^ base[i]));
        }
        for (uint64_t i = 0; i < len - k + 1; ++i) {
            key_shl2(fwd, words, k);
            key_set(fwd, 0, base[i + k - 1]);
            key_shr2(rev, words);
            key_set(rev, k - 1, (unsigned char)(0x3 ^ base[i + k - 1]));
            if ((uint64_t)npos[cur] < i + k) {
                if ((uint64_t)npos[cur] <= i) ++cur;
                continue;
            }
            const uint64_t *key = pbo_key_cmp(fwd, rev, words) <= 0 ? fwd : rev;   /* std::min */
            if (seed_keys && table_find(seed_keys, seed_counts, n_seed, words, key) != 0) continue;   /* counter.h:853 */
            memcpy(inst + n_inst * words, key, words * sizeof(uint64_t));
            ++n_inst;
        }
    }
    free(fwd); free(rev); free(base); free(npos);

    g_sort_words = words;
    qsort(inst, n_inst, words * sizeof(uint64_t), sort_cmp);

    /* run-length encode in place; saturate at 65534 (counter.h:468) */
    uint64_t nd = 0;
    uint16_t *counts = (uint16_t *)malloc((n_inst ? n_inst : 1) * sizeof(uint16_t));
    if (!counts) { free(inst); return PBO_E_NOMEM; }
    for (uint64_t i = 0; i < n_inst;) {
        uint64_t j = i + 1;
        while (j < n_inst && pbo_key_cmp(inst + i * words, inst + j * words, words) == 0) ++j;
        uint64_t c = j - i;
        if (c > PBO_COUNT_SAT) c = PBO_COUNT_SAT;
        memmove(inst + nd * words, inst + i * words, words * sizeof(uint64_t));
        counts[nd] = (uint16_t)c;
        res->occ_hist[c] += 1;                                  /* counter.h:496 */
        ++nd;
        i = j;
    }
    /* the table's own entries, with their values (counter.h:695-705); none of them was counted above */
    uint64_t n_add = 0;
    for (uint64_t i = 0; i < n_seed; ++i) n_add += seed_counts[i] != 0;
    if (n_add) {
        uint64_t tot = nd + n_add;
        uint64_t *mk = (uint64_t *)malloc(tot * (words + 1) * sizeof(uint64_t));      /* key words + count, sorted together */
        if (!mk) { free(inst); free(counts); return PBO_E_NOMEM; }
        uint64_t m = 0;
        for (uint64_t i = 0; i < nd; ++i, ++m) { memcpy(mk + m * (words + 1), inst + i * words, words * 8); mk[m * (words + 1) + words] = counts[i]; }
        for (uint64_t i = 0; i < n_seed; ++i) {
            if (seed_counts[i] == 0) continue;
            memcpy(mk + m * (words + 1), seed_keys + i * words, words * 8);
            mk[m * (words + 1) + words] = seed_counts[i];
            res->occ_hist[seed_counts[i] < PBO_OCC_BINS ? seed_counts[i] : PBO_OCC_BINS - 1] += 1;   /* counter.h:701 */
            ++m;
        }
        /* sort records by key: pbo_key_cmp looks at the first `words` words only, the count rides along */
        g_sort_words = words;
        qsort(mk, tot, (words + 1) * sizeof(uint64_t), sort_cmp);
        free(inst); free(counts);
        inst = (uint64_t *)malloc(tot * words * sizeof(uint64_t));
        counts = (uint16_t *)malloc(tot * sizeof(uint16_t));
        if (!inst || !counts) { free(mk); free(inst); free(counts); return PBO_E_NOMEM; }
        for (uint64_t i = 0; i < tot; ++i) { memcpy(inst + i * words, mk + i * (words + 1), words * 8); counts[i] = (uint16_t)mk[i * (words + 1) + words]; }
        free(mk);
        nd = tot;
    }
    res->keys = inst; res->counts = counts;
    res->n_distinct = nd; res->n_instances = n_inst;
    res->max_occ = 0;
    for (unsigned i = PBO_OCC_BINS - 1; i > 0; --i)             /* counter.h:371-376 */
        if (res->occ_hist[i] > 0) { res->max_occ = i; break; }
    return PBO_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* occurrence lookup                                                                          */
/* ------------------------------------------------------------------------------------------ */

/* Counter::findValue -> DoubleHash::findValue: the stored count, 0 when the key is absent.  The
 * table here is the sorted (key, count) dump, searched by bisection. */
static uint16_t table_find(const uint64_t *keys, const uint16_t *counts, uint64_t n, unsigned words, const uint64_t *key)
{
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        uint64_t mid = lo + (hi - lo) / 2;
        int c = pbo_key_cmp(keys + mid * words, key, words);
        if (c == 0) return counts[mid];
        if (c < 0) lo = mid + 1; else hi = mid;
    }
    return 0;
}

/* Counter::makeKmerReadDistributionFromContig (counter.h:511-593): occurrenceTable[key] = max(itself,
 * max(coverage[contig], minOccurrence)) over every window of every contig of length >= k, then
 * writeKmerDistribution (entries with a non-zero value, counter.h:483-507).  Sequences are loaded as Contig::setSeq does
 * (common.h:528-537: base = Char2Bin(c)); like the reference, windows with an N are NOT skipped (the skip is commented
 * out, counter.h:559-566) -- the code 4 goes into the 2-bit fields as it is. */
int pbo_count_contigs(const pbo_reads *r, unsigned k, const uint16_t *coverage, uint64_t min_occ, pbo_result *res)
{
    if (k == 0) return PBO_E_ARG;
    unsigned words = (k + 31) / 32;
    if (words > MAXW) return PBO_E_ARG;
    memset(res, 0, sizeof(*res));
    res->k = k; res->words = words;
    res->len_hist = (uint64_t *)calloc(PBO_MAX_READ_LEN + 1, sizeof(uint64_t));
    if (!res->len_hist) return PBO_E_NOMEM;
    uint64_t max_inst = 0;
    for (uint64_t i = 0; i < r->n_reads; ++i) {
        uint64_t len = r->offsets[i + 1] - r->offsets[i];
        if (len >= k) max_inst += len - k + 1;
    }
    const unsigned rw = words + 1;                                    /* key words + value, sorted together */
    uint64_t *rec = (uint64_t *)malloc((max_inst ? max_inst : 1) * rw * sizeof(uint64_t));
    uint64_t *fwd = (uint64_t *)calloc(words, sizeof(uint64_t)), *rev = (uint64_t *)calloc(words, sizeof(uint64_t));
    if (!rec || !fwd || !rev) { free(rec); free(fwd); free(rev); return PBO_E_NOMEM; }
    uint64_t n = 0;
    for (uint64_t ri = 0; ri < r->n_reads; ++ri) {
        const char *s = r->bases + r->offsets[ri];
        uint64_t len = r->offsets[ri + 1] - r->offsets[ri];
        if (len < k) continue;                                        /* counter.h:543-544 */
        uint64_t v = coverage[ri] > min_occ ? coverage[ri] : min_occ;  /* counter.h:573 */
        v &= 0xFFFF;                                                  /* stored in an unsigned short */
        for (unsigned i = 0; i + 1 < k; ++i) {
            unsigned char b = pbo_char2bin(s[i]);
            key_set(fwd, k - i - 2, b);
            key_set(rev, i + 1, (unsigned char)(0x3 ^ b));
        }
        for (uint64_t i = 0; i < len - k + 1; ++i) {
            unsigned char b = pbo_char2bin(s[i + k - 1]);
            key_shl2(fwd, words, k);
            key_set(fwd, 0, b);
            key_shr2(rev, words);
            key_set(rev, k - 1, (unsigned char)(0x3 ^ b));
            const uint64_t *key = pbo_key_cmp(fwd, rev, words) <= 0 ? fwd : rev;
            memcpy(rec + n * rw, key, words * 8);
            rec[n * rw + words] = v;
            ++n;
        }
    }
    free(fwd); free(rev);
    g_sort_words = words;
    qsort(rec, n, rw * sizeof(uint64_t), sort_cmp);
    uint64_t *keys = (uint64_t *)malloc((n ? n : 1) * words * sizeof(uint64_t));
    uint16_t *counts = (uint16_t *)malloc((n ? n : 1) * sizeof(uint16_t));
    if (!keys || !counts) { free(rec); free(keys); free(counts); return PBO_E_NOMEM; }
    uint64_t nd = 0;
    for (uint64_t i = 0; i < n;) {
        uint64_t j = i, best = 0;
        while (j < n && pbo_key_cmp(rec + i * rw, rec + j * rw, words) == 0) { if (rec[j * rw + words] > best) best = rec[j * rw + words]; ++j; }
        if (best != 0) {                                              /* counter.h:490 */
            memcpy(keys + nd * words, rec + i * rw, words * 8);
            counts[nd] = (uint16_t)best;
            res->occ_hist[best < PBO_OCC_BINS ? best : PBO_OCC_BINS - 1] += 1;
            ++nd;
        }
        i = j;
    }
    free(rec);
    res->keys = keys; res->counts = counts; res->n_distinct = nd; res->n_instances = 0;
    res->max_occ = 0;
    for (unsigned i = PBO_OCC_BINS - 1; i > 0; --i)
        if (res->occ_hist[i] > 0) { res->max_occ = i; break; }
    return PBO_OK;
}

/* Counter::pickupReadMatchedEdgeKmer (counter.h:870-910): out[r] = 1 iff read r is kept -- it is at least k long and
 * one of its windows without an N has a k-mer with a non-zero value in the table.  Reads go through
 * SEQ::convertFromString like in pbo_count (N positions as a list, counter.h:882-884, 895-899). */
int pbo_match_reads(const pbo_reads *r, unsigned k, const uint64_t *keys, const uint16_t *counts, uint64_t n, uint8_t *out)
{
    if (k == 0) return PBO_E_ARG;
    unsigned words = (k + 31) / 32;
    if (words > MAXW) return PBO_E_ARG;
    uint64_t *fwd = (uint64_t *)calloc(words, sizeof(uint64_t));
    uint64_t *rev = (uint64_t *)calloc(words, sizeof(uint64_t));
    if (!fwd || !rev) { free(fwd); free(rev); return PBO_E_NOMEM; }
    for (uint64_t ri = 0; ri < r->n_reads; ++ri) {
        const char *s = r->bases + r->offsets[ri];
        uint64_t len = r->offsets[ri + 1] - r->offsets[ri];
        out[ri] = 0;
        if (len < k) continue;                                        /* counter.h:880 */
        for (unsigned i = 0; i + 1 < k; ++i) {
            unsigned char b = pbo_char2bin(s[i]);
            if (b == 4) b = 0;                                        /* the byte under an N is never part of a used window */
            key_set(fwd, k - i - 2, b);
            key_set(rev, i + 1, (unsigned char)(0x3 ^ b));
        }
        int64_t last_n = -1;
        for (unsigned i = 0; i + 1 < k; ++i) if (pbo_char2bin(s[i]) == 4) last_n = (int64_t)i;
        for (uint64_t i = 0; i < len - k + 1; ++i) {
            unsigned char b = pbo_char2bin(s[i + k - 1]);
            if (b == 4) { last_n = (int64_t)(i + k - 1); b = 0; }
            key_shl2(fwd, words, k);
            key_set(fwd, 0, b);
            key_shr2(rev, words);
            key_set(rev, k - 1, (unsigned char)(0x3 ^ b));
            if (last_n >= (int64_t)i) continue;                       /* the window [i, i + k) holds an N (counter.h:895-899) */
            const uint64_t *key = pbo_key_cmp(fwd, rev, words) <= 0 ? fwd : rev;
            if (table_find(keys, counts, n, words, key) > 0) { out[ri] = 1; break; }   /* counter.h:900-903 */
        }
    }
    free(fwd); free(rev);
    return PBO_OK;
}

/* ContigDivider::getOccurrenceArray (kmer_divide.cpp:151-197) over sequences loaded the way
 * Contig::setSeq does (common.h:528-537: base[i] = Char2Bin(c), an N stays 4).  out has one entry per
 * BASE, out[offsets[r] + start] for the window starting at `start` of sequence r (the reference's
 * OccurrenceArray has length - k + 1 entries per sequence, zero-initialised, kmer_divide.h:45-50);
 * entries no window starts at stay 0.  Pinned by tests/golden/occ_*.npz (oracle/ref_occ_harness.cpp). */
int pbo_occurrence_array(const pbo_reads *r, unsigned k, const uint64_t *keys, const uint16_t *counts,
                         uint64_t n, uint16_t *out)
{
    if (k == 0) return PBO_E_ARG;
    unsigned words = (k + 31) / 32;
    if (words > MAXW) return PBO_E_ARG;
    uint64_t *fwd = (uint64_t *)calloc(words, sizeof(uint64_t));
    uint64_t *rev = (uint64_t *)calloc(words, sizeof(uint64_t));
    if (!fwd || !rev) { free(fwd); free(rev); return PBO_E_NOMEM; }
    memset(out, 0, r->offsets[r->n_reads] * sizeof(uint16_t));
    for (uint64_t ri = 0; ri < r->n_reads; ++ri) {
        const char *s = r->bases + r->offsets[ri];
        uint64_t len = r->offsets[ri + 1] - r->offsets[ri];
        if (len < k) continue;                                   /* numKmer == 0 (kmer_divide.cpp:160) */
        uint64_t num_kmer = len - k + 1, start = 0;
        uint16_t *o = out + r->offsets[ri];
        int is_init = 1;
        while (start < num_kmer) {
            if (is_init) {
                uint64_t j = 0;
                for (; j < k - 1; ++j) {
                    unsigned char b = pbo_char2bin(s[start + j]);
                    if (b == 4) break;
                    key_set(fwd, k - 2 - (unsigned)j, b);
                    key_set(rev, (unsigned)j + 1, (unsigned char)(0x3 ^ b));
                }
                if (j == k - 1) is_init = 0;
                else { start += j + 1; continue; }
            }
            unsigned char b = pbo_char2bin(s[start + k - 1]);
            if (b == 4) { start += k; is_init = 1; continue; }
            key_shl2(fwd, words, k);
            key_shr2(rev, words);
            key_set(fwd, 0, b);
            key_set(rev, k - 1, (unsigned char)(0x3 ^ b));
            const uint64_t *key = pbo_key_cmp(fwd, rev, words) <= 0 ? fwd : rev;
            o[start] = table_find(keys, counts, n, words, key);
            ++start;
        }
    }
    free(fwd); free(rev);
    return PBO_OK;
}

/* The eight neighbour probes BruijnGraph::makeInitialBruijnGraph makes per k-mer (graph.h:337-375): for the k-mer as it
 * stands in sortedKeyFP (its forward orientation = the canonical key), bit b of leftFlags says that the k-mer
 * "b + first k-1 bases" is in the table, bit b of rightFlags that "last k-1 bases + b" is (canonical form looked up,
 * findValue != 0); out[i] = (leftFlags << 4) | rightFlags, the layout of Junction::out (graph.h:398).  The table is the one
 * loadKmer builds: only keys with count >= min_count count as present.  Pinned by tests/golden/flags_k*.npz
 * (oracle/ref_iter_harness.cpp, mode flags: the reference's own KMER primitives and Counter::findValue). */
int pbo_neighbor_flags(unsigned k, const uint64_t *keys, const uint16_t *counts, uint64_t n, uint32_t min_count, uint8_t *out)
{
    if (k == 0) return PBO_E_ARG;
    unsigned words = (k + 31) / 32;
    if (words > MAXW) return PBO_E_ARG;
    uint64_t *buf = (uint64_t *)calloc(4 * (size_t)words, sizeof(uint64_t));
    if (!buf) return PBO_E_NOMEM;
    uint64_t *fwd = buf, *rev = buf + words, *nf = buf + 2 * words, *nr = buf + 3 * words;
    for (uint64_t i = 0; i < n; ++i) {
        out[i] = 0;
        if (counts[i] < min_count) continue;
        const uint64_t *key = keys + i * words;
        memcpy(fwd, key, words * sizeof(uint64_t));
        memset(rev, 0, words * sizeof(uint64_t));                    /* KmerBase::reverseComplement */
        for (unsigned j = 0; j < k; ++j)
            key_set(rev, k - 1 - j, (unsigned char)(0x3 ^ ((fwd[j / 32] >> (2 * (j % 32))) & 3)));
        unsigned left = 0, right = 0;
        for (unsigned char b = 0; b < 4; ++b) {
            /* leftKmer: forward >>= 2, reverse <<= 2 (masked), base b at the top of forward / its complement at the bottom of reverse */
            memcpy(nf, fwd, words * sizeof(uint64_t)); memcpy(nr, rev, words * sizeof(uint64_t));
            key_shr2(nf, words); key_shl2(nr, words, k);
            key_set(nf, k - 1, b); key_set(nr, 0, (unsigned char)(0x3 ^ b));
            const uint64_t *c = pbo_key_cmp(nf, nr, words) <= 0 ? nf : nr;
            if (table_find(keys, counts, n, words, c) >= (min_count ? min_count : 1)) left |= 1u << b;
            /* rightKmer: forward <<= 2 (masked), reverse >>= 2, base b at the bottom of forward / its complement at the top of reverse */
            memcpy(nf, fwd, words * sizeof(uint64_t)); memcpy(nr, rev, words * sizeof(uint64_t));
            key_shl2(nf, words, k); key_shr2(nr, words);
            key_set(nf, 0, b); key_set(nr, k - 1, (unsigned char)(0x3 ^ b));
            c = pbo_key_cmp(nf, nr, words) <= 0 ? nf : nr;
            if (table_find(keys, counts, n, words, c) >= (min_count ? min_count : 1)) right |= 1u << b;
        }
        out[i] = (uint8_t)((left << 4) | right);
    }
    free(buf);
    return PBO_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* histogram statistics                                                                       */
/* ------------------------------------------------------------------------------------------ */

/* Counter::getLeftLocalMinimalValue (counter.h:245-267) */
uint64_t pbo_left_local_min(const uint64_t *occ, uint64_t max_occ, uint64_t w)
{
    if (max_occ <= w) return 0;
    uint64_t n = max_occ - w + 2;
    uint64_t *win = (uint64_t *)calloc(n, sizeof(uint64_t));
    uint64_t i;
    for (i = 0; i < w; ++i) win[1] += occ[1 + i];
    for (i = 2; i < n; ++i) {
        win[i] = win[i - 1] - occ[i - 1] + occ[i + w - 1];
        if (win[i] >= win[i - 1]) break;
    }
    free(win);
    return (i <= max_occ) ? (i - 1 + w / 2) : (1 + w / 2);
}

/* Counter::calcDistributionAverage (counter.h:221-238) */
int pbo_dist_average(const uint64_t *dist, uint64_t size, uint64_t start, uint64_t end, double *out)
{
    if (end > size || start > end) return PBO_E_KMER_DIST;
    uint64_t sum = 0, num = 0;
    for (uint64_t i = start; i <= end; ++i) { sum += i * dist[i]; num += dist[i]; }
    if (num == 0) return PBO_E_KMER_DIST;
    *out = (double)sum / (double)num;
    return PBO_OK;
}

/* assemble.cpp:318-321, SMOOTHING_WINDOW = 1 (assemble.cpp:42) */
uint64_t pbo_coverage_cutoff(const uint64_t *occ, uint64_t max_occ, int n_opt, int repeat)
{
    if (n_opt != 0) return (uint64_t)(int64_t)n_opt;
    uint64_t v = pbo_left_local_min(occ, max_occ, 1);
    if (!repeat) v /= 2;
    return v > 2 ? v : 2;
}

/* ------------------------------------------------------------------------------------------ */
/* sizes                                                                                      */
/* ------------------------------------------------------------------------------------------ */

/* sizeof(KEY): unsigned long long (kmer.h:30); Binstr63/95/127/159 = vptr + value* + len +
 * entity[2..5] (binstr.h:292-296, 484-486 ...); binstr_t = vptr + value* + len = 24 bytes
 * (binstr.h:36-38, 51); confirmed by the record sizes of the golden .bin files. */
uint64_t pbo_key_raw_size(unsigned k)
{
    if (k <= 32) return 8;
    if (k <= 64) return 24 + 16;
    if (k <= 96) return 24 + 24;
    if (k <= 128) return 24 + 32;
    if (k <= 160) return 24 + 40;
    return 24;
}

uint64_t pbo_pair_size(unsigned k)
{
    /* std::pair<KEY, unsigned short>, 8-byte aligned */
    return pbo_key_raw_size(k) + 8;
}

/* counter.h:300-309 */
uint64_t pbo_double_hash_size(uint64_t memory, unsigned k)
{
    long base = (long)pbo_pair_size(k);
    unsigned long long tmp = memory / (unsigned long long)base;
    base = (long)(log((double)tmp) / log(2));
    tmp = (unsigned long long)pow(2, (double)base);
    while (tmp > memory) tmp >>= 1;
    return tmp;
}

/* counter.h:621-622 */
uint64_t pbo_load_size(uint64_t total)
{
    unsigned long long size = (unsigned long long)(log((double)total / 0.9) / log(2));
    size = (unsigned long long)pow(2, (double)(size + 1));
    return size;
}

/* ------------------------------------------------------------------------------------------ */
/* DoubleHash emulation and kmer_occ.bin                                                      */
/* ------------------------------------------------------------------------------------------ */

/* DoubleHash::calcLength (doubleHash.h:107-115) */
static uint64_t dh_calc_length(uint64_t len)
{
    for (uint64_t i = 1; i < 64; ++i) if ((len >> i) == 0) return i;
    return 64;
}

/* makeHashKey / reHashKey (doubleHash.h:118-146) */
void pbo_probe_start(const uint64_t *key, unsigned words, uint64_t slots, uint64_t *home, uint64_t *step)
{
    uint64_t index_size = slots - 1;
    uint64_t index_length = dh_calc_length(slots);
    uint64_t shifter = index_length >= 32 ? 0 : 2 * index_length;
    uint64_t h = 0, s = 0;
    for (unsigned i = 0; i < words; ++i) {
        h += key[i] + (key[i] >> index_length) + (key[i] >> shifter);
        s += ~key[i] ^ (key[i] >> index_length) ^ (key[i] >> shifter);
    }
    *home = h & index_size;
    *step = s | 1;
}

typedef struct { uint64_t slot; uint64_t idx; } slot_rec;
static int slot_cmp(const void *a, const void *b)
{
    uint64_t x = ((const slot_rec *)a)->slot, y = ((const slot_rec *)b)->slot;
    return x < y ? -1 : (x > y);
}

/* a small open-addressing map slot -> entry index so that the emulation does not need the
 * max(size, doubleHashSize) zero-filled table the reference allocates (counter.h:627) */
typedef struct { uint64_t *slot; uint64_t *idx; uint64_t cap; } occ_map;
static uint64_t mix64(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}
static int64_t occ_find(const occ_map *m, uint64_t slot)
{
    uint64_t p = mix64(slot) & (m->cap - 1);
    while (m->idx[p] != UINT64_MAX) {
        if (m->slot[p] == slot) return (int64_t)m->idx[p];
        p = (p + 1) & (m->cap - 1);
    }
    return -1;
}
static void occ_put(occ_map *m, uint64_t slot, uint64_t idx)
{
    uint64_t p = mix64(slot) & (m->cap - 1);
    while (m->idx[p] != UINT64_MAX) p = (p + 1) & (m->cap - 1);
    m->slot[p] = slot; m->idx[p] = idx;
}

int pbo_write_bin(const char *path, unsigned k, const uint64_t *keys, const uint16_t *counts,
                  uint64_t n, uint64_t min_count, uint64_t double_hash_size)
{
    unsigned words = (k + 31) / 32;
    uint64_t total = 0;
    for (uint64_t i = 0; i < n; ++i) if (counts[i] >= min_count) ++total;      /* counter.h:612-617 */
    uint64_t size = pbo_load_size(total);
    uint64_t slots = size > double_hash_size ? size : double_hash_size;          /* counter.h:627 */

    occ_map m; m.cap = 16;
    while (m.cap < 2 * total + 16) m.cap *= 2;
    m.slot = (uint64_t *)malloc(m.cap * sizeof(uint64_t));
    m.idx = (uint64_t *)malloc(m.cap * sizeof(uint64_t));
    slot_rec *recs = (slot_rec *)malloc((total ? total : 1) * sizeof(slot_rec));
    if (!m.slot || !m.idx || !recs) { free(m.slot); free(m.idx); free(recs); return PBO_E_NOMEM; }
    memset(m.idx, 0xFF, m.cap * sizeof(uint64_t));

    uint64_t nrec = 0;
    for (uint64_t i = 0; i < n; ++i) {
        if (counts[i] < min_count) continue;
        /* operator[] -> find_any (doubleHash.h:170-184, 221-226); keys are distinct */
        uint64_t home, step;
        pbo_probe_start(keys + i * words, words, slots, &home, &step);
        uint64_t v = home;
        while (occ_find(&m, v) >= 0) v = (v + step) & (slots - 1);
        occ_put(&m, v, i);
        recs[nrec].slot = v; recs[nrec].idx = i; ++nrec;
    }
    free(m.slot); free(m.idx);
    qsort(recs, nrec, sizeof(slot_rec), slot_cmp);

    FILE *fp = fopen(path, "wb");
    if (!fp) { free(recs); return PBO_E_IO; }
    uint64_t k64 = k, index_size = slots - 1;
    fwrite(&k64, 8, 1, fp);                                   /* counter.h:960 */
    fwrite(&index_size, 8, 1, fp);                            /* doubleHash.h:268 */
    uint64_t raw = pbo_key_raw_size(k);
    unsigned char *buf = (unsigned char *)calloc(1, raw + 8 * words);
    for (uint64_t j = 0; j < nrec; ++j) {
        const uint64_t *key = keys + recs[j].idx * words;
        fwrite(&recs[j].slot, 8, 1, fp);                      /* doubleHash.h:272 */
        memset(buf, 0, raw);
        if (k <= 32) {
            memcpy(buf, key, 8);
            fwrite(buf, 1, 8, fp);
        } else if (k <= 160) {
            /* raw object bytes: [vptr][value*][len][entity...] (doubleHash.h:273); vptr/value are
             * process-specific in the reference and ignored by its reader (doubleHash.h:288-291) */
            memcpy(buf + 16, &k64, 8);
            memcpy(buf + 24, key, 8 * words);
            fwrite(buf, 1, raw, fp);
        } else {
            memcpy(buf + 16, &k64, 8);                        /* binstr_t {vptr, value*, len} (binstr.h:36-38, 51) */
            fwrite(buf, 1, raw, fp);
            fwrite(key, 8, words, fp);                        /* doubleHash.h:77-80 */
        }
        fwrite(&counts[recs[j].idx], 2, 1, fp);               /* doubleHash.h:275 */
    }
    free(buf); free(recs);
    return fclose(fp) == 0 ? PBO_OK : PBO_E_IO;
}

void pbo_bin_release(pbo_bin *b)
{
    if (!b) return;
    free(b->slots); free(b->keys); free(b->counts);
    b->slots = NULL; b->keys = NULL; b->counts = NULL;
}

int pbo_read_bin(const char *path, pbo_bin *out)
{
    memset(out, 0, sizeof(*out));
    FILE *fp = fopen(path, "rb");
    if (!fp) return PBO_E_IO;
    if (fread(&out->k, 8, 1, fp) != 1 || fread(&out->index_size, 8, 1, fp) != 1) { fclose(fp); return PBO_E_IO; }
    unsigned k = (unsigned)out->k, words = (k + 31) / 32;
    out->words = words;
    uint64_t raw = pbo_key_raw_size(k);
    uint64_t rec = 8 + raw + (k > 160 ? 8ull * words : 0) + 2;
    long pos = ftell(fp);
    fseek(fp, 0, SEEK_END);
    long end = ftell(fp);
    fseek(fp, pos, SEEK_SET);
    if ((uint64_t)(end - pos) % rec != 0) { fclose(fp); return PBO_E_FORMAT; }
    uint64_t n = (uint64_t)(end - pos) / rec;
    out->n = n;
    out->slots = (uint64_t *)malloc((n ? n : 1) * 8);
    out->keys = (uint64_t *)malloc((n ? n : 1) * 8 * words);
    out->counts = (uint16_t *)malloc((n ? n : 1) * 2);
    unsigned char *buf = (unsigned char *)malloc(rec);
    if (!out->slots || !out->keys || !out->counts || !buf) { fclose(fp); free(buf); pbo_bin_release(out); return PBO_E_NOMEM; }
    for (uint64_t i = 0; i < n; ++i) {
        if (fread(buf, 1, rec, fp) != rec) { fclose(fp); free(buf); return PBO_E_IO; }
        memcpy(&out->slots[i], buf, 8);
        if (k <= 32) memcpy(&out->keys[i], buf + 8, 8);
        else if (k <= 160) memcpy(&out->keys[i * words], buf + 8 + 24, 8 * words);
        else memcpy(&out->keys[i * words], buf + 8 + raw, 8 * words);
        memcpy(&out->counts[i], buf + rec - 2, 2);
    }
    free(buf);
    fclose(fp);
    return PBO_OK;
}

int pbo_bin_check_reachable(const pbo_bin *b)
{
    uint64_t slots = b->index_size + 1;
    if (slots & (slots - 1)) return PBO_E_FORMAT;
    occ_map m; m.cap = 16;
    while (m.cap < 2 * b->n + 16) m.cap *= 2;
    m.slot = (uint64_t *)malloc(m.cap * 8);
    m.idx = (uint64_t *)malloc(m.cap * 8);
    if (!m.slot || !m.idx) { free(m.slot); free(m.idx); return PBO_E_NOMEM; }
    memset(m.idx, 0xFF, m.cap * 8);
    int rc = PBO_OK;
    for (uint64_t i = 0; i < b->n; ++i) {
        if (b->slots[i] >= slots || occ_find(&m, b->slots[i]) >= 0) { rc = PBO_E_FORMAT; goto out; }
        occ_put(&m, b->slots[i], i);
    }
    for (uint64_t i = 0; i < b->n; ++i) {
        uint64_t home, step;
        pbo_probe_start(b->keys + i * b->words, b->words, slots, &home, &step);
        uint64_t v = home, guard = 0;
        for (;;) {                                             /* find_any, doubleHash.h:170-184 */
            int64_t at = occ_find(&m, v);
            if (at < 0) { rc = PBO_E_FORMAT; goto out; }       /* hits an empty slot first: lost */
            if ((uint64_t)at == i) break;
            if (pbo_key_cmp(b->keys + (uint64_t)at * b->words, b->keys + i * b->words, b->words) == 0) { rc = PBO_E_FORMAT; goto out; }
            v = (v + step) & (slots - 1);
            if (++guard > b->n + 1) { rc = PBO_E_FORMAT; goto out; }
        }
    }
out:
    free(m.slot); free(m.idx);
    return rc;
}

/* Counter::outputOccurrenceDistribution (counter.h:1000-1007) */
int pbo_write_tsv(const char *path, const uint64_t *occ, uint64_t max_occ)
{
    FILE *fp = fopen(path, "w");
    if (!fp) return PBO_E_IO;
    for (uint64_t i = 1; i <= max_occ; ++i)
        fprintf(fp, "%llu\t%llu\n", (unsigned long long)i, (unsigned long long)occ[i]);
    return fclose(fp) == 0 ? PBO_OK : PBO_E_IO;
}
