// pbk_ingest.hpp -- FASTA/FASTQ text -> reads, with the exact semantics of the reference's serial parsers
// (Assemble::readFastaUncompressed / readFastqUncompressed, assemble.cpp:816-848, 902-942) but parallel over byte
// ranges of one file.
//
// Why ranges can be parsed independently: in both reference loops a header line ('>' resp. '@' as first character)
// puts the parser into one and the same state whatever came before -- pending read flushed if non-empty, read = "",
// FASTQ flag = true.  So every header line is a synchronisation point: a worker that starts AT a header line with an
// empty read and stops BEFORE the first header line at or after the end of its range (flushing its pending read, which
// is what that header line would have done) produces exactly the reads the serial loop produces for those lines.  This
// also reproduces the reference's quirks: a FASTQ quality line that starts with '@' acts as a header, lines after
// it are taken as sequence until the next '+' line; the read pending at the end of the file is flushed
// unconditionally, even when it is empty (assemble.cpp:844-845, 938-939).
#ifndef PBK_INGEST_HPP
#define PBK_INGEST_HPP

#include <cstddef>
#include <cstring>
#include <string>

namespace pbk {
namespace ingest {

struct Line { const char *s; size_t len; };

struct LineReader {                      // std::getline semantics: '\n' stripped, the last line may lack it
    const char *cur, *end;
    LineReader(const char *p, size_t n) : cur(p), end(p + n) {}
    bool next(Line &l)
    {
        if (cur >= end) return false;
        const char *nl = (const char *)memchr(cur, '\n', (size_t)(end - cur));
        l.s = cur;
        if (nl) { l.len = (size_t)(nl - cur); cur = nl + 1; } else { l.len = (size_t)(end - cur); cur = end; }
        return true;
    }
};

// where parsed reads go: SEQ::convertFromString + writeTemporaryFile in the reference (common.h:460, 426)
struct ReadSink {
    std::string read;                    // the read being assembled from lines
    virtual ~ReadSink() {}
    void append(const Line &l) { read.append(l.s, l.len); }
    void flush() { emit(); read.clear(); }
    virtual void emit() = 0;             // consume `read` (implementations throw on length >= MAX_READ_LEN, common.h:465)
};

// offset of the first line that starts at or after `from` and begins with `mark`; n if there is none
inline size_t sync_point(const char *p, size_t n, size_t from, char mark)
{
    size_t q = from;
    if (q > 0 && q <= n && p[q - 1] != '\n') {           // inside a line: go to the start of the next one
        const char *nl = (const char *)memchr(p + q, '\n', n - q);
        if (!nl) return n;
        q = (size_t)(nl - p) + 1;
    }
    while (q < n) {
        if (p[q] == mark) return q;
        const char *nl = (const char *)memchr(p + q, '\n', n - q);
        if (!nl) return n;
        q = (size_t)(nl - p) + 1;
    }
    return n;
}

// Lines [s_begin, s_end) of the file, s_begin a synchronisation point (or n).  `final_flush`: this worker owns the end
// of the file and performs the reference's unconditional last flush.
inline void parse_range(const char *p, size_t s_begin, size_t s_end, bool fastq, bool final_flush, ReadSink &out)
{
    const char mark = fastq ? '@' : '>';
    LineReader r(p + s_begin, s_end - s_begin);
    Line l;
    bool flag = true;
    while (r.next(l)) {
        if (fastq) {
            if (l.len == 0) continue;                                   // assemble.cpp:922
            if (l.s[0] != mark) {
                if (flag && l.s[0] != '+') out.append(l); else flag = false;
            } else {
                if (!out.read.empty()) out.flush();
                flag = true;
            }
        } else {
            if (!(l.len && l.s[0] == mark)) out.append(l);
            else if (!out.read.empty()) out.flush();
        }
    }
    if (final_flush) out.flush();                                       // also when the read is empty
    else if (!out.read.empty()) out.flush();                            // what the next worker's header line would do
}

// The byte ranges [i n / T, (i + 1) n / T) of a file as synchronisation points: s[0..T], s[T] = n.  Worker t parses
// [s[t], s[t+1]); `final_owner` is the worker that performs the unconditional last flush.
struct Plan { size_t s[65]; unsigned n_workers, final_owner; };
inline Plan plan_ranges(const char *p, size_t n, bool fastq, unsigned T)
{
    Plan pl;
    if (T < 1) T = 1;
    if (T > 64) T = 64;
    pl.n_workers = T;
    const char mark = fastq ? '@' : '>';
    for (unsigned t = 0; t < T; ++t) pl.s[t] = sync_point(p, n, (size_t)((unsigned long long)n * t / T), mark);
    pl.s[T] = n;
    for (unsigned t = 1; t <= T; ++t) if (pl.s[t] < pl.s[t - 1]) pl.s[t] = pl.s[t - 1];   // (monotone by construction; belt and braces)
    pl.final_owner = 0;                                                  // no header line at all: worker 0 emits the empty read
    for (unsigned t = 0; t < T; ++t) if (pl.s[t] < n) pl.final_owner = t;    // the last worker that has lines to parse
    return pl;
}

}  // namespace ingest
}  // namespace pbk

#endif  // PBK_INGEST_HPP
