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
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>

namespace pbk {
namespace ingest {

// ---- input files: plain, gzip or bzip2 (platanus::checkFileCompression / openFileAllowingCompression,
//      common.cpp:88-143; readFastaCompressed / readFastqCompressed, assemble.cpp:851-885, 945-986) ----------------------
// The reference asks `file -bL` whether a file is "gzip compressed" / "bzip2 compressed" and then reads the output of
// `gzip -cd FILE` / `bzip2 -cd FILE` through a pipe with the same line loop as for plain files.  Here the type comes
// from the magic numbers `file` itself goes by (1f 8b; "BZh"), and the same two programs write the plain text into
// an unlinked temporary file in the -tmp directory, which is then mapped and parsed in ranges like any other input.
// As in the reference, the decompressor's exit status is not looked at: a truncated archive yields the reads that
// came out of it.
enum Compression { UNCOMPRESSED = 0, GZIP = 1, BZIP2 = 2 };

inline Compression sniff_compression(int fd)
{
    unsigned char m[3] = {0, 0, 0};
    const ssize_t got = pread(fd, m, 3, 0);
    if (got >= 2 && m[0] == 0x1f && m[1] == 0x8b) return GZIP;
    if (got >= 3 && m[0] == 'B' && m[1] == 'Z' && m[2] == 'h') return BZIP2;
    return UNCOMPRESSED;
}

// fd of an unlinked temp file holding the decompressed bytes of `name`; -1 if the temp file cannot be made
inline int decompress_to_tmp(const std::string &name, Compression c, const std::string &tmp_dir)
{
    std::string templ = tmp_dir + "/XXXXXX";
    std::vector<char> buf(templ.begin(), templ.end());
    buf.push_back('\0');
    const int out = mkstemp(buf.data());
    if (out < 0) return -1;
    unlink(buf.data());
    const pid_t pid = fork();
    if (pid < 0) { close(out); return -1; }
    if (pid == 0) {                                          // child: PROGRAM -cd NAME > temp file (no shell in between)
        dup2(out, 1);
        close(out);
        const char *prog = c == GZIP ? "gzip" : "bzip2";
        execlp(prog, prog, "-cd", name.c_str(), (char *)NULL);
        _exit(127);
    }
    int status = 0;
    while (waitpid(pid, &status, 0) < 0) {}
    return out;
}

struct MappedFile {
    const char *p; size_t n; int fd; Compression compression;
    MappedFile() : p(NULL), n(0), fd(-1), compression(UNCOMPRESSED) {}
    ~MappedFile() { unmap(); }
    // 0 = ok, -1 = cannot open / map the file (FILEError in the reference), -2 = no temp file (TMPError)
    int map(const std::string &name, const std::string &tmp_dir)
    {
        unmap();
        fd = open(name.c_str(), O_RDONLY);
        if (fd < 0) return -1;
        compression = sniff_compression(fd);
        if (compression != UNCOMPRESSED) {
            const int plain = decompress_to_tmp(name, compression, tmp_dir);
            close(fd);
            fd = plain;
            if (fd < 0) return -2;
        }
        struct stat st;
        if (fstat(fd, &st) != 0) { unmap(); return -1; }
        n = (size_t)st.st_size;
        if (n) {
            void *m = mmap(NULL, n, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m == MAP_FAILED) { n = 0; unmap(); return -1; }
            madvise(m, n, MADV_SEQUENTIAL);
            p = (const char *)m;
        }
        return 0;
    }
    void unmap()
    {
        if (p) munmap((void *)p, n);
        if (fd >= 0) close(fd);
        p = NULL; n = 0; fd = -1;
    }
private:
    MappedFile(const MappedFile &);
    MappedFile &operator=(const MappedFile &);
};

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
    void flush() { emit_span(read.data(), read.size()); read.clear(); }
    // consume one read (implementations throw on length >= MAX_READ_LEN, common.h:465).  Single-line records -- the usual
    // FASTQ -- are handed over straight from the mapped file, without a detour through `read`.
    virtual void emit_span(const char *s, size_t len) = 0;
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
    // `early`: the record in progress was a single line whose end was in sight (the next line closes it: '+' in FASTQ,
    // a header in FASTA), so it has been emitted already and `read` stays empty until the next header.  The reference
    // would emit it at that header, or at the end of the file: same reads, same order.
    bool early = false;
    while (r.next(l)) {
        if (fastq) {
            if (l.len == 0) continue;                                   // assemble.cpp:922
            if (l.s[0] != mark) {
                if (flag && l.s[0] != '+') {
                    // (after an early emit the next line starts with '+' and clears `flag`: no line is appended while `early`)
                    if (out.read.empty() && r.cur < r.end && *r.cur == '+') { out.emit_span(l.s, l.len); early = true; }
                    else out.append(l);
                } else flag = false;
            } else {
                if (!out.read.empty()) out.flush();
                flag = true; early = false;
            }
        } else {
            if (!(l.len && l.s[0] == mark)) {
                if (out.read.empty() && l.len && r.cur < r.end && *r.cur == mark) { out.emit_span(l.s, l.len); early = true; }
                else out.append(l);
            } else {
                if (!out.read.empty()) out.flush();
                early = false;
            }
        }
    }
    if (final_flush) { if (!early) out.flush(); }                       // also when the read is empty (assemble.cpp:844-845, 938-939)
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
