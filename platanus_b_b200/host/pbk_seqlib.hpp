// pbk_seqlib.hpp -- the paired / tagged read ingest of the reference's other commands (scaffold, gap_close, polish ...):
// ReadFastaSingleMT / ReadFastaPairMT / their *Tagged* forms (seqlib.cpp:365-443, 445-543, 544-636, 638-742) with the record
// readers they are built on (seqlib.cpp:740-953).  SURVEY.md section 8f row 3: same SEQ temp-file format as the counting path
// (common.h:426-433), same serial `getline` bottleneck in the reference.
//
// What is reproduced bit for bit: the per-thread temp files lib[i].pairFP (pair j goes to file j % numThread: forward record,
// then reverse record, each `int numUnknown (always 0 on this path: SEQ::put never fills the N list) | int length | one
// byte per base, Char2Bin codes, N = 4`, tagged forms followed by `int tagID`), lib[0]'s totalLength / numPair, and the
// FormatError conditions (odd number of reads, files of different length).
//
// How it differs from the reference's loop: the files are memory-mapped and each is cut into lines and converted by its OWN
// thread (two in pair mode) into chunks of ready-made records; the calling thread only zips the chunks into the per-thread
// output buffers, which are written with large fwrite calls.  The line grammar is restated exactly, including std::getline's
// end-of-file behaviour (an unterminated last line is a line and sets eof; a failed read sets fail), because the
// reference's loops are controlled by those bits.
#ifndef PBK_SEQLIB_HPP
#define PBK_SEQLIB_HPP

#include "pbk_counter.hpp"
#include "pbk_ingest.hpp"

#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace pbk {
namespace seqlib {

struct FormatError : ErrorBase { explicit FormatError(const std::string &m) : ErrorBase(E_FORMAT, "Error, File format exception!!\n" + m) {} };

// the part of SeqLib (seqlib.h:31-98) these functions touch
struct PairLibrary {
    std::vector<FILE *> pairFP;        // lib[i].pairFP, i < numThread
    long numPair = 0, totalLength = 0; // lib[0].addNumPair / addTotalLength
};

// platanus::Char2Bin (common.h:256): only the low nibble counts
inline uint8_t char2bin(unsigned char c)
{
    static const unsigned char table[17] = ".\x0.\x1\x3..\x2......\x4";
    return table[c & 0xF];
}

// std::getline on an ifstream, over a mapped file
struct LineReader {
    const char *p, *end;
    bool eof, fail;
    LineReader(const char *b, size_t n) : p(b), end(b + n), eof(false), fail(false) {}
    bool good() const { return !fail; }                       // `while (ifs && ...)`: operator bool is !fail()
    bool getline(const char *&line, size_t &len)
    {
        if (fail) return false;
        if (p == end) { eof = true; fail = true; len = 0; return false; }
        const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
        line = p;
        if (nl) { len = (size_t)(nl - p); p = nl + 1; }
        else { len = (size_t)(end - p); p = end; eof = true; }
        return true;
    }
};

struct Record { uint64_t off; uint32_t len; int32_t tag; };
struct Chunk {
    std::string codes;                 // Char2Bin codes of the chunk's records, back to back
    std::vector<Record> rec;
    bool eof_after = false;            // stream state after the chunk's last record: the reference's loops test eof() there
};

// one record, as ReadFast[aq]Seq[Tagged]Uncompressed read it (seqlib.cpp:793-840, 888-940)
inline void put(std::string &codes, const char *line, size_t len)
{
    const size_t at = codes.size();
    codes.resize(at + len);
    for (size_t i = 0; i < len; ++i) codes[at + i] = (char)char2bin((unsigned char)line[i]);
}
inline bool tag_position_inline(const char *line, size_t len, size_t &start, size_t &end)     // seqlib.cpp:866-884
{
    static const char T[] = "BX:Z:";
    const char *hit = len >= 5 ? (const char *)memmem(line, len, T, 5) : NULL;
    if (!hit) return false;
    start = (size_t)(hit - line) + 5;
    for (end = start + 1; end < len; ++end)
        if (!isalnum((unsigned char)line[end])) return true;
    return true;
}
inline void read_record(LineReader &in, bool fastq, bool tagged, std::unordered_map<std::string, int> *tags, Chunk &c)
{
    Record r;
    r.off = c.codes.size(); r.tag = 0;
    const char *line; size_t len;
    if (tagged) {                                            // the header line itself is read here (no header skip in the tagged loops)
        r.tag = -1;
        if (in.getline(line, len)) {
            size_t s = 0, e = 0;
            if (tag_position_inline(line, len, s, e)) {      // the reference's operator[] gives 0 for a tag it has never seen; a lookup
                const auto it = tags->find(std::string(line + s, e - s));   // without the insertion is safe from two parser threads
                r.tag = it == tags->end() ? 0 : it->second;
            }
        }
    }
    unsigned read_line = 0;
    while (in.good() && in.getline(line, len)) {
        if (len > 0 && line[0] == (fastq ? '+' : '>')) break;
        ++read_line;
        put(c.codes, line, len);
    }
    if (fastq) {                                             // quality lines (+ the next header when the header skip is ours)
        const unsigned skip = tagged ? read_line : read_line + 1;
        for (unsigned i = 0; i < skip; ++i) in.getline(line, len);
    }
    r.len = (uint32_t)(c.codes.size() - r.off);
    c.rec.push_back(r);
    c.eof_after = in.eof;
}
inline void read_header(LineReader &in, bool fastq)          // ReadFast[aq]HeaderUncompressed (seqlib.cpp:740-772)
{
    const char *line; size_t len;
    while (in.good() && in.getline(line, len))
        if (len > 0 && line[0] == (fastq ? '@' : '>')) break;
}

// a file parsed on its own thread into chunks of records
class RecordStream {
public:
    RecordStream(const std::string &path, bool fastq, bool tagged, std::unordered_map<std::string, int> *tags, const std::string &tmp_dir)
        : fastq_(fastq), tagged_(tagged), tags_(tags), done_(false), at_(0), cur_(NULL)
    {
        const int rc = map_.map(path, tmp_dir);              // plain, gzip or bzip2 (by magic number, pbk_ingest.hpp)
        if (rc == -2) throw TMPError();
        if (rc != 0) throw FILEError(path);
        worker_ = std::thread([this]() { produce(); });
    }
    ~RecordStream() { { std::lock_guard<std::mutex> g(m_); stop_ = true; } cv_.notify_all(); worker_.join(); delete cur_; }
    // the reference's loop condition `ifs && !ifs.eof()` before a record is read
    bool more()
    {
        if (cur_ && at_ < cur_->rec.size()) return true;
        return next_chunk() ? true : false;
    }
    // eof() after the most recently returned record
    bool eof_now() { return (cur_ && at_ < cur_->rec.size()) ? false : peek_end(); }
    const Record &next(const char *&codes) { const Record &r = cur_->rec[at_++]; codes = cur_->codes.data() + r.off; return r; }

private:
    void produce()
    {
        LineReader in(map_.p, map_.n);
        if (!tagged_) read_header(in, fastq_);
        while (in.good() && !in.eof) {                        // one record per iteration, like the reference's loops
            Chunk *c = new Chunk();
            while (in.good() && !in.eof && c->rec.size() < 65536 && c->codes.size() < ((size_t)16 << 20)) read_record(in, fastq_, tagged_, tags_, *c);
            std::unique_lock<std::mutex> g(m_);
            cv_.wait(g, [this]() { return q_.size() < 4 || stop_; });
            if (stop_) { delete c; return; }
            q_.push_back(c);
            cv_.notify_all();
        }
        std::lock_guard<std::mutex> g(m_);
        done_ = true;
        cv_.notify_all();
    }
    bool next_chunk()
    {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [this]() { return !q_.empty() || done_; });
        if (q_.empty()) return false;
        delete cur_;
        cur_ = q_.front(); q_.pop_front(); at_ = 0;
        cv_.notify_all();
        return !cur_->rec.empty();
    }
    bool peek_end()                                          // no record left in the current chunk: is the stream at its end?
    {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [this]() { return !q_.empty() || done_; });
        return q_.empty();
    }
    ingest::MappedFile map_;
    bool fastq_, tagged_;
    std::unordered_map<std::string, int> *tags_;
    std::thread worker_;
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<Chunk *> q_;
    bool done_, stop_ = false;
    size_t at_;
    Chunk *cur_;
};

// per-thread output buffers, flushed with large writes
class PairWriter {
public:
    PairWriter(PairLibrary &lib, int numThread) : lib_(lib), buf_(numThread), i_(0)
    {
        for (int t = 0; t < numThread; ++t) fseek(lib.pairFP[t], 0, SEEK_END);          // seqlib.cpp:387-388
    }
    ~PairWriter() { for (size_t t = 0; t < buf_.size(); ++t) flush(t); }
    void pair(const Record &f, const char *fc, const Record &r, const char *rc, bool isMate, bool tagged)
    {
        std::string &b = buf_[i_];
        seq(b, f, fc, isMate, tagged);
        seq(b, r, rc, isMate, tagged);
        if (b.size() >= ((size_t)8 << 20)) flush(i_);
        i_ = (i_ + 1) % buf_.size();
        lib_.totalLength += (long)f.len + (long)r.len;
        lib_.numPair += 1;
    }

private:
    static void seq(std::string &b, const Record &r, const char *codes, bool isMate, bool tagged)
    {
        const int32_t zero = 0, len = (int32_t)r.len;
        b.append((const char *)&zero, 4);                    // numUnknown: SEQ::put leaves it 0, the N stay in base[] as 4
        b.append((const char *)&len, 4);
        const size_t at = b.size();
        b.append(codes, r.len);
        if (isMate)                                          // SEQ::reverse (common.h:414-423): reverse complement, 4 stays 4
            for (uint32_t i = 0; i < r.len; ++i) { const char c = codes[r.len - 1 - i]; b[at + i] = c != 4 ? (char)(c ^ 0x3) : (char)4; }
        if (tagged) b.append((const char *)&r.tag, 4);
    }
    void flush(size_t t)
    {
        if (!buf_[t].empty() && fwrite(buf_[t].data(), 1, buf_[t].size(), lib_.pairFP[t]) != buf_[t].size()) throw TMPError();
        buf_[t].clear();
    }
    PairLibrary &lib_;
    std::vector<std::string> buf_;
    size_t i_;
};

// ReadFastaSingleMT / ReadFastaSingleTaggedMT (seqlib.cpp:365-443, 445-543): one interleaved file
inline void ReadFastaSingleMT(PairLibrary &lib, const std::string &filename, int numThread, bool isMate = false, bool isFastq = false,
                              bool notPair = false, std::unordered_map<std::string, int> *tagStringConverter = NULL, const std::string &tmp_dir = ".")
{
    const bool tagged = tagStringConverter != NULL;
    RecordStream in(filename, isFastq, tagged, tagStringConverter, tmp_dir);
    PairWriter out(lib, numThread);
    const Record none = {0, 0, tagged ? -1 : 0};
    while (in.more()) {
        const char *fc, *rc = "";
        const Record f = in.next(fc);
        if (!notPair && in.eof_now()) throw FormatError("the number of read is odd in file.");
        Record r = none;
        if (in.more()) r = in.next(rc);                      // (past the end the reference reads an empty record)
        if (tagged && (f.tag == -1 || r.tag == -1)) continue;
        out.pair(f, fc, r, rc, isMate, tagged);
    }
}

// ReadFastaPairMT / ReadFastaPairTaggedMT (seqlib.cpp:544-636, 638-742): mates in two files
inline void ReadFastaPairMT(PairLibrary &lib, const std::string &filename1, const std::string &filename2, int numThread, bool isMate = false,
                            bool isFastq = false, std::unordered_map<std::string, int> *tagStringConverter = NULL, const std::string &tmp_dir = ".")
{
    const bool tagged = tagStringConverter != NULL;
    RecordStream in1(filename1, isFastq, tagged, tagStringConverter, tmp_dir);
    RecordStream in2(filename2, isFastq, tagged, tagStringConverter, tmp_dir);
    PairWriter out(lib, numThread);
    bool more1 = in1.more(), more2 = in2.more();
    while (more1 && more2) {
        const char *fc, *rc;
        const Record f = in1.next(fc), r = in2.next(rc);
        // tagged pairs are kept only when both mates carry the SAME tag (seqlib.cpp:674)
        if (!tagged || (f.tag == r.tag && f.tag != -1)) out.pair(f, fc, r, rc, isMate, tagged);
        more1 = in1.more(); more2 = in2.more();
    }
    if (more1 || more2) throw FormatError("the number of read is different in paired-file.");
}

}  // namespace seqlib
}  // namespace pbk

#endif  // PBK_SEQLIB_HPP
