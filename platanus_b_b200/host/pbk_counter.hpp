// pbk_counter.hpp -- host C++ side of the drop-in: the part of the reference's Counter<KMER>
// (counter.h:36-201) that `platanus_b assemble -kmer_occ_only` uses, re-implemented on top of the
// libpbk C ABI (include/pbk.h).  Same member names, argument meaning, return values and error
// behaviour as the reference, so Assemble::initialKmerAssemble (assemble.cpp:303-350) reads the same
// with either class; INTEGRATION.md shows the wiring.  Keys never cross this boundary as binstr
// objects: they are plain little-endian u64 words, value[0] (the last 32 bases) first, exactly what
// Kmer31::writeKey / KmerN::writeKey put into kmerFP (kmer.h:119, 237).
//
// There is no CPU fallback: every counting member throws pbk::GPUError when libpbk cannot reach a
// CUDA device.
#ifndef PBK_COUNTER_HPP
#define PBK_COUNTER_HPP

#include "../../include/pbk.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include <unistd.h>

namespace pbk {

// platanus::ERROR ids (common.h:55-56): the process exit code of the reference's main (main.cpp:121-124)
enum ErrorId { E_IO = 0, E_FOPEN = 1, E_TMP = 2, E_FORMAT = 3, E_READ = 4, E_KMERDIST = 6, E_GPU = 64 };

class ErrorBase {
public:
    ErrorBase(int id, const std::string &msg) : id_(id), msg_(msg) {}
    void showErrorMessage() const { std::cerr << "Error(" << id_ << "): " << msg_ << std::endl; }   // common.h:100-103
    int getID() const { return id_; }
    const std::string &what() const { return msg_; }
private:
    int id_;
    std::string msg_;
};
struct FILEError : ErrorBase { explicit FILEError(const std::string &f) : ErrorBase(E_FOPEN, "Error, File not open exception!!\n" + f + " is not open!!") {} };
struct TMPError : ErrorBase { TMPError() : ErrorBase(E_TMP, "Error, temporary file exception!!") {} };
struct ReadError : ErrorBase { explicit ReadError(const std::string &m = "") : ErrorBase(E_READ, "Error, Read file exception!!" + (m.empty() ? m : "\n" + m)) {} };
struct KmerDistError : ErrorBase { KmerDistError() : ErrorBase(E_KMERDIST, "Error, kmer distribution exception!!\nkmer distribution can't be calculated.") {} };
struct GPUError : ErrorBase { explicit GPUError(const std::string &m) : ErrorBase(E_GPU, "Error, GPU k-mer counter exception!!\n" + m) {} };

class Counter {
public:
    typedef unsigned long long u64_t;

    FILE *kmerFP;      // kept for source compatibility (counter.h:77); the (key, count) stream stays in HBM

    Counter() : kmerFP(NULL), ctx_(NULL), group_(NULL), kmerLength_(0), maxOccurrence_(0), doubleHashSize_(0), nInstances_(0), nDistinct_(0),
                device_(-1), flags_(0), numDevices_(1) {}
    explicit Counter(u64_t k) : kmerFP(NULL), ctx_(NULL), group_(NULL), kmerLength_(k), maxOccurrence_(0), doubleHashSize_(0), nInstances_(0),
                                nDistinct_(0), device_(-1), flags_(0), numDevices_(1) {}
    ~Counter() { if (ctx_) pbk_destroy(ctx_); if (group_) pbk_group_destroy(group_); }
    Counter(const Counter &) = delete;
    Counter &operator=(const Counter &) = delete;

    void setDevice(int device) { device_ = device; }
    // count on the first n visible GPUs (CUDA_VISIBLE_DEVICES selects which): the library shards the keys by hash range and
    // exchanges them between the devices itself (pbk_group_*).  Must be called before the first beginCounting.  The counting
    // members (beginCounting .. exportKmers and what builds on them) work on the group; the table consumers (occurrenceArray,
    // matchReads, seedEntries, pushContigs, readOccurrenceTableBinary) need a single device.
    void setNumDevices(unsigned n) { numDevices_ = n < 1 ? 1 : n; deviceList_.clear(); }
    // the same with explicit device ordinals (an ordinal may repeat: logical shards on one GPU, used by the tests)
    void setDevices(const std::vector<int32_t> &devices) { deviceList_ = devices; numDevices_ = devices.empty() ? 1u : (unsigned)devices.size(); }
    unsigned getNumDevices() const { return numDevices_; }
    void setFlags(unsigned flags) { flags_ = flags; }
    pbk_ctx *context() { return ctx_; }

    // ---- getters (counter.h:86-92) ------------------------------------------------------------
    u64_t getMaxOccurrence() const { return maxOccurrence_; }
    u64_t getLengthDistributionI(u64_t position) const { return lengthDistribution_[position]; }
    u64_t getKmerLength() const { return kmerLength_; }
    void setKmerLength(u64_t k) { kmerLength_ = k; }
    u64_t getNumInstances() const { return nInstances_; }
    u64_t getNumDistinct() const { return nDistinct_; }
    const std::vector<u64_t> &occurrenceDistribution() const { return occurrenceDistribution_; }

    // ---- counter.h:146-150, 221-238 -------------------------------------------------------------
    double calcLengthDistributionAverage(u64_t start, u64_t end) const { return average(lengthDistribution_, start, end); }
    double calcOccurrenceDistributionAverage(u64_t start, u64_t end) const { return average(occurrenceDistribution_, start, end); }

    // ---- counter.h:245-267 ----------------------------------------------------------------------
    u64_t getLeftLocalMinimalValue(u64_t windowSize) const
    {
        return pbk_left_local_min((const uint64_t *)occurrenceDistribution_.data(), maxOccurrence_, windowSize);
    }

    // ---- the wide seam: reads straight from the parser (what Assemble::readInputFile hands to
    //      SEQ::convertFromString, assemble.cpp:790-986), no temp-file detour ----------------------
    void beginCounting(u64_t kLength)
    {
        kmerLength_ = kLength;
        keptValid_ = false;
        if (group_) { check(pbk_group_reset(group_, (uint32_t)kLength), "pbk_group_reset"); return; }
        if (ctx_) { check(pbk_reset(ctx_, (uint32_t)kLength), "pbk_reset"); return; }
        pbk_config cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.struct_size = sizeof cfg; cfg.k = (uint32_t)kLength; cfg.device = device_; cfg.flags = flags_;
        if (numDevices_ > 1) {
            const int rc = pbk_group_create(&group_, &cfg, deviceList_.empty() ? NULL : deviceList_.data(), numDevices_);
            if (rc != PBK_OK) { group_ = NULL; throw GPUError(pbk_strerror(rc)); }
            return;
        }
        const int rc = pbk_create(&ctx_, &cfg);
        if (rc != PBK_OK) { ctx_ = NULL; throw GPUError(pbk_strerror(rc)); }
    }
    // ASCII bases of n reads, concatenated; offsets[n + 1]
    void pushReads(const uint8_t *bases, const uint64_t *offsets, uint64_t n)
    {
        if (group_) check(pbk_group_push_reads(group_, bases, offsets, n, PBK_ENC_ASCII, NULL, NULL), "pbk_group_push_reads");
        else check(pbk_push_reads(ctx_, bases, offsets, n, PBK_ENC_ASCII, NULL, NULL), "pbk_push_reads");
    }
    // finish: fills lengthDistribution, occurrenceDistribution, maxOccurrence; returns doubleHashSize
    u64_t endCounting(u64_t memory)
    {
        occurrenceDistribution_.assign(PBK_OCC_BINS, 0);
        lengthDistribution_.assign(PBK_LEN_BINS, 0);
        uint64_t nd = 0, ni = 0, mx = 0;
        if (group_)
            check(pbk_group_finalize(group_, (uint64_t *)occurrenceDistribution_.data(), (uint64_t *)lengthDistribution_.data(), &nd, &ni, &mx),
                  "pbk_group_finalize");
        else
            check(pbk_finalize(ctx_, (uint64_t *)occurrenceDistribution_.data(), (uint64_t *)lengthDistribution_.data(), &nd, &ni, &mx),
                  "pbk_finalize");
        nDistinct_ = nd; nInstances_ = ni;
        if (nd) maxOccurrence_ = mx;                       // counter.h:371-376 leaves it unchanged when nothing was counted
        doubleHashSize_ = pbk_double_hash_size(memory, (uint32_t)kmerLength_);
        return doubleHashSize_;
    }

    // ---- the narrow seam: counter.h:276-383 -------------------------------------------------------
    // readFP[0..numThread) are the per-thread SEQ temp files (common.h:426-448).  They are left
    // readable (counter.h:874 rewinds them again).  numThread only says how many files there are.
    u64_t makeKmerReadDistributionMT(u64_t kLength, FILE **readFP, u64_t memory, u64_t numThread)
    {
        beginCounting(kLength);
        pushSeqTempFiles(readFP, numThread);
        return endCounting(memory);
    }
    // the reads of numThread SEQ temp files (common.h:426-448), in batches of 256 MB, through PBK_ENC_PLATANUS
    void pushSeqTempFiles(FILE **readFP, u64_t numThread)
    {
        std::vector<uint8_t> bases;
        std::vector<uint64_t> offsets(1, 0), nposOff(1, 0);
        std::vector<int32_t> npos;
        const size_t BATCH = (size_t)256 << 20;
        for (u64_t t = 0; t < numThread; ++t) {
            FILE *fp = readFP[t];
            rewind(fp);
            int32_t numUnknown, length;
            while (fread(&numUnknown, sizeof(int32_t), 1, fp) == 1) {
                const size_t n0 = npos.size();
                npos.resize(n0 + (size_t)numUnknown);
                if (numUnknown > 0 && fread(&npos[n0], sizeof(int32_t), (size_t)numUnknown, fp) != (size_t)numUnknown) throw ReadError();
                if (fread(&length, sizeof(int32_t), 1, fp) != 1) throw ReadError();
                const size_t b0 = bases.size();
                bases.resize(b0 + (size_t)length);
                if (length > 0 && fread(&bases[b0], 1, (size_t)length, fp) != (size_t)length) throw ReadError();
                offsets.push_back(bases.size());
                nposOff.push_back(npos.size());
                if (bases.size() >= BATCH) flushPlatanus(bases, offsets, npos, nposOff);
            }
        }
        flushPlatanus(bases, offsets, npos, nposOff);
    }

    // ---- counter.h:967-993: PREFIX_kmer_occ.bin -> table (device resident), occurrenceDistribution, maxOccurrence ------
    void readOccurrenceTableBinary(const std::string &filename)
    {
        uint32_t k = 0;
        uint64_t indexSize = 0, n = 0, *keys = NULL;
        uint16_t *counts = NULL;
        const int rc = pbk_read_kmer_occ_bin(filename.c_str(), &k, &indexSize, &keys, &counts, &n);
        if (rc != PBK_OK) throw FILEError(filename);
        struct Release { uint64_t *k; uint16_t *c; ~Release() { pbk_free(k); pbk_free(c); } } release = {keys, counts};
        beginCounting(k);                                            // counter.h:972: kmerLength comes from the file
        check(pbk_load_entries(ctx_, keys, counts, n), "pbk_load_entries");
        occurrenceDistribution_.assign(PBK_OCC_BINS, 0);
        uint64_t nd = 0, ni = 0, mx = 0;
        check(pbk_finalize(ctx_, (uint64_t *)occurrenceDistribution_.data(), NULL, &nd, &ni, &mx), "pbk_finalize");
        nDistinct_ = nd; nInstances_ = ni;
        if (nd) maxOccurrence_ = mx;
        doubleHashSize_ = indexSize + 1;
    }

    // ---- ContigDivider::getOccurrenceArray (kmer_divide.cpp:151-197), which calls Counter::findValue once per window:
    //      the whole array in one device pass.  ASCII sequences, concatenated; occ[offsets[r] + i] = occurrence of the
    //      k-mer starting at base i of sequence r (0: absent, contains N, or no window starts there).
    void occurrenceArray(const uint8_t *bases, const uint64_t *offsets, uint64_t n, uint16_t *occ)
    {
        check(pbk_lookup(ctx_, bases, offsets, n, PBK_ENC_ASCII, NULL, NULL, occ), "pbk_lookup");
    }

    // ---- Counter::pickupReadMatchedEdgeKmer (counter.h:870-910) on reads held in memory: matched[r] = the read stays
    void matchReads(const uint8_t *bases, const uint64_t *offsets, uint64_t n, uint8_t *matched)
    {
        check(pbk_match_reads(ctx_, bases, offsets, n, PBK_ENC_ASCII, NULL, NULL, matched), "pbk_match_reads");
    }

    // the same on reads in the SEQ temp-file form (bytes 0..3 + N position list), as the patched reference holds them
    void matchReadsPlatanus(const uint8_t *bases, const uint64_t *offsets, uint64_t n, const int32_t *npos, const uint64_t *nposOffsets,
                            uint8_t *matched)
    {
        static const int32_t none = 0;
        check(pbk_match_reads(ctx_, bases, offsets, n, PBK_ENC_PLATANUS, npos ? npos : &none, nposOffsets, matched), "pbk_match_reads");
    }
    // (key, value) entries into the table: what setOccurrenceValue (counter.h:92) does on the host DoubleHash
    void loadEntries(const uint64_t *keys, const uint16_t *values, uint64_t n)
    {
        check(pbk_load_entries(ctx_, keys, values, n), "pbk_load_entries");
    }

    // ---- Counter::makeKmerReadDistributionConsideringPreviousGraph (counter.h:663-750), wide seam: beginCounting(k);
    //      seedEntries(table of the previous round's contigs, assemble.cpp:400-403); pushReads(...); endCounting(memory)
    void seedEntries(const uint64_t *keys, const uint16_t *values, uint64_t n)
    {
        check(pbk_seed_entries(ctx_, keys, values, n), "pbk_seed_entries");
    }

    // ---- Counter::makeKmerReadDistributionFromContig (counter.h:511-593) on contigs held in memory (ASCII, concatenated):
    //      beginCounting(k); pushContigs(...); endCounting(memory) leaves the distribution writeKmerDistribution would
    void pushContigs(const uint8_t *bases, const uint64_t *offsets, uint64_t n, const uint16_t *coverage, u64_t minOccurrence)
    {
        check(pbk_push_contigs(ctx_, bases, offsets, n, coverage, minOccurrence), "pbk_push_contigs");
    }

    // ---- counter.h:1000-1007 ----------------------------------------------------------------------
    void outputOccurrenceDistribution(const std::string &filename)
    {
        if (pbk_write_frq_tsv(filename.c_str(), (const uint64_t *)occurrenceDistribution_.data(), maxOccurrence_) != PBK_OK) throw FILEError(filename);
    }

    // ---- counter.h:917-951: keys with count >= minOccurrence, ascending; written to a temp file of
    //      key words like the reference's sortedKeyFP.  The sorted (key, count) list is kept for
    //      loadKmer / outputOccurrenceTableBinary. --------------------------------------------------
    FILE *sortedKeyFromKmerFile(u64_t minOccurrence, const std::string &tmpDir = ".")
    {
        exportKmers(minOccurrence, true);
        FILE *fp = makeTemporaryFile(tmpDir);
        if (!keptKeys_.empty() && fwrite(keptKeys_.data(), 8, keptKeys_.size(), fp) != keptKeys_.size()) throw TMPError();
        return fp;
    }

    // ---- counter.h:600-640: the reference rebuilds a host DoubleHash here; on this path only its
    //      return value and (later) the slot placement matter, both produced by pbk_write_kmer_occ_bin.
    u64_t loadKmer(u64_t minOccurrence, u64_t doubleHashSize)
    {
        std::cerr << "loading kmers..." << std::endl;                     // counter.h:607
        if (keptMin_ != minOccurrence || !keptValid_) exportKmers(minOccurrence, true);
        doubleHashSize_ = doubleHashSize;
        const double total = (double)keptCounts_.size();
        u64_t size = (u64_t)(std::log(total / 0.9) / std::log(2.0));     // counter.h:621-622
        size = (u64_t)std::pow(2.0, (double)(size + 1));
        if (size > doubleHashSize) std::cerr << "WARNING:: Sorry, memory exceeds specified value!!" << std::endl;   // common.h:259-262
        loadSize_ = size;
        return size;
    }

    // ---- counter.h:955-963 + doubleHash.h:266-278 --------------------------------------------------
    void outputOccurrenceTableBinary(const std::string &filename)
    {
        if (!keptValid_) throw GPUError("outputOccurrenceTableBinary before loadKmer");
        uint64_t loadSize = 0;
        const int rc = pbk_write_kmer_occ_bin(filename.c_str(), (uint32_t)kmerLength_, keptKeys_.data(), keptCounts_.data(),
                                              keptCounts_.size(), doubleHashSize_, &loadSize);
        if (rc == PBK_E_IO) throw FILEError(filename);
        if (rc != PBK_OK) throw GPUError(pbk_strerror(rc));
    }

    // every kept (key, count), ascending when `sorted`: the hand-off to a host DoubleHash for the graph stage
    void exportKmers(u64_t minOccurrence, bool sorted)
    {
        uint64_t n = 0;
        const size_t W = (size_t)((kmerLength_ + 31) / 32);
        if (group_) {
            check(pbk_group_export(group_, (uint32_t)minOccurrence, sorted ? 1 : 0, NULL, NULL, 0, &n), "pbk_group_export");
            keptKeys_.assign(n * W, 0);
            keptCounts_.assign(n, 0);
            if (n) check(pbk_group_export(group_, (uint32_t)minOccurrence, sorted ? 1 : 0, keptKeys_.data(), keptCounts_.data(), n, &n), "pbk_group_export");
            keptMin_ = minOccurrence; keptValid_ = true;
            return;
        }
        check(pbk_export(ctx_, (uint32_t)minOccurrence, sorted ? 1 : 0, NULL, NULL, 0, &n), "pbk_export");
        keptKeys_.assign(n * W, 0);
        keptCounts_.assign(n, 0);
        if (n) check(pbk_export(ctx_, (uint32_t)minOccurrence, sorted ? 1 : 0, keptKeys_.data(), keptCounts_.data(), n, &n), "pbk_export");
        keptMin_ = minOccurrence; keptValid_ = true;
    }
    // ---- graph.h:337-375: the eight findValue probes makeInitialBruijnGraph makes per k-mer of sortedKeyFP, for all kept k-mers
    //      (the list of the last exportKmers / sortedKeyFromKmerFile) in one device pass: flags[i] = (leftFlags << 4) | rightFlags
    std::vector<uint8_t> neighborFlags()
    {
        if (!keptValid_) throw GPUError("neighborFlags before sortedKeyFromKmerFile / exportKmers");
        std::vector<uint8_t> flags(keptCounts_.size(), 0);
        if (flags.empty()) return flags;
        if (group_) check(pbk_group_neighbor_flags(group_, (uint32_t)keptMin_, keptKeys_.data(), flags.size(), flags.data()), "pbk_group_neighbor_flags");
        else check(pbk_neighbor_flags(ctx_, (uint32_t)keptMin_, keptKeys_.data(), flags.size(), flags.data()), "pbk_neighbor_flags");
        return flags;
    }
    const std::vector<uint64_t> &keptKeys() const { return keptKeys_; }
    const std::vector<uint16_t> &keptCounts() const { return keptCounts_; }

    // platanus::makeTemporaryFile (common.h:276-293): unlinked mkstemp file in the -tmp directory
    static FILE *makeTemporaryFile(const std::string &dir)
    {
        std::string templ = dir + "/XXXXXX";
        std::vector<char> buf(templ.begin(), templ.end());
        buf.push_back('\0');
        const int fd = mkstemp(buf.data());
        if (fd == -1) throw TMPError();
        FILE *fp = fdopen(fd, "wb+");
        unlink(buf.data());
        if (!fp) throw TMPError();
        return fp;
    }

private:
    void check(int rc, const char *what)
    {
        if (rc == PBK_OK) return;
        if (!ctx_ && !group_) throw GPUError(std::string(what) + ": no device context (several GPUs: only the counting members are available)");
        if (rc == PBK_E_READ_TOO_LONG) throw ReadError();                      // common.h:465
        if (rc == PBK_E_KMER_DIST) throw KmerDistError();
        std::string m = std::string(what) + ": " + pbk_strerror(rc);
        if (ctx_ && pbk_last_error(ctx_)[0]) m += std::string(" (") + pbk_last_error(ctx_) + ")";
        if (group_ && pbk_group_last_error(group_)[0]) m += std::string(" (") + pbk_group_last_error(group_) + ")";
        if (rc == PBK_E_BAD_BASE) throw ReadError(m);
        throw GPUError(m);
    }
    double average(const std::vector<u64_t> &dist, u64_t start, u64_t end) const
    {
        double out = 0;
        if (pbk_distribution_average((const uint64_t *)dist.data(), dist.size(), start, end, &out) != PBK_OK) throw KmerDistError();
        return out;
    }
    void flushPlatanus(std::vector<uint8_t> &bases, std::vector<uint64_t> &offsets, std::vector<int32_t> &npos,
                       std::vector<uint64_t> &nposOff)
    {
        const uint64_t n = offsets.size() - 1;
        if (n) {
            static const int32_t none = 0;
            check(pbk_push_reads(ctx_, bases.data(), offsets.data(), n, PBK_ENC_PLATANUS, npos.empty() ? &none : npos.data(), nposOff.data()),
                  "pbk_push_reads");
        }
        bases.clear(); npos.clear();
        offsets.assign(1, 0); nposOff.assign(1, 0);
    }

    pbk_ctx *ctx_;
    pbk_group *group_;
    u64_t kmerLength_, maxOccurrence_, doubleHashSize_, nInstances_, nDistinct_;
    int device_;
    unsigned flags_, numDevices_;
    std::vector<int32_t> deviceList_;
    std::vector<u64_t> lengthDistribution_, occurrenceDistribution_;
    std::vector<uint64_t> keptKeys_;
    std::vector<uint16_t> keptCounts_;
    u64_t keptMin_ = 0, loadSize_ = 0;
    bool keptValid_ = false;
};

}  // namespace pbk

#endif  // PBK_COUNTER_HPP
