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

#include <algorithm>
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
enum ErrorId { E_IO = 0, E_FOPEN = 1, E_TMP = 2, E_FORMAT = 3, E_READ = 4, E_KMERDIST = 6, E_GPU = 64, E_GPU_NOMEM = 65 };

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
struct GPUError : ErrorBase {
    explicit GPUError(const std::string &m, int id = E_GPU) : ErrorBase(id, "Error, GPU k-mer counter exception!!\n" + m) {}
};
// the table does not fit the HBM budget (PBK_E_NOMEM): what the counting members answer with hash-range passes
struct NoMemory : GPUError { explicit NoMemory(const std::string &m) : GPUError(m, E_GPU_NOMEM) {} };

class Counter {
public:
    typedef unsigned long long u64_t;

    FILE *kmerFP;      // kept for source compatibility (counter.h:77); the (key, count) stream stays in HBM

    Counter() : kmerFP(NULL), ctx_(NULL), group_(NULL), kmerLength_(0), maxOccurrence_(0), doubleHashSize_(0), nInstances_(0), nDistinct_(0),
                device_(-1), flags_(0), numDevices_(1) {}
    explicit Counter(u64_t k) : kmerFP(NULL), ctx_(NULL), group_(NULL), kmerLength_(k), maxOccurrence_(0), doubleHashSize_(0), nInstances_(0),
                                nDistinct_(0), device_(-1), flags_(0), numDevices_(1) {}
    ~Counter() { if (ctx_) pbk_destroy(ctx_); if (group_) pbk_group_destroy(group_); if (passFP_) fclose(passFP_); }
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
    // cap on the device memory of a context (default: 85 % of the free HBM; PBK_HBM_BUDGET_MB overrides both) and the directory
    // of the temporary file the hash-range passes collect their entries in (the reference's -tmp)
    void setHbmBudget(u64_t bytes) { hbmBudget_ = bytes; }
    void setTmpDir(const std::string &dir) { tmpDir_ = dir; }
    // how many hash-range passes the last count took (1 = the table fitted)
    unsigned getNumPasses() const { return passes_; }

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
    void beginCounting(u64_t kLength) { beginPass(kLength, 1, 0); multiPass_ = false; passes_ = 1; }
    // a context that counts hash range `pass` of `passes` (pbk_config.n_passes); passes == 1: everything
    void beginPass(u64_t kLength, unsigned passes, unsigned pass)
    {
        kmerLength_ = kLength;
        keptValid_ = false;
        if (group_) { check(pbk_group_reset(group_, (uint32_t)kLength), "pbk_group_reset"); return; }
        if (ctx_ && passes == 1 && !ctxIsPass_) { check(pbk_reset(ctx_, (uint32_t)kLength), "pbk_reset"); return; }
        dropContext();                                       // (a pass context is bound to its range: a new one per pass)
        pbk_config cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.struct_size = sizeof cfg; cfg.k = (uint32_t)kLength; cfg.device = device_; cfg.flags = flags_;
        cfg.hbm_budget_bytes = hbmBudget_;
        if (const char *e = getenv("PBK_HBM_BUDGET_MB")) { const long long mb = atoll(e); if (mb > 0) cfg.hbm_budget_bytes = (uint64_t)mb << 20; }
        if (passes > 1) { cfg.n_passes = passes; cfg.pass_index = pass; }
        ctxIsPass_ = passes > 1;
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
    // If the table does not fit the HBM budget (PBK_E_NOMEM) the count is repeated in 2, 4, 8 ... hash-range passes over the
    // same temp files -- the counterpart of the reference's own memory-limited mode, which writes the k-mers that found no
    // room to temporary files and counts them in further rounds (counter.h:340-364, 442-449).
    u64_t makeKmerReadDistributionMT(u64_t kLength, FILE **readFP, u64_t memory, u64_t numThread)
    {
        const SeqTempFeed feed = {readFP, numThread};
        return countFed(kLength, memory, feed, NULL, NULL, 0);
    }
    // feeder over the per-thread SEQ temp files (rewound and read once per pass)
    struct SeqTempFeed { FILE **fp; u64_t n; void operator()(Counter &c) const { c.pushSeqTempFiles(fp, n); } };
    // Counter::makeKmerReadDistributionConsideringPreviousGraph (counter.h:663-750) on the same files: the seeded k-mers keep
    // their value, every other k-mer of the reads is counted -- in hash-range passes, too, if the table does not fit
    u64_t makeKmerReadDistributionSeeded(u64_t kLength, FILE **readFP, u64_t memory, u64_t numThread, const uint64_t *seedKeys,
                                         const uint16_t *seedValues, uint64_t nSeeds)
    {
        const SeqTempFeed feed = {readFP, numThread};
        return countFed(kLength, memory, feed, seedKeys, seedValues, nSeeds);
    }
    // The general form: feed(*this) pushes ALL reads (pushReads / pushSeqTempFiles) and is called once per pass; seeds
    // (makeKmerReadDistributionConsideringPreviousGraph) are handed to every pass, the library keeps those of its range.
    template <typename Feed>
    u64_t countFed(u64_t kLength, u64_t memory, const Feed &feed, const uint64_t *seedKeys, const uint16_t *seedValues, uint64_t nSeeds)
    {
        unsigned passes = 1;
        if (const char *e = getenv("PBK_NUM_PASSES")) passes = (unsigned)std::max(1, atoi(e));       // (tests; a power of two is not required)
        for (;; passes *= 2) {
            try {
                if (passes == 1) {
                    beginCounting(kLength);
                    if (nSeeds) seedEntries(seedKeys, seedValues, nSeeds);
                    feed(*this);
                    return endCounting(memory);
                }
                return countInPasses(kLength, memory, passes, feed, seedKeys, seedValues, nSeeds);
            } catch (const NoMemory &) {
                dropContext();
                if (group_ || passes >= 4096) throw;
            }
        }
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
        if (multiPass_) { exportFromPassFile(minOccurrence, sorted); return; }
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
        if (multiPass_ && !ctx_) {
            // the passes' tables are gone; the kept entries (count >= the cutoff: what loadKmer puts into the host DoubleHash,
            // counter.h:600-640) fit -- they are what the graph stage works on
            beginPass(kmerLength_, 1, 0);
            keptValid_ = true;
            check(pbk_load_entries(ctx_, keptKeys_.data(), keptCounts_.data(), keptCounts_.size()), "pbk_load_entries");
            check(pbk_finalize(ctx_, NULL, NULL, NULL, NULL, NULL), "pbk_finalize");
        }
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
    void dropContext()
    {
        if (ctx_) { pbk_destroy(ctx_); ctx_ = NULL; }
        ctxIsPass_ = false;
    }
    // passes hash-range passes over everything feed pushes; every pass's entries go to an unlinked temporary file of
    // (key words, u16 count) records -- the layout of the reference's kmerFP (counter.h:494-495)
    template <typename Feed>
    u64_t countInPasses(u64_t kLength, u64_t memory, unsigned passes, const Feed &feed, const uint64_t *seedKeys, const uint16_t *seedValues,
                        uint64_t nSeeds)
    {
        const size_t W = (size_t)((kLength + 31) / 32);
        if (passFP_) { fclose(passFP_); passFP_ = NULL; }
        passFP_ = makeTemporaryFile(tmpDir_);
        std::vector<u64_t> occTotal(PBK_OCC_BINS, 0), occ(PBK_OCC_BINS, 0);
        lengthDistribution_.assign(PBK_LEN_BINS, 0);
        uint64_t ndTotal = 0, niTotal = 0;
        std::vector<uint64_t> keys;
        std::vector<uint16_t> counts;
        for (unsigned p = 0; p < passes; ++p) {
            beginPass(kLength, passes, p);
            if (nSeeds) seedEntries(seedKeys, seedValues, nSeeds);
            feed(*this);
            uint64_t nd = 0, ni = 0, mx = 0, n = 0;
            // (every pass sees every read: the read-length distribution is the same each time, the first pass's is kept)
            check(pbk_finalize(ctx_, (uint64_t *)occ.data(), p == 0 ? (uint64_t *)lengthDistribution_.data() : NULL, &nd, &ni, &mx), "pbk_finalize");
            for (size_t i = 0; i < occ.size(); ++i) occTotal[i] += occ[i];
            ndTotal += nd; niTotal += ni;
            check(pbk_export(ctx_, 1, 0, NULL, NULL, 0, &n), "pbk_export");
            keys.assign(n * W, 0); counts.assign(n, 0);
            if (n) check(pbk_export(ctx_, 1, 0, keys.data(), counts.data(), n, &n), "pbk_export");
            for (uint64_t i = 0; i < n; ++i)
                if (fwrite(&keys[i * W], 8, W, passFP_) != W || fwrite(&counts[i], 2, 1, passFP_) != 1) throw TMPError();
            dropContext();
        }
        occurrenceDistribution_.swap(occTotal);
        nDistinct_ = ndTotal; nInstances_ = niTotal;
        if (ndTotal) {
            for (size_t i = PBK_OCC_BINS - 1; i > 0; --i) if (occurrenceDistribution_[i]) { maxOccurrence_ = i; break; }     // counter.h:371-376
        }
        multiPass_ = true; passes_ = passes; keptValid_ = false;
        doubleHashSize_ = pbk_double_hash_size(memory, (uint32_t)kLength);
        return doubleHashSize_;
    }
    // exportKmers after a multi-pass count: the records of the pass file with count >= minOccurrence; sorted = ascending in
    // the reference's numeric order (top word first, binstr.h:460-466)
    void exportFromPassFile(u64_t minOccurrence, bool sorted)
    {
        const size_t W = (size_t)((kmerLength_ + 31) / 32);
        std::vector<uint64_t> keys;
        std::vector<uint16_t> counts;
        rewind(passFP_);
        std::vector<uint64_t> key(W);
        uint16_t c;
        while (fread(key.data(), 8, W, passFP_) == W) {
            if (fread(&c, 2, 1, passFP_) != 1) throw TMPError();
            if (c < minOccurrence) continue;
            keys.insert(keys.end(), key.begin(), key.end());
            counts.push_back(c);
        }
        const size_t n = counts.size();
        if (sorted && n > 1) {
            std::vector<size_t> order(n);
            for (size_t i = 0; i < n; ++i) order[i] = i;
            const uint64_t *kp = keys.data();
            std::sort(order.begin(), order.end(), [kp, W](size_t a, size_t b) {
                for (size_t j = W; j-- > 0;) if (kp[a * W + j] != kp[b * W + j]) return kp[a * W + j] < kp[b * W + j];
                return false;
            });
            keptKeys_.resize(n * W); keptCounts_.resize(n);
            for (size_t i = 0; i < n; ++i) {
                for (size_t j = 0; j < W; ++j) keptKeys_[i * W + j] = keys[order[i] * W + j];
                keptCounts_[i] = counts[order[i]];
            }
        } else {
            keptKeys_.swap(keys); keptCounts_.swap(counts);
        }
        keptMin_ = minOccurrence; keptValid_ = true;
    }
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
        if (rc == PBK_E_NOMEM) throw NoMemory(m);
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
    // hash-range passes (a table beyond the HBM budget)
    u64_t hbmBudget_ = 0;
    std::string tmpDir_ = ".";
    FILE *passFP_ = NULL;           // (key words, u16 count) records of all passes
    bool multiPass_ = false, ctxIsPass_ = false;
    unsigned passes_ = 1;
};

}  // namespace pbk

#endif  // PBK_COUNTER_HPP
