// pbk_assemble.cpp -- `platanus_b assemble -kmer_occ_only` on the B200 counter: the host side of the
// drop-in as a stand-alone program.  Same command line as the reference for this path
// (assemble.cpp:52-71, baseCommand.cpp:56-136), same stderr markers (assemble.cpp:306, 334, 663-665,
// 189; counter.h:607), same outputs: PREFIX_<k>merFrq.tsv and PREFIX_kmer_occ.bin.
//
//   pbk_assemble assemble -kmer_occ_only -k 32 -t 8 -m 16 -o out -f r_1.fq r_2.fq [-n N] [-repeat]
//
// What is different by design: the reads never take the temp-file detour (SEQ::writeTemporaryFile,
// common.h:426) -- one parser thread per input file hands ASCII batches in pinned memory to
// pbk_push_reads, and the GPU does the 2-bit packing.  `-t` bounds the parser threads; `-m` only
// determines doubleHashSize (the kmer_occ.bin header), as in counter.h:300-309.
#include "pbk_counter.hpp"
#include "pbk_ingest.hpp"

#include <chrono>
#include <cmath>
#include <condition_variable>
#include <deque>
#include <fstream>
#include <map>
#include <mutex>
#include <sstream>
#include <thread>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace {

typedef unsigned long long u64;

// PBK_TIMING=1: phase times on stderr (the reference has no timers at all)
struct PhaseTimer {
    bool on; std::chrono::steady_clock::time_point t0, last;
    PhaseTimer() : on(getenv("PBK_TIMING") != NULL), t0(std::chrono::steady_clock::now()), last(t0) {}
    void mark(const char *what)
    {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[pbk_assemble] %-28s %8.3f s  (total %.3f s)\n", what, std::chrono::duration<double>(now - last).count(),
                std::chrono::duration<double>(now - t0).count());
        last = now;
    }
};

// ---- options: the reference's three maps (assemble.cpp:52-71) -----------------------------------
struct Options {
    std::map<std::string, std::string> single;
    std::map<std::string, std::vector<std::string> > multi;
    std::map<std::string, bool> flag;
    Options()
    {
        single["-o"] = "out"; single["-k"] = "32"; single["-K"] = "0.5"; single["-s"] = "10"; single["-n"] = "0";
        single["-c"] = "1"; single["-a"] = "10.0"; single["-u"] = "0"; single["-d"] = "0.5"; single["-e"] = "";
        single["-t"] = "1"; single["-m"] = "16"; single["-tmp"] = ".";
        multi["-f"] = std::vector<std::string>();
        flag["-kmer_occ_only"] = false; flag["-repeat"] = false;
        flag["-seq_tmp"] = false;        // not a reference option: go through SEQ temp files + makeKmerReadDistributionMT
    }
    // baseCommand.cpp:56-136: unknown options and missing -f make the command print its usage
    bool parse(int argc, char **argv)
    {
        for (int i = 2; i < argc;) {
            const std::string a = argv[i];
            if (flag.count(a)) { flag[a] = true; ++i; }
            else if (single.count(a)) { if (i + 1 >= argc) return false; single[a] = argv[i + 1]; i += 2; }
            else if (multi.count(a)) {
                ++i;
                while (i < argc && argv[i][0] != '-') multi[a].push_back(argv[i++]);
            } else return false;
        }
        return !multi["-f"].empty();
    }
};

// ---- input files ---------------------------------------------------------------------------------
struct Mapped : pbk::ingest::MappedFile {                  // plain, gzip or bzip2 (pbk_ingest.hpp)
    Mapped(const std::string &name, const std::string &tmp_dir)
    {
        const int rc = map(name, tmp_dir);
        if (rc == -2) throw pbk::TMPError();
        if (rc != 0) throw pbk::FILEError(name);
    }
};

using pbk::ingest::Line;
using pbk::ingest::LineReader;
using pbk::ingest::ReadSink;

// BaseCommand::checkFileFormat (baseCommand.cpp:29-50): 0 unknown, 1 FASTA, 2 FASTQ
int check_file_format(const Mapped &f)
{
    LineReader r(f.p, f.n);
    Line l[4] = {{"", 0}, {"", 0}, {"", 0}, {"", 0}};
    for (int i = 0; i < 4; ++i) if (!r.next(l[i])) break;
    bool acgtn = true;
    for (size_t i = 0; i < l[1].len; ++i) acgtn &= (memchr("ACGTN", l[1].s[i], 5) != NULL);
    const char c0 = l[0].len ? l[0].s[0] : '\0', c2 = l[2].len ? l[2].s[0] : '\0';
    if (c0 == '>' && acgtn) return 1;
    if (c0 == '@' && acgtn && c2 == '+') return 2;
    return 0;
}

// one batch of reads in (pageable, page-aligned) host memory: the parsers must be able to run before the CUDA
// context exists -- creating it takes seconds on a 180 GB GPU and is the largest fixed cost of the program
struct Batch {
    uint8_t *bases; size_t cap, used;
    std::vector<uint64_t> offsets;
    Batch() : bases(NULL), cap(0), used(0), offsets(1, 0) {}
};

class BatchQueue {                       // parser threads -> the thread that owns the pbk context
public:
    BatchQueue(size_t batch_bytes, int n_buffers) : batch_bytes_(batch_bytes), producers_(0)
    {
        for (int i = 0; i < n_buffers; ++i) {
            Batch *b = new Batch();
            void *p = NULL;
            if (posix_memalign(&p, 4096, batch_bytes) != 0) throw pbk::ErrorBase(pbk::E_IO, "Error, out of host memory");
            b->bases = (uint8_t *)p; b->cap = batch_bytes;
            free_.push_back(b); all_.push_back(b);
        }
    }
    ~BatchQueue() { for (size_t i = 0; i < all_.size(); ++i) { free(all_[i]->bases); delete all_[i]; } }
    void add_producer() { std::lock_guard<std::mutex> g(m_); ++producers_; }
    void producer_done() { std::lock_guard<std::mutex> g(m_); --producers_; cv_.notify_all(); }
    Batch *get_free()
    {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [&] { return !free_.empty(); });
        Batch *b = free_.front(); free_.pop_front();
        b->used = 0; b->offsets.assign(1, 0);
        return b;
    }
    void put_full(Batch *b) { std::lock_guard<std::mutex> g(m_); full_.push_back(b); cv_.notify_all(); }
    Batch *get_full()                    // NULL once every producer has finished and nothing is queued
    {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [&] { return !full_.empty() || producers_ == 0; });
        if (full_.empty()) return NULL;
        Batch *b = full_.front(); full_.pop_front();
        return b;
    }
    void put_free(Batch *b) { std::lock_guard<std::mutex> g(m_); free_.push_back(b); cv_.notify_all(); }
    size_t batch_bytes() const { return batch_bytes_; }
private:
    size_t batch_bytes_;
    int producers_;
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<Batch *> free_, full_;
    std::vector<Batch *> all_;
};

// the fast path: ASCII batches in pinned memory; a read longer than a batch cannot exist (MAX_READ_LEN)
struct BatchSink : ReadSink {
    BatchQueue &q; Batch *cur;
    explicit BatchSink(BatchQueue &q_) : q(q_), cur(q_.get_free()) {}
    void emit_span(const char *s, size_t len)
    {
        if (len >= 500000) throw pbk::ReadError();                          // common.h:465
        if (cur->used + len > cur->cap) { q.put_full(cur); cur = q.get_free(); }
        memcpy(cur->bases + cur->used, s, len);
        cur->used += len;
        cur->offsets.push_back(cur->used);
    }
    void finish() { q.put_full(cur); cur = NULL; }
};

// the reference's own route (-seq_tmp): per-thread SEQ temp files, dealt round-robin (assemble.cpp:160-163,
// 836-845), later consumed by Counter::makeKmerReadDistributionMT
struct SeqTmpSink : ReadSink {
    std::vector<FILE *> &fp; size_t i;
    std::vector<int32_t> pos;
    std::string codes;
    explicit SeqTmpSink(std::vector<FILE *> &fp_) : fp(fp_), i(0) {}
    void emit_span(const char *rd, size_t rd_len)
    {
        if (rd_len >= 500000) throw pbk::ReadError();                  // common.h:465
        static const char code[] = ".\x0.\x1\x3..\x2......\x4";               // platanus::Char2Bin (common.h:256)
        pos.clear();
        codes.resize(rd_len);
        for (size_t j = 0; j < rd_len; ++j) {
            const char c = code[rd[j] & 0xF];
            if (c == 4) { pos.push_back((int32_t)j); codes[j] = 0; } else codes[j] = c;
        }
        const int32_t nn = (int32_t)pos.size(), len = (int32_t)rd_len;
        fwrite(&nn, 4, 1, fp[i]);
        if (nn) fwrite(pos.data(), 4, (size_t)nn, fp[i]);
        fwrite(&len, 4, 1, fp[i]);
        fwrite(codes.data(), 1, (size_t)len, fp[i]);
        i = (i + 1) % fp.size();
    }
};

// one file, serially (the -seq_tmp route deals reads round-robin in file order like the reference)
void parse_whole_file(const Mapped &f, bool fastq, ReadSink &out)
{
    const pbk::ingest::Plan pl = pbk::ingest::plan_ranges(f.p, f.n, fastq, 1);
    pbk::ingest::parse_range(f.p, pl.s[0], pl.s[1], fastq, true, out);
}

// ---- Assemble::extendKmer and friends (assemble.cpp:657-736): the stderr schedule ------------------
double log_probability_join(u64 cutoff, double cov, double len, u64 large_k, u64 small_k)
{
    const double c = cov * (len - large_k + 1.0) / len;
    double s = 0;
    for (u64 i = 0; i < cutoff; ++i) {
        double p = 0;
        for (u64 j = 1; j <= i; ++j) p += std::log(c) - std::log((double)j);
        s += std::exp(p);
    }
    s = std::exp(-c + std::log(s));
    return ((large_k - small_k) + 1.0) * (-s);
}
u64 decrease_cutoff(u64 cutoff, double cov, double len, double min_log_p, u64 large_k, u64 small_k)
{
    if (cutoff <= 1) return 1;
    u64 i = cutoff;
    for (; i > 1; --i) if (log_probability_join(i, cov, len, large_k, small_k) > min_log_p) break;
    return i;
}
void extend_kmer(double min_log_p, double cov, double len, u64 min_cov, double max_k_ratio, u64 step, std::vector<unsigned> &kmer,
                 std::vector<u64> &cutoff)
{
    const u64 min_max_k = (u64)(long)(len * max_k_ratio + 0.5);
    std::cerr << "\nKMER_EXTENSION:" << std::endl;
    std::cerr << "K=" << kmer[0] << ", KMER_COVERAGE=" << cov * (len - kmer[0] + 1.0) / len;
    std::cerr << " (>= " << cutoff[0] << "), COVERAGE_CUTOFF=" << cutoff[0] << std::endl;
    for (u64 i = 1; kmer[i - 1] <= len; ++i) {
        kmer.push_back(0); cutoff.push_back(0);
        for (u64 j = 1; j <= step + 1; ++j) {
            kmer[i] = kmer[i - 1] + (unsigned)j;
            cutoff[i] = std::max(decrease_cutoff(cutoff[i - 1], cov, len, min_log_p, kmer[i], kmer[i - 1]), min_cov);
            if (kmer[i - 1] + j > min_max_k && log_probability_join(cutoff[i], cov, len, kmer[i], kmer[i - 1]) < min_log_p) break;
        }
        --kmer[i];
        cutoff[i] = std::max(decrease_cutoff(cutoff[i - 1], cov, len, min_log_p, kmer[i], kmer[i - 1]), min_cov);
        if (kmer[i] == kmer[i - 1]) break;
        std::cerr << "K=" << kmer[i] << ", KMER_COVERAGE=" << cov * (len - kmer[i] + 1.0) / len;
        std::cerr << ", COVERAGE_CUTOFF=" << cutoff[i] << ", PROB_SPLIT=10e"
                  << std::log10(1.0 - std::exp(log_probability_join(cutoff[i], cov, len, kmer[i], kmer[i - 1]))) << std::endl;
    }
}

void usage()
{
    std::cerr << "\nUsage: pbk_assemble assemble -kmer_occ_only [Options]\nOptions:\n"
              << "    -o STR               : prefix of output files (default out)\n"
              << "    -f FILE1 [FILE2 ...] : reads file (fasta or fastq; plain, gzip or bzip2)\n"
              << "    -k INT               : k-mer size (default 32)\n"
              << "    -n INT               : initial k-mer coverage cutoff (default 0, 0 means auto)\n"
              << "    -t INT               : number of threads (parsers; default 1)\n"
              << "    -m INT               : memory limit for making kmer distribution (GB, default 16; sets the table size in the .bin header)\n"
              << "    -tmp DIR             : directory for temporary files (default .)\n"
              << "    -repeat              : mode to assemble repetitive sequences\n"
              << "    -kmer_occ_only       : only output k-mer occurrence table (required: the graph stages are not part of this tool)\n";
}

void exec(Options &opt)
{
    const bool repeat = opt.flag["-repeat"];
    const double min_log_p = std::log(1.0 - std::pow(10.0, -atof(opt.single["-a"].c_str())));   // assemble.cpp:117
    const u64 num_thread = (u64)std::max(1, atoi(opt.single["-t"].c_str()));
    const u64 step = (u64)atoi(opt.single["-s"].c_str());
    const double max_k_ratio = atof(opt.single["-K"].c_str());
    const u64 min_cov = (u64)atoi(opt.single["-c"].c_str());
    const u64 memory = (u64)atoi(opt.single["-m"].c_str()) * 1000000000ull;                     // assemble.cpp:124
    const unsigned k0 = (unsigned)atoi(opt.single["-k"].c_str());
    const std::string prefix = opt.single["-o"];
    const std::vector<std::string> &files = opt.multi["-f"];
    const std::string tmp_dir = opt.single["-tmp"];

    PhaseTimer pt;
    pbk::Counter counter;
    counter.setTmpDir(tmp_dir);
    // open (and, for gzip/bzip2 inputs, decompress: one `gzip -cd` per file, -t at a time) and sniff the inputs; the
    // first failure in file order is the one reported, as in the reference's serial loop (assemble.cpp:162-163)
    std::vector<int> types(files.size(), 0);
    std::vector<Mapped *> maps(files.size(), (Mapped *)NULL);
    {
        std::vector<pbk::ErrorBase *> open_err(files.size(), (pbk::ErrorBase *)NULL);
        std::vector<std::thread> openers;
        std::mutex m;
        size_t next = 0;
        for (size_t w = 0; w < std::min<size_t>((size_t)num_thread, files.size()); ++w)
            openers.push_back(std::thread([&]() {
                for (;;) {
                    size_t i;
                    { std::lock_guard<std::mutex> g(m); i = next++; }
                    if (i >= files.size()) return;
                    try {
                        maps[i] = new Mapped(files[i], tmp_dir);
                        types[i] = check_file_format(*maps[i]);
                        if (types[i] == 0) throw pbk::ReadError("Read file is not FASTA/FASTQ format.");   // assemble.cpp:798
                    } catch (pbk::ErrorBase &e) { open_err[i] = new pbk::ErrorBase(e); }
                }
            }));
        for (size_t w = 0; w < openers.size(); ++w) openers[w].join();
        for (size_t i = 0; i < files.size(); ++i) if (open_err[i]) throw *open_err[i];
    }
    u64 double_hash_size = 0;
    // narrow seam: exactly the reference's data flow up to Counter::makeKmerReadDistributionMT -- the reads go through the SEQ
    // temp files (-seq_tmp), which is also where the streaming path below falls back to when the table does not fit the HBM
    // budget: the temp files can be read once per hash-range pass (pbk::Counter::countFed), a parsed stream cannot
    auto count_through_temp_files = [&](bool announce) {
        std::vector<FILE *> read_fp;
        for (u64 i = 0; i < num_thread; ++i) read_fp.push_back(pbk::Counter::makeTemporaryFile(opt.single["-tmp"]));
        SeqTmpSink sink(read_fp);
        for (size_t i = 0; i < files.size(); ++i) parse_whole_file(*maps[i], types[i] == 2, sink);
        if (announce) std::cerr << "K = " << k0 << ", saving kmers from reads..." << std::endl;
        double_hash_size = counter.makeKmerReadDistributionMT(k0, read_fp.data(), memory, num_thread);
        for (size_t i = 0; i < read_fp.size(); ++i) fclose(read_fp[i]);
        if (counter.getNumPasses() > 1 && getenv("PBK_TIMING")) std::cerr << "[pbk] table beyond the HBM budget: counted in " << counter.getNumPasses() << " hash-range passes" << std::endl;
    };
    if (opt.flag["-seq_tmp"]) {
        count_through_temp_files(true);
        for (size_t i = 0; i < maps.size(); ++i) delete maps[i];
    } else {
    pt.mark("open + sniff inputs");
    std::cerr << "K = " << k0 << ", saving kmers from reads..." << std::endl;                   // assemble.cpp:306

    // ingest: the files are cut into byte ranges that are parsed independently (pbk_ingest.hpp), -t workers at a
    // time, started BEFORE the CUDA context is created so that parsing runs behind that fixed cost; the main thread
    // then feeds the GPU.  Enough batch buffers for ~2 GB of bases.
    size_t total_bytes = 0;
    for (size_t i = 0; i < maps.size(); ++i) total_bytes += maps[i]->n;
    // PBK_INGEST_RANGE_BYTES: smallest byte range handed to one worker (default 32 MiB; tests set it low so that small
    // files are cut into many ranges too)
    size_t range_min = (size_t)32 << 20;
    if (const char *e = getenv("PBK_INGEST_RANGE_BYTES")) { const long long v = atoll(e); if (v > 0) range_min = (size_t)v; }
    const size_t range_target = std::max<size_t>(range_min, total_bytes / (size_t)(num_thread * 2) + 1);
    struct Item { size_t file; unsigned t; };
    std::vector<pbk::ingest::Plan> plans;
    std::vector<Item> items;
    for (size_t i = 0; i < maps.size(); ++i) {
        const unsigned T = (unsigned)std::min<size_t>(64, std::max<size_t>(1, (maps[i]->n + range_target - 1) / range_target));
        plans.push_back(pbk::ingest::plan_ranges(maps[i]->p, maps[i]->n, types[i] == 2, T));
        for (unsigned t = 0; t < T; ++t)
            if (plans[i].s[t] < plans[i].s[t + 1] || t == plans[i].final_owner) items.push_back(Item{i, t});
    }
    // every worker holds one batch buffer while it parses: at most 12 workers on 16 buffers, so that full batches can
    // queue up behind the CUDA context creation without a worker waiting for a buffer to start with
    const size_t n_par = std::max<size_t>(1, std::min<size_t>(std::min<size_t>((size_t)num_thread, 12), items.size()));
    const size_t batch_bytes = (size_t)128 << 20;
    const int n_buf = (int)std::min<size_t>(16, std::max<size_t>(n_par * 2 + 1, total_bytes / 2 / batch_bytes + n_par + 1));
    BatchQueue q(batch_bytes, n_buf);
    std::mutex err_m;
    std::vector<pbk::ErrorBase> errors;
    std::vector<std::thread> workers;
    size_t next_item = 0;
    std::mutex next_m;
    for (size_t w = 0; w < n_par; ++w) {
        q.add_producer();
        workers.push_back(std::thread([&]() {
            BatchSink sink(q);
            for (;;) {
                size_t i;
                { std::lock_guard<std::mutex> g(next_m); i = next_item++; }
                if (i >= items.size()) break;
                const Item it = items[i];
                const pbk::ingest::Plan &pl = plans[it.file];
                try {
                    pbk::ingest::parse_range(maps[it.file]->p, pl.s[it.t], pl.s[it.t + 1], types[it.file] == 2, it.t == pl.final_owner, sink);
                } catch (pbk::ErrorBase &e) { std::lock_guard<std::mutex> g(err_m); errors.push_back(e); sink.read.clear(); }
            }
            sink.finish();
            q.producer_done();
        }));
    }
    // How many GPUs: PBK_NUM_GPUS (a number, or "all"), else one per 2 GiB of input text, never more than are visible
    // (CUDA_VISIBLE_DEVICES chooses which; there is no new command-line flag, iterate.cpp:244-251 spawns a fixed command line).
    {
        const int visible = std::max(1, pbk_device_count());
        unsigned want = (unsigned)std::max<size_t>(1, total_bytes >> 31);
        if (const char *e = getenv("PBK_NUM_GPUS")) want = strcmp(e, "all") == 0 ? (unsigned)visible : (unsigned)std::max(1, atoi(e));
        if (getenv("PBK_GROUP_LOGICAL") && want > (unsigned)visible) {      // tests: more shards than GPUs, several members per device
            std::vector<int32_t> devs;
            for (unsigned i = 0; i < std::min(want, 16u); ++i) devs.push_back((int32_t)(i % (unsigned)visible));
            counter.setDevices(devs);
        } else {
            counter.setNumDevices(std::min<unsigned>(want, (unsigned)visible));
        }
        if (counter.getNumDevices() > 1 && getenv("PBK_TIMING")) std::cerr << "[pbk] counting on " << counter.getNumDevices() << " GPUs" << std::endl;
    }
    pbk::ErrorBase *push_error = NULL;
    try { counter.beginCounting(k0); } catch (pbk::ErrorBase &e) { push_error = new pbk::ErrorBase(e); }
    pt.mark("pbk_create (CUDA context)");
    while (Batch *b = q.get_full()) {
        if (!push_error && b->offsets.size() > 1) {
            try { counter.pushReads(b->bases, b->offsets.data(), b->offsets.size() - 1); }
            catch (pbk::ErrorBase &e) { push_error = new pbk::ErrorBase(e); }
        }
        q.put_free(b);
    }
    for (size_t w = 0; w < workers.size(); ++w) workers[w].join();
    if (!errors.empty()) throw errors[0];
    bool beyond_budget = push_error && push_error->getID() == pbk::E_GPU_NOMEM && counter.getNumDevices() == 1;
    if (push_error && !beyond_budget) throw *push_error;
    pt.mark("parse + push (overlapped)");
    if (!beyond_budget) {
        try { double_hash_size = counter.endCounting(memory); }
        catch (pbk::NoMemory &) { if (counter.getNumDevices() > 1) throw; beyond_budget = true; }
    }
    if (beyond_budget) count_through_temp_files(false);        // the table does not fit: hash-range passes over the SEQ temp files
    for (size_t i = 0; i < maps.size(); ++i) delete maps[i];
    pt.mark("pbk_finalize");
    }

    // assemble.cpp:318-335
    std::vector<u64> cutoff;
    const int n_opt = atoi(opt.single["-n"].c_str());
    cutoff.push_back(n_opt != 0 ? (u64)n_opt
                                : std::max<u64>(repeat ? counter.getLeftLocalMinimalValue(1) : counter.getLeftLocalMinimalValue(1) / 2, 2ull));
    double ave_cov = counter.calcOccurrenceDistributionAverage(cutoff[0], counter.getMaxOccurrence());
    if (opt.single["-e"] != "") ave_cov = atof(opt.single["-e"].c_str());
    const double ave_len = counter.calcLengthDistributionAverage(0, 500000);
    ave_cov = ave_cov * ave_len / (ave_len - k0 + 1.0);
    std::cerr << "AVE_READ_LEN=" << ave_len << std::endl;
    std::vector<unsigned> kmer(1, k0);
    extend_kmer(min_log_p, ave_cov, ave_len, min_cov, max_k_ratio, step, kmer, cutoff);

    std::ostringstream oss;
    oss << prefix << '_' << k0 << "merFrq.tsv";
    counter.outputOccurrenceDistribution(oss.str());

    pt.mark("cutoff + schedule + tsv");
    FILE *sorted_fp = counter.sortedKeyFromKmerFile(cutoff[0], opt.single["-tmp"]);
    pt.mark("export sorted + sortedKeyFP");
    double_hash_size = counter.loadKmer(cutoff[0], double_hash_size);
    (void)double_hash_size;
    if (opt.flag["-kmer_occ_only"]) {
        counter.outputOccurrenceTableBinary(prefix + "_kmer_occ.bin");
        pt.mark("kmer_occ.bin");
        fclose(sorted_fp);
        std::cerr << "assemble completed!" << std::endl;                                       // assemble.cpp:189
        // outputs are closed; skip the teardown of the CUDA context (the OS reclaims it faster than the driver frees it)
        fflush(NULL);
        _exit(0);
    }
    fclose(sorted_fp);
    throw pbk::ErrorBase(13, "pbk_assemble only implements the -kmer_occ_only path; run the reference for graph construction");
}

}  // namespace

int main(int argc, char **argv)
{
    if (argc < 2 || strcmp(argv[1], "assemble") != 0) { usage(); return 1; }
    Options opt;
    if (!opt.parse(argc, argv)) { usage(); return 1; }
    try {
        exec(opt);
    } catch (pbk::ErrorBase &e) {        // main.cpp:121-124: message on stderr, exit code = error id
        e.showErrorMessage();
        return e.getID();
    }
    return 0;
}
