// Parse throughput of the range-parallel FASTA/FASTQ reader (platanus_b_b200/host/pbk_ingest.hpp) on the host alone:
// the file is mapped, cut into T ranges, T threads parse their range into a sink that copies every read into a batch
// buffer (what pbk_assemble's BatchSink does) -- no GPU involved.  Prints one JSON line.
//   g++ -O2 -std=c++11 -pthread -o ingest_bench scripts/host/ingest_bench.cpp && ./ingest_bench reads.fq fq 8
#include "../../platanus_b_b200/host/pbk_ingest.hpp"

#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>

struct CopySink : pbk::ingest::ReadSink {
    std::vector<char> buf; size_t used, reads, bases;
    CopySink() : buf((size_t)64 << 20), used(0), reads(0), bases(0) {}
    void emit_span(const char *s, size_t len)
    {
        if (used + len > buf.size()) used = 0;                       // "hand the batch over"
        memcpy(buf.data() + used, s, len);
        used += len; reads += 1; bases += len;
    }
};

int main(int argc, char **argv)
{
    if (argc < 4) return 2;
    const bool fastq = std::string(argv[2]) == "fq";
    const unsigned T = (unsigned)atoi(argv[3]);
    pbk::ingest::MappedFile f;
    if (f.map(argv[1], "/tmp") != 0) return 3;
    double best = 1e30; size_t reads = 0, bases = 0;
    std::vector<CopySink> sinks(T < 1 ? 1 : (T > 64 ? 64 : T));      // batch buffers exist before the clock starts, as in pbk_assemble
    for (int rep = 0; rep < 3; ++rep) {
        for (auto &k : sinks) { k.used = k.reads = k.bases = 0; k.read.clear(); }
        const auto t0 = std::chrono::steady_clock::now();
        const pbk::ingest::Plan pl = pbk::ingest::plan_ranges(f.p, f.n, fastq, T);
        std::vector<std::thread> th;
        for (unsigned t = 0; t < pl.n_workers; ++t)
            th.push_back(std::thread([&, t]() {
                if (pl.s[t] < pl.s[t + 1] || t == pl.final_owner)
                    pbk::ingest::parse_range(f.p, pl.s[t], pl.s[t + 1], fastq, t == pl.final_owner, sinks[t]);
            }));
        for (auto &x : th) x.join();
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (s < best) best = s;
        reads = bases = 0;
        for (auto &k : sinks) { reads += k.reads; bases += k.bases; }
    }
    printf("{\"file_bytes\": %zu, \"threads\": %u, \"reads\": %zu, \"bases\": %zu, \"seconds\": %.4f, \"GB_per_s\": %.3f}\n",
           f.n, T, reads, bases, best, f.n / best / 1e9);
    return 0;
}
